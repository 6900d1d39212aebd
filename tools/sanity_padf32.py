"""Small launches of the padded float32 rows kernels (per time step and ensemble, both plane sizes) for compute-sanitizer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyorc_b200.engine import Engine
from pyorc_b200 import synth
e = Engine(0)
for ws, ov, shape in (((26, 26), (12, 12), (3, 96, 127)), ((10, 14), (5, 7), (3, 60, 83))):
    imgs = synth.particle_frames(*shape, dtype=np.float32)
    u, v, c, s = e.pairs(imgs, ws, ov)
    assert e.last_variant == 4
    e.ens_begin(shape[1:], ws, ov, np.float32)
    e.ens_add(imgs, ws, ov, corr_min=0.0, s2n_min=0.0)
    e.ens_finish(0.0)
    print(ws, "ok", float(np.nanmean(u)), float(np.nanmean(v)), flush=True)
e.close()
