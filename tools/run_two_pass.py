"""Two-pass timing (development aid): repeated device-resident runs + per-stage CUDA-event times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyorc_b200.engine import Engine
from pyorc_b200 import synth
e = Engine(0)
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 41
fr = synth.particle_frames_torch(n, 1080, 1920, dev, dtype="uint8")
coarse, fine = ((64, 64), (48, 48)), ((32, 32), (24, 24))
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(6):
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record(); u1, v1, _, _ = e.pairs(fr, *coarse)
    b.record(); sh = e.predictor(u1, v1, (1080, 1920), coarse, fine)
    c.record(); out = e.pairs_shifted(fr, *fine, sh)
    d.record(); torch.cuda.synchronize()
    print(f"rep {rep}: pass1 {a.elapsed_time(b):.3f} ms  predictor {b.elapsed_time(c):.3f} ms  pass2 {c.elapsed_time(d):.3f} ms  total {a.elapsed_time(d):.3f} ms  -> {out[0].numel() / a.elapsed_time(d) / 1e3:.1f} Mwin/s", flush=True)
