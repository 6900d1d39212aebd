"""The reference's shipped mask example (examples/ngwerere/ngwerere_piv.nc -> ngwerere_masked.nc, notebook 03) through the DEVICE mask
stack: `pyorc_b200.mask.Masks` with the accessor calls of the notebook, on tests/golden/ngwerere_masks.npz.  Prints how many of the
486 750 values differ from the reference's output (the CPU restatement, oracle/mask_oracle.py, reproduces it exactly:
tests/test_mask.py::test_oracle_reproduces_the_reference_mask_example_exactly; the kernels equal that restatement bit for bit on the
test fields, tests/test_mask.py -m gpu).  Needs a GPU; development aid, not part of the test suite (added after the round's GPU budget
was spent, so it has not run on hardware yet)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyorc_b200 import _xr  # noqa: E402
from pyorc_b200.mask import Masks, decode_int16, pack_dataset  # noqa: E402

if __name__ == "__main__":
    d = np.load(os.path.join(ROOT, "tests", "golden", "ngwerere_masks.npz"))
    shape = tuple(int(v) for v in d["shape"])
    kept = np.unpackbits(d["kept_bits"])[: int(np.prod(shape))].reshape(shape).astype(bool)
    f = {k: np.asarray(decode_int16(d[k])) for k in ("v_x", "v_y", "corr")}
    f["s2n"] = np.ones(shape, np.float32)          # not used by these masks (the file's s2n holds int16 wrap-around values)
    dims = ("time", "y", "x")
    ds = _xr.Dataset({k: (dims, f[k]) for k in ("v_x", "v_y", "corr", "s2n")}, {"time": d["time"], "y": d["y"], "x": d["x"]})
    mk = Masks(ds)
    mk.corr(inplace=True)
    mk.minmax(inplace=True)
    mk.rolling(inplace=True)
    mk.outliers(inplace=True)
    mk.variance(inplace=True)
    mk.angle(angle_tolerance=0.5 * np.pi)          # not in place, as in the notebook
    mk.count(inplace=True)
    mk.window_mean(wdw=2, inplace=True, tolerance=0.5, reduce_time=True)
    got = np.isfinite(np.asarray(ds["v_x"].values))
    print("survivors", int(got.sum()), "reference", int(kept.sum()), "differing values", int((got != kept).sum()), "of", kept.size)
    packed = pack_dataset(ds)
    same = all(np.array_equal(np.asarray(packed[k])[kept], d[k][kept]) for k in ("v_x", "v_y", "corr"))
    print("survivors re-encode to the file's int16 values:", same)
    sys.exit(0 if (got == kept).all() and same else 1)
