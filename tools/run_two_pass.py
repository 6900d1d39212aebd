"""Two-pass PIV of BASELINE configs[2] on 1080p: discrete window offset vs window deformation, coarse pass at 75 % / 50 % overlap
(development aid; bench.py reports the same rows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyorc_b200.engine import Engine
from pyorc_b200 import synth
from oracle import ffpiv_oracle as O

H, W, N = 1080, 1920, 41
dev = torch.device("cuda", 0)
e = Engine(0)
fr = synth.particle_frames_torch(N, H, W, dev, dtype="uint8")
FINE = ((32, 32), (24, 24))
y0, x0 = O.window_origins((H, W), *FINE)
yc, xc = np.meshgrid(y0 + 16.0, x0 + 16.0, indexing="ij")
tx, ty = synth.displacement_field(H, W, yc, xc)

def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), out

def report(name, ms, out):
    u, v = out[0].cpu().numpy(), out[1].cpu().numpy()
    nwin = u.size
    rmse = np.sqrt(np.nanmean((u - tx[None]) ** 2 + (v - ty[None]) ** 2))
    print(f"{name:58s} {ms:7.3f} ms  {nwin / ms / 1e3:7.1f} M fine windows/s  rmse vs imposed field {rmse:.4f} px  finite {np.isfinite(u).mean():.4f}  corr {float(torch.nanmean(out[2])):.4f}", flush=True)

ms, out = timed(lambda: e.pairs(fr, *FINE)); report("single pass 32x32 / 75 %", ms, out)
for coarse in (((64, 64), (48, 48)), ((64, 64), (32, 32))):
    for mode in ("offset", "deform"):
        ms, out = timed(lambda: e.pairs_two_pass(fr, coarse, FINE, mode=mode))
        report(f"two-pass {mode}, coarse {coarse[0][0]}x{coarse[0][1]} / {100 * coarse[1][0] // coarse[0][0]} %", ms, out)
