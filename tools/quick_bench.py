"""Quick device-resident timing of the fused PIV kernel (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyorc_b200.engine import Engine
from pyorc_b200 import synth

def run(H, W, ws, ov, n_frames, reps=5, dtype="uint8"):
    dev = torch.device("cuda", 0)
    e = Engine(0)
    fr = synth.particle_frames_torch(n_frames, H, W, dev, dtype=dtype)
    torch.cuda.synchronize()
    for _ in range(3):
        out = e.pairs(fr, ws, ov)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = e.pairs(fr, ws, ov); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    nr, nc = out[0].shape[1:]
    nwin = (n_frames - 1) * nr * nc
    t = float(np.median(ts))
    print(f"{H}x{W} win {ws} ov {ov} {dtype}: {nwin} windows in {t:.3f} ms -> {nwin / t / 1e3:.2f} Mwin/s  (u mean {float(torch.nanmean(out[0])):.3f} v mean {float(torch.nanmean(out[1])):.3f})", flush=True)
    e.close()

if __name__ == "__main__":
    if "--single" in sys.argv:
        run(1080, 1920, (64, 64), (32, 32), 21, reps=2)
        sys.exit(0)
    run(1080, 1920, (64, 64), (32, 32), 101)
    run(1080, 1920, (64, 64), (32, 32), 101, dtype="float32")
    run(1080, 1920, (32, 32), (16, 16), 41)
    run(1080, 1920, (32, 32), (24, 24), 11)
    run(2160, 3840, (128, 128), (64, 64), 21)
    run(1080, 1920, (16, 16), (8, 8), 11)
