"""End-to-end time of Engine.pairs for pinned vs ordinary (pageable) numpy frames, against the number of staging threads."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyorc_b200.engine import Engine
from pyorc_b200 import synth
e = Engine(0)
H, W, n = 1080, 1920, 101
fr = synth.particle_frames_torch(n, H, W, torch.device("cuda", 0), dtype="uint8")
pinned = e.pinned_empty((n, H, W), np.uint8)
pinned[...] = fr.cpu().numpy()
pageable = np.array(pinned, copy=True)
def run(name, host):
    for _ in range(2):
        e.pairs(host, (64, 64), (32, 32))
    t0 = time.perf_counter()
    for _ in range(10):
        e.pairs(host, (64, 64), (32, 32))
    dt = (time.perf_counter() - t0) / 10
    print(f"{name:24s}: {dt * 1e3:.3f} ms per 100-pair step -> {100 * 32 * 59 / dt / 1e6:.2f} Mwin/s end to end", flush=True)
run("pinned", pinned)
for nt in (1, 2, 4, 8, 12, 16):
    e.set_option("stage_threads", nt)
    run(f"pageable, {nt} threads", pageable)
print("host cores", os.cpu_count())
