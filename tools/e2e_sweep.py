"""End-to-end (pinned host frames -> H2D -> kernel -> D2H) time of Engine.pairs against the number of copy chunks."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyorc_b200.engine import Engine
from pyorc_b200 import synth

if __name__ == "__main__":
    e = Engine(0)
    H, W, n = 1080, 1920, 101
    fr = synth.particle_frames_torch(n, H, W, torch.device("cuda", 0), dtype="uint8")
    host = e.pinned_empty((n, H, W), np.uint8)
    host[...] = fr.cpu().numpy()
    for chunks in (1, 4, 8, 12, 16, 20, 25, 33, 50):
        e.set_option("copy_chunks", chunks)
        for _ in range(3):
            e.pairs(host, (64, 64), (32, 32))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            out = e.pairs(host, (64, 64), (32, 32))
        dt = (time.perf_counter() - t0) / 20
        print(f"copy_chunks {chunks:3d}: {dt * 1e3:.3f} ms per step -> {100 * 32 * 59 / dt / 1e6:.2f} Mwin/s end to end (kernel {e.last_kernel_ms:.3f} ms)", flush=True)
    e.close()
