"""Where does a get_b2piv call on pageable numpy frames spend its time? (development aid)"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyorc_b200 import _xr, synth, velocimetry
from pyorc_b200.engine import get_engine

H, W, WS, OV, N = 1080, 1920, (64, 64), (32, 32), 101
dev = torch.device("cuda", 0)
host = synth.particle_frames_torch(N, H, W, dev, dtype="uint8").cpu().numpy()
da = _xr.DataArray(host, ("time", "y", "x"), {"time": np.arange(N) / 30.0})
eng = get_engine(0)
nr, nc = eng.plan((H, W), WS, OV, np.uint8)
args = (da, np.arange(nr), np.arange(nc), np.full(N - 1, 1 / 30.0), WS, OV, WS, 0.01, 0.01)
for _ in range(3):
    velocimetry.get_b2piv(*args)
t0 = time.perf_counter()
for _ in range(10):
    velocimetry.get_b2piv(*args)
print("get_b2piv pageable: %.2f ms" % ((time.perf_counter() - t0) * 100))
t0 = time.perf_counter()
for _ in range(10):
    eng.pairs(host, WS, OV)
print("Engine.pairs pageable: %.2f ms" % ((time.perf_counter() - t0) * 100))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    velocimetry.get_b2piv(*args)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
from pyorc_b200 import window
t0 = time.perf_counter()
for _ in range(20):
    window.available_memory(0)
print("available_memory: %.3f ms" % ((time.perf_counter() - t0) * 50))
pin = eng.pinned_empty(host.shape, np.uint8) if hasattr(eng, "pinned_empty") else None
if pin is not None:
    pin[...] = host
    t0 = time.perf_counter()
    for _ in range(10):
        eng.pairs(pin, WS, OV)
    print("Engine.pairs page-locked: %.2f ms" % ((time.perf_counter() - t0) * 100))
for thr in (4, 8, 12, 16):
    eng.set_option("stage_threads", thr)
    eng.pairs(host, WS, OV)
    t0 = time.perf_counter()
    for _ in range(10):
        eng.pairs(host, WS, OV)
    print("Engine.pairs pageable, %d stage threads: %.2f ms" % (thr, (time.perf_counter() - t0) * 100))
