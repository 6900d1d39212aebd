"""Device-resident timing of the mask / packing kernels against the HBM copy peak.

Fields: the result grid of BASELINE configs[2] (132 x 237 windows), 1000 time steps = 31.3 M elements (125 MB) per field.
Algorithmic bytes: every input field read once (twice where the reference's two-pass std needs it), the mask written once.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyorc_b200 import mask as M

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

if __name__ == "__main__":
    peak = 6568.0
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p): peak = float(json.load(open(p))["hbm_gbs"])
    T, ny, nx = (int(v) for v in os.environ.get("B2_MASK_SHAPE", "1000,132,237").split(","))
    g = torch.Generator(device="cuda").manual_seed(3)
    vx = 0.6 + 0.35 * torch.randn((T, ny, nx), device="cuda", generator=g)
    vy = -0.2 + 0.25 * torch.randn((T, ny, nx), device="cuda", generator=g)
    c = torch.rand((T, ny, nx), device="cuda", generator=g)
    s = 1 + 39 * torch.rand((T, ny, nx), device="cuda", generator=g)
    bad = torch.rand((T, ny, nx), device="cuda", generator=g) < 0.1
    for a in (vx, vy, c, s): a[bad] = float("nan")
    n = T * ny * nx
    m3 = M.minmax(vx, vy)
    rows = [
        ("minmax", timeit(lambda: M.minmax(vx, vy)), n * 9),
        ("angle", timeit(lambda: M.angle(vx, vy)), n * 9),
        ("corr", timeit(lambda: M.corr(c)), n * 5),
        ("count", timeit(lambda: M.count(vx)), n * 4),
        ("outliers", timeit(lambda: M.outliers(vx, vy)), n * (16 + 9)),
        ("variance", timeit(lambda: M.variance(vx, vy)), n * 16),
        ("rolling wdw=5", timeit(lambda: M.rolling(vx, vy)), n * 9),
        ("window_nan wdw=1", timeit(lambda: M.window_nan(vx)), n * 5),
        ("window_mean wdw=1", timeit(lambda: M.window_mean(vx, vy)), n * 9),
        ("window_replace x4 (incl. clone)", timeit(lambda: M.window_replace([vx, vy, c, s])), n * 4 * (8 + 8 + 8)),
        ("apply x4 (incl. clone)", timeit(lambda: M.apply_masks([vx, vy, c, s], [m3])), n * (4 * 8 + 1 + 4 * 4 * 0.1)),
        ("encode_int16", timeit(lambda: M.encode_int16(vx)), n * 6),
    ]
    for name, ms, byts in rows:
        gbs = byts / ms / 1e6
        print(f"{name:32s} {ms:8.3f} ms  {gbs:8.1f} GB/s algorithmic  = {gbs / peak:.3f} of measured HBM copy peak {peak:.0f} GB/s", flush=True)
    if "--cpu" in sys.argv:
        from oracle import mask_oracle as MO
        hx, hy = vx[:100].cpu().numpy(), vy[:100].cpu().numpy()
        for name, fn in (("minmax", lambda: MO.minmax(hx, hy)), ("outliers", lambda: MO.outliers(hx, hy)), ("window_mean", lambda: MO.window_mean(hx, hy))):
            t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
            print(f"cpu numpy restatement {name:12s} {dt * 1e3 * T / 100:9.1f} ms (scaled from 100 of {T} time steps, 1 core)", flush=True)
