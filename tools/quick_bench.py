"""Quick device-resident timing of the fused PIV kernel (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyorc_b200.engine import Engine
from pyorc_b200 import synth

def run(H, W, ws, ov, n_frames, reps=5, dtype="uint8", variant=0, run_len=0, **opts):
    """opts: further engine options (name=value) for A/B runs."""
    dev = torch.device("cuda", 0)
    e = Engine(0)
    e.set_option("kernel_variant", variant)
    e.set_option("run_len", run_len)
    for k, v in opts.items():
        e.set_option(k, v)
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
    print(f"variant {variant} run_len {run_len} {opts if opts else ""} {H}x{W} win {ws} ov {ov} {dtype}: {nwin} windows in {t:.3f} ms -> {nwin / t / 1e3:.2f} Mwin/s  (u mean {float(torch.nanmean(out[0])):.3f} v mean {float(torch.nanmean(out[1])):.3f})", flush=True)
    e.close()

def run_two_pass(H, W, n_frames, coarse=((64, 64), (48, 48)), fine=((32, 32), (24, 24)), reps=5):
    """BASELINE configs[2]: pass 1 + validation / predictor + displaced pass 2, device resident."""
    dev = torch.device("cuda", 0)
    e = Engine(0)
    fr = synth.particle_frames_torch(n_frames, H, W, dev, dtype="uint8")
    for _ in range(3):
        out = e.pairs_two_pass(fr, coarse, fine)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = e.pairs_two_pass(fr, coarse, fine); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    nr, nc = out[0].shape[1:]
    nwin = (n_frames - 1) * nr * nc
    t = float(np.median(ts))
    print(f"TWO-PASS {coarse} -> {fine} {H}x{W} uint8: {nwin} fine windows in {t:.3f} ms -> {nwin / t / 1e3:.2f} Mwin/s  (u mean {float(torch.nanmean(out[0])):.3f} v mean {float(torch.nanmean(out[1])):.3f})", flush=True)
    e.close()

def run_ens(H, W, ws, ov, n_frames, reps=5, dtype="uint8", variant=0):
    """Ensemble mode on device-resident frames: begin + add (one launch) + finish, thresholds at pyorc's defaults."""
    dev = torch.device("cuda", 0)
    e = Engine(0)
    e.set_option("kernel_variant", variant)
    fr = synth.particle_frames_torch(n_frames, H, W, dev, dtype=dtype)
    npdt = np.uint8 if dtype == "uint8" else np.float32
    torch.cuda.synchronize()
    ts = []
    for i in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nr, nc = e.ens_begin((H, W), ws, ov, npdt)
        a.record(); cm, sn = e.ens_add(fr, ws, ov, corr_min=0.2, s2n_min=3.0); b.record(); torch.cuda.synchronize()
        u, v, cnt = e.ens_finish(0.2)
        if i >= 2: ts.append(a.elapsed_time(b))
    nwin = (n_frames - 1) * nr * nc
    t = float(np.median(ts))
    print(f"ENSEMBLE variant {variant} {H}x{W} win {ws} ov {ov} {dtype}: {nwin} windows in {t:.3f} ms -> {nwin / t / 1e3:.2f} Mwin/s  (u mean {np.nanmean(u):.3f} v mean {np.nanmean(v):.3f} count mean {cnt.mean():.1f})", flush=True)
    e.close()

if __name__ == "__main__":
    if "--pad" in sys.argv:   # non power-of-two windows: padded row-per-thread kernel (4) vs shared-memory padded (1) vs direct (3)
        for ws, ov in (((26, 26), (12, 12)), ((20, 20), (10, 10)), ((30, 30), (15, 15)), ((10, 10), (5, 5)), ((12, 12), (6, 6)), ((16, 16), (8, 8)), ((14, 14), (7, 7))):
            for variant in (4, 1, 3):
                run(1080, 1920, ws, ov, 11, variant=variant)
        run_ens(1080, 1920, (26, 26), (12, 12), 21, variant=0)
        run_ens(1080, 1920, (26, 26), (12, 12), 21, variant=1)
        sys.exit(0)
    if "--ens128" in sys.argv:
        for variant in (0, 1):
            run_ens(2160, 3840, (128, 128), (64, 64), 21, variant=variant)
        sys.exit(0)
    if "--ens" in sys.argv:
        for variant in (1, 0):
            run_ens(1080, 1920, (64, 64), (32, 32), 101, variant=variant)
            run_ens(2160, 3840, (64, 64), (32, 32), 41, variant=variant)
            run_ens(1080, 1920, (32, 32), (16, 16), 41, variant=variant)
            run_ens(1080, 1920, (64, 64), (32, 32), 51, dtype="float32", variant=variant)
        sys.exit(0)
    if "--configs" in sys.argv:   # the five BASELINE.json geometries, one shard of frames each
        run(475, 371, (32, 32), (16, 16), 3)
        run(1080, 1920, (64, 64), (32, 32), 101)
        run(1080, 1920, (32, 32), (24, 24), 41)
        run_two_pass(1080, 1920, 41)
        run(2160, 3840, (64, 64), (32, 32), 41)
        run(4320, 7680, (128, 128), (64, 64), 41)
        run(4320, 7680, (128, 128), (64, 64), 6, dtype="float32")
        sys.exit(0)
    if "--single" in sys.argv:
        run(1080, 1920, (64, 64), (32, 32), 21, reps=2, variant=int(os.environ.get("B2_VARIANT", "0")))
        sys.exit(0)
    if "--variants" in sys.argv:
        run(1080, 1920, (64, 64), (32, 32), 101, variant=1)
        run(1080, 1920, (64, 64), (32, 32), 101, variant=2)
        run(1080, 1920, (64, 64), (32, 32), 101, variant=2, run_len=50)
        run(2160, 3840, (64, 64), (32, 32), 41, variant=2)
        run(1080, 1920, (32, 32), (16, 16), 41, variant=1)
        run(1080, 1920, (32, 32), (16, 16), 41, variant=2)
        run(1080, 1920, (32, 32), (24, 24), 11, variant=0)
        run(1080, 1920, (64, 64), (32, 32), 51, variant=0, dtype="float32")
        run(1080, 1920, (64, 64), (32, 32), 51, variant=1, dtype="float32")
        run(1080, 1920, (32, 32), (16, 16), 41, variant=0, dtype="float32")
        run(1080, 1920, (10, 10), (5, 5), 5, variant=0)
        run(1080, 1920, (26, 26), (12, 12), 5, variant=0)
        run(1080, 1920, (50, 50), (25, 25), 3, variant=0)
        sys.exit(0)
    run(1080, 1920, (64, 64), (32, 32), 101)
    run(1080, 1920, (64, 64), (32, 32), 101, dtype="float32")
    run(1080, 1920, (32, 32), (16, 16), 41)
    run(1080, 1920, (32, 32), (24, 24), 11)
    run(2160, 3840, (128, 128), (64, 64), 21)
    run(1080, 1920, (16, 16), (8, 8), 11)
