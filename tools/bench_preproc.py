"""Device-resident timing of the pre-processing kernels against the HBM copy peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyorc_b200 import preprocess as G, synth

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

if __name__ == "__main__":
    peak = 6568.0
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p): peak = float(json.load(open(p))["hbm_gbs"])
    n, H, W = 101, 1080, 1920
    fr = synth.particle_frames_torch(n, H, W, torch.device("cuda", 0), dtype="uint8")
    ff = fr.float()
    px = n * H * W
    rows = []
    # algorithmic bytes: every input byte read once, every output byte written once (normalize: input read twice + mean)
    rows.append(("normalize u8->u8", timeit(lambda: G.normalize(fr, 15)), px * 1 * 2 + px * 1 + 7 * H * W * 1 + H * W * 4 * 3))
    rows.append(("time_diff u8->f32", timeit(lambda: G.time_diff(fr, 2.0)), px * 1 + (px - H * W) * 4))
    rows.append(("time_diff f32->f32", timeit(lambda: G.time_diff(ff, 2.0)), px * 4 + (px - H * W) * 4))
    rows.append(("minmax f32", timeit(lambda: G.minmax(ff, 10, 200)), px * 8))
    rows.append(("smooth k=3 u8->f32", timeit(lambda: G.smooth(fr, 1)), px * 1 + px * 4))
    rows.append(("edge_detect 3/5 f32->f32", timeit(lambda: G.edge_detect(ff, 1, 2)), px * 8))
    # orthoprojection with the Ngwerere camera configuration's index maps (1080p -> 475x371), tests/golden/ngwerere_maps.npz
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mp = os.path.join(root, "tests", "golden", "ngwerere_maps.npz")
    if os.path.exists(mp):
        from pyorc_b200 import project as PJ
        d = np.load(mp)
        und = lambda a: np.cumsum(a.astype(np.int64))
        maps = dict(idx_img=und(d["d_idx_img"]), idx_ortho=np.unpackbits(d["idx_ortho_bits"])[: int(d["n_ortho"])].astype(bool),
                    src_idx=und(d["d_src_idx"]), uidx=und(d["d_uidx"]), norm_idx=und(d["d_norm_idx"]))
        oh, ow = (int(v) for v in d["ortho_shape"])
        proj = PJ.OrthoProjector((H, W), (oh, ow), **maps)
        grp = np.zeros(oh * ow, bool); grp[maps["uidx"]] = True
        nn_used = maps["idx_img"][~grp[np.flatnonzero(maps["idx_ortho"])]]
        touched = np.unique(np.concatenate([nn_used, maps["src_idx"]])).size
        rows.append((f"project u8 1080p->{oh}x{ow}", timeit(lambda: proj(fr)), n * (touched + oh * ow)))
        rows.append((f"project f32 1080p->{oh}x{ow}", timeit(lambda: proj(ff)), n * (touched + oh * ow) * 4))
        if "--cpu" in sys.argv:
            import time
            from oracle import project_oracle as PO
            host = fr[:4].cpu().numpy()
            t0 = time.perf_counter(); PO.project_stack(host, np.arange(ow), np.arange(oh), **maps); dt = (time.perf_counter() - t0) / 4
            print(f"cpu oracle project: {dt * 1e3:.1f} ms / frame (1 core, numpy restatement of project.py:123-157)", flush=True)
    for name, ms, byts in rows:
        gbs = byts / ms / 1e6
        print(f"{name:28s} {ms:8.3f} ms  {gbs:8.1f} GB/s algorithmic  = {gbs / peak:.3f} of measured HBM copy peak {peak:.0f} GB/s", flush=True)
