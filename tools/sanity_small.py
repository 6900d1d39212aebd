"""Small invocations of every fused kernel family (for compute-sanitizer memcheck / racecheck runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyorc_b200.engine import Engine
from pyorc_b200 import synth

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    e = Engine(0)
    dev = torch.device("cuda", 0)
    def frames(n, h, w, dtype="uint8"):
        return synth.particle_frames_torch(n, h, w, dev, dtype=dtype)
    if which in ("all", "rows64"):
        out = e.pairs(frames(4, 200, 304), (64, 64), (32, 32)); print("rows64", float(torch.nanmean(out[0])))
    if which in ("all", "rows32"):
        out = e.pairs(frames(4, 120, 176), (32, 32), (24, 24)); print("rows32 unaligned", float(torch.nanmean(out[0])))
    if which in ("all", "rows128"):
        out = e.pairs(frames(4, 300, 432), (128, 128), (64, 64)); print("rows128", float(torch.nanmean(out[0])))
    if which in ("all", "rows128", "rows128ens"):
        fr = frames(4, 300, 432)            # device tensors are used in place: the TMA kernels need a 16-byte pitch
        e.ens_begin((300, 432), (128, 128), (64, 64), np.uint8); e.ens_add(fr, (128, 128), (64, 64), corr_min=0.2, s2n_min=3.0)
        u, v, cnt = e.ens_finish(0.2); print("rows128 ensemble", float(np.nanmean(u)))
    if which in ("all", "rows128", "rows128f32"):
        out = e.pairs(frames(3, 300, 432, "float32"), (128, 128), (64, 64)); print("rows128 f32", float(torch.nanmean(out[0])))
        fr = frames(3, 300, 432, "float32")
        e.ens_begin((300, 432), (128, 128), (64, 64), np.float32); e.ens_add(fr, (128, 128), (64, 64), corr_min=0.2, s2n_min=3.0)
        u, v, cnt = e.ens_finish(0.2); print("rows128 f32 ensemble", float(np.nanmean(u)))
    if which in ("all", "rows128", "pad128"):
        out = e.pairs(frames(4, 160, 224), (50, 50), (25, 25)); print("pad128 50x50", float(torch.nanmean(out[0])))
        out = e.pairs(frames(3, 130, 176), (36, 20), (18, 10)); print("pad128 36x20", float(torch.nanmean(out[0])))
        fr = frames(4, 160, 224)
        e.ens_begin((160, 224), (50, 50), (25, 25), np.uint8); e.ens_add(fr, (50, 50), (25, 25), corr_min=0.2, s2n_min=3.0)
        u, v, cnt = e.ens_finish(0.2); print("pad128 ensemble", float(np.nanmean(u)))
    if which in ("all", "push"):
        res = e.pairs(frames(4, 200, 304), (64, 64), (32, 32))
        buf = torch.zeros((4, 5, res[0].shape[1], res[0].shape[2]), device=dev)
        from pyorc_b200.parallel import _result_block
        e.peer_push(_result_block(res), [buf.data_ptr()], 5, 2); torch.cuda.synchronize(); print("peer push", float(torch.nanmean(buf[0, 2:])))
    if which in ("all", "generic"):
        e.set_option("kernel_variant", 1.0)
        out = e.pairs(frames(3, 130, 170), (50, 50), (25, 25)); print("shared-memory kernel, padded 50x50", float(torch.nanmean(out[0])))
        e.set_option("kernel_variant", 0.0)
    if which in ("all", "pad"):
        out = e.pairs(frames(3, 120, 176), (26, 26), (12, 12)); print("pad26", float(torch.nanmean(out[0])))
    if which in ("all", "twopass"):
        out = e.pairs_two_pass(frames(4, 200, 288)); print("two-pass", float(torch.nanmean(out[0])))
    if which in ("all", "ens"):
        fr = frames(5, 200, 304)
        e.ens_begin((200, 304), (64, 64), (32, 32), np.uint8); e.ens_add(fr, (64, 64), (32, 32), corr_min=0.2, s2n_min=3.0)
        u, v, cnt = e.ens_finish(0.2); print("ens", float(np.nanmean(u)))
    if which in ("all", "rows64smem"):
        e.set_option("tmem", 0.0)
        out = e.pairs(frames(4, 200, 304), (64, 64), (32, 32)); print("rows64 (parked spectra in shared memory)", float(torch.nanmean(out[0])))
        e.set_option("tmem", 1.0)
    if which in ("all", "parts"):
        e.set_option("unit_parts", 12.0)
        out = e.pairs(frames(6, 200, 304), (64, 64), (32, 32)); print("rows64 tmem, forced even partition", float(torch.nanmean(out[0])))
        e.set_option("unit_parts", 0.0)
    if which in ("all", "big"):
        out = e.pairs(frames(3, 200, 260), (100, 100), (50, 50)); print("direct 100x100", float(torch.nanmean(out[0])))
    if which in ("all", "deform"):
        out = e.pairs_two_pass(frames(4, 200, 288), mode="deform"); print("two-pass deform", float(torch.nanmean(out[0])))
    if which in ("all", "f32"):
        out = e.pairs(frames(3, 200, 304, "float32"), (64, 64), (32, 32)); print("rows64 f32", float(torch.nanmean(out[0])))
    torch.cuda.synchronize()
    e.close()
