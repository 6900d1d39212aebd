#!/usr/bin/env python
"""bench.py - interrogation windows / s of the fused LSPIV engine (BASELINE.json metric).

Workload (BASELINE.json configs[1]): synthetic 1080p, 100 frame pairs, 64x64 windows, 50 % overlap, single pass,
uint8 frames.  A "step" is one pass of the hot path over that batch (188 800 windows) on every GPU (weak scaling:
each rank owns its own 100-pair shard, frame pairs shard with no data-path collective; for N > 1 the 16 B/window
results reach every rank through P2P stores fused into the kernel epilogue).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port, all host cores)

Prints ONE JSON line (rank 0).
  value      windows/s with frames resident in HBM (CUDA events around K steps, max over ranks)
  e2e        the same metric through the reference-facing binding `pyorc_b200.velocimetry.get_b2piv` on ORDINARY (pageable)
             numpy frames - what pyorc hands over (`frame_chunk.values`, ffpiv.py:223,451): chunk loop, H2D, kernel, D2H,
             unit conversion and Dataset packaging inside the timed region; N > 1: plus the gather of every rank's fields
  roofline   algorithmic HBM bytes of the dominant kernel against the measured copy bandwidth (MEASURED_PEAKS.json)
  fp32       algorithmic FFT flops against the fp32 FMA peak MEASURED in this run (the bound that actually binds)
  cpu_baseline  the oracle port on all host cores (process pool), a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "interrogation windows/sec at 64x64, 50% overlap"
H, W = 1080, 1920
WS, OV = (64, 64), (32, 32)
N_PAIRS = 100
NR, NC = (H - WS[0]) // (WS[0] - OV[0]) + 1, (W - WS[1]) // (WS[1] - OV[1]) + 1     # 32 x 59
WORKLOAD = "synthetic 1080p, 100 frame pairs, 64x64 windows 50% overlap, single-pass, 1xB200 (per GPU)"
# compulsory HBM traffic and FFT flops per window
B_ALG = 2 * (WS[0] - OV[0]) * (WS[1] - OV[1]) * 1 + 16                      # 2064 B (uint8)
F_ALG = 3 * 2.5 * (WS[0] * WS[1]) * np.log2(WS[0] * WS[1]) + 6 * WS[0] * (WS[1] // 2 + 1)  # 381 312 flop


def config_dict(world):
    """The `config` object - identical for both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "frame": [H, W], "pairs_per_gpu": N_PAIRS, "window": list(WS), "overlap": list(OV),
            "input_dtype": "uint8", "windows_per_step": N_PAIRS * NR * NC * world,
            "l2": f"inputs {(N_PAIRS + 1) * H * W / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)",
            "parallelism": f"frame-pair shard x{world}"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: NVML polled in-process every ~2 ms (the timed
    region of the default run lasts tens of milliseconds, too short for `nvidia-smi -lms`), nvidia-smi as fall-back."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, uuid=None):
        self.index, self.uuid = index, uuid
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max, self.how = None, None
        self._stop = threading.Event()
        self._thread, self._nvml, self._h = None, None, None

    def _open_nvml(self):
        import pynvml

        pynvml.nvmlInit()
        h = None
        if self.uuid is not None:
            for u in (f"GPU-{self.uuid}", str(self.uuid)):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(u)
                    break
                except Exception:
                    h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self._nvml, self._h = pynvml, h
        self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

    def _poll_nvml(self):
        n, h = self._nvml, self._h
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
                self.power.append(n.nvmlDeviceGetPowerUsage(h) / 1e3)
            except Exception:
                pass
            time.sleep(0.002)

    def _poll_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
                r = [c.strip() for c in out.split(",")]
                self.sm.append(float(r[0]))
                self.sm_max = float(r[1])
                self.power.append(float(r[2]))
                for nme, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                time.sleep(0.05)

    def start(self):
        try:
            self._open_nvml()
            self.how = "nvml, 2 ms period, during the timed regions"
            target = self._poll_nvml
        except Exception:
            self.how = "nvidia-smi, back to back, during the timed regions"
            target = self._poll_smi
        self._thread = threading.Thread(target=target, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=15)
        sm = self.sm
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm), "sm_mhz_min": float(min(sm)) if sm else None,
                "power_w_max": float(max(self.power)) if self.power else None, "how": self.how}


# ---- the reference's CPU path ----------------------------------------------------------------------------------------
def cpu_reference(frames_np, pool):
    """The reference's CPU path (oracle port: float64 pocketfft, all passes over memory pyorc makes) on `frames_np`."""
    from oracle import ffpiv_oracle as O

    t0 = time.perf_counter()
    u, v, c, s = O.cpu_reference_pairs(frames_np, WS, OV, pool=pool)
    dt = time.perf_counter() - t0
    return (u, v, c, s), dt


def cpu_sample_note(pairs, pool):
    return (f"{pairs} of the workload's 100 frame pairs per step ({pairs * NR * NC} windows); restated ffpiv CPU path (upstream "
            f"ffpiv / rocket-fft are not installable offline): float64 pocketfft, gather -> normalise -> rfft2.conj.irfft2 -> f32 planes -> "
            f"nanmax / nanmean -> peak fit, one frame pair per job on a pool of {pool.n_workers} worker processes")


def run_reference(args, rank, world):
    """--impl reference: the CPU path on the box's host cores (rank 0 only), every step a bounded sample of the workload sized so
    that the whole run takes about a minute."""
    if rank != 0:
        return
    from oracle import ffpiv_oracle as O
    from pyorc_b200 import synth

    cores = os.cpu_count() or 1
    pool = O.make_pool(cores)
    # warm-up steps also size the sample: K steps of `pairs` frame pairs should take about 45 s
    probe = min(N_PAIRS, max(cores, 8))
    frames = synth.particle_frames(probe + 1, H, W, dtype=np.uint8)
    rate = None
    for _ in range(max(1, min(args.warmup, 2))):
        _, dt = cpu_reference(frames, pool)
        rate = probe / dt
    pairs = args.cpu_pairs if args.cpu_pairs > 0 else int(min(N_PAIRS, max(cores, 4, rate * 45.0 / max(args.steps, 1))))
    if pairs >= cores:
        pairs = min(N_PAIRS, (pairs // cores) * cores) or pairs      # whole rounds of the pool
    if pairs + 1 > frames.shape[0]:
        frames = synth.particle_frames(pairs + 1, H, W, dtype=np.uint8)   # same seed: the same frames, more of them
    nwin = pairs * NR * NC
    ts = []
    for _ in range(args.steps):
        _, dt = cpu_reference(frames[: pairs + 1], pool)
        ts.append(dt)
    total = float(np.sum(ts))
    val = nwin * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world),
        "cpu_baseline": {"value": val, "unit": "windows/s", "cores": pool.n_workers, "kind": "port", "sample": cpu_sample_note(pairs, pool),
                         "host_cores": cores},
        "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    pool.shutdown()


# ---- the other BASELINE.json configurations ------------------------------------------------------------------------------
def _timed(torch, dev, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    ts = []
    out = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), out


def _alg(ws, ov):
    b_alg = 2 * (ws[0] - ov[0]) * (ws[1] - ov[1]) + 16
    f_alg = 3 * 2.5 * ws[0] * ws[1] * np.log2(ws[0] * ws[1]) + 6 * ws[0] * (ws[1] // 2 + 1)
    return b_alg, f_alg


def other_configs(eng, dev, fp32_peak):
    """windows/s of BASELINE.json configs[0] and [2] (their geometry, a shard of frame pairs each, frames resident in HBM, CUDA
    events, median of 5 after 2 warm-ups) with the algorithmic HBM bytes / fp32 flops of SURVEY.md 8(d).  configs[3] / [4]: see
    sharded_configs."""
    import torch

    from pyorc_b200 import synth

    peak, _ = measured_peaks()
    rows = []
    c75, c50 = ((64, 64), (48, 48)), ((64, 64), (32, 32))
    cases = [("configs[0] geometry: 475x371, 32x32 / 50 %", 475, 371, (32, 32), (16, 16), 3, None),
             ("configs[2] single pass: 1080p, 32x32 / 75 %", 1080, 1920, (32, 32), (24, 24), 21, None),
             ("configs[2] two-pass, discrete window offset (64x64 / 75 % -> 32x32 / 75 %; round 1's row)", 1080, 1920, (32, 32), (24, 24), 21, (c75, "offset")),
             ("configs[2] two-pass, discrete window offset (64x64 / 50 % -> 32x32 / 75 %)", 1080, 1920, (32, 32), (24, 24), 21, (c50, "offset")),
             ("configs[2] two-pass DEFORM: bilinear window deformation (64x64 / 50 % -> 32x32 / 75 %)", 1080, 1920, (32, 32), (24, 24), 21, (c50, "deform")),
             ("f-2: 1080p, 50x50 / 50 % (windows of 34 .. 64 px: padded mode of the 128-plane polyphase kernel)", 1080, 1920, (50, 50), (25, 25), 21, None),
             ("f-2: 1080p, 50x50 / 50 %, FLOAT32 frames (padded mode of the 128-plane polyphase kernel behind pyorc's float32 filters)", 1080, 1920, (50, 50), (25, 25), 21, "float32"),
             ("f-2: 1080p, 26x26 / overlap 12, uint8 (pyorc's window_size=25 -> 26; padded mode of the row-per-thread kernel)", 1080, 1920, (26, 26), (12, 12), 21, None),
             ("f-2: 1080p, 26x26 / overlap 12, FLOAT32 frames (pyorc's own recipe: normalize -> edge_detect -> minmax -> get_piv(window_size=25))", 1080, 1920, (26, 26), (12, 12), 21, "float32")]
    for name, h, w, ws, ov, n, two_pass in cases:
        try:
            in_dtype = "uint8"
            if two_pass == "float32":
                in_dtype, two_pass = "float32", None
            fr = synth.particle_frames_torch(n, h, w, dev, dtype=in_dtype)
            if two_pass:
                ms, out = _timed(torch, dev, lambda: eng.pairs_two_pass(fr, two_pass[0], (ws, ov), mode=two_pass[1]))
            else:
                ms, out = _timed(torch, dev, lambda: eng.pairs(fr, ws, ov))
            nwin = int(out[0].numel())
            b_alg, f_alg = _alg(ws, ov)
            if in_dtype == "float32":
                b_alg = 2 * (ws[0] - ov[0]) * (ws[1] - ov[1]) * 4 + 16
            rows.append({"config": name, "pairs": n - 1, "windows": nwin, "ms": ms, "windows_per_s": nwin / (ms * 1e-3),
                         "hbm_frac": b_alg * nwin / (ms * 1e-3) / 1e9 / peak,
                         "fp32_tflops": f_alg * nwin / (ms * 1e-3) / 1e12,
                         "fp32_frac_of_measured": f_alg * nwin / (ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None})
            if h == 1080:   # accuracy against the imposed synthetic field (synth.displacement_field), px
                y0 = np.arange(out[0].shape[1]) * (ws[0] - ov[0]) + ws[0] / 2.0
                x0 = np.arange(out[0].shape[2]) * (ws[1] - ov[1]) + ws[1] / 2.0
                yc, xc = np.meshgrid(y0, x0, indexing="ij")
                tx, ty = synth.displacement_field(h, w, yc, xc)
                uu, vv = out[0].cpu().numpy(), out[1].cpu().numpy()
                rows[-1]["rmse_vs_imposed_field_px"] = float(np.sqrt(np.nanmean((uu - tx[None]) ** 2 + (vv - ty[None]) ** 2)))
            del fr, out
        except Exception as exc:   # a shard that does not fit must not take the headline down with it
            rows.append({"config": name, "error": str(exc)[:200]})
    eng.plan((H, W), WS, OV, np.uint8)
    return rows


def sharded_configs(eng, dev, rank, world, fp32_peak):
    """BASELINE.json configs[3] (4K, 2000 pairs, 64x64 / 50 %, 4 GPUs) and configs[4] (8K, 5000 pairs, 128x128 / 50 %, 8 GPUs).

    Frame pairs shard over the ranks; every rank renders ITS frames on the device from (seed, frame index) (the stacks - 17 GB
    and 166 GB - never exist on the host), holds them in HBM and walks through them in chunks with the reference's 1-frame halo
    (pyorc/velocimetry/ffpiv.py:140, :399-440), one kernel launch per chunk; for N > 1 every chunk's 16 B / window are
    pushed into every rank's gather buffer (P2P stores on a side stream, parallel.PeerGather).  The per-rank share is BASELINE's own
    at the N it names - 500 pairs of 4K, 625 pairs of 8K - so the row at N = 4 (configs[3]) and at N = 8 (configs[4]) IS that
    configuration in full, and the rows at the other N are its weak-scaling series (N = 1: the single-GPU reference point)."""
    import torch
    import torch.distributed as dist

    from pyorc_b200 import parallel, synth

    peak, _ = measured_peaks()
    rows = []
    cases = [("configs[3]: synthetic 4K, 64x64 / 50 %, frame-pair shard", 2160, 3840, (64, 64), (32, 32), 500, 125, 4),
             ("configs[4]: synthetic 8K, 128x128 / 50 %, frame-pair shard", 4320, 7680, (128, 128), (64, 64), 625, 125, 8)]
    for name, h, w, ws, ov, pairs_rank, chunk, named_n in cases:
        row = {"config": name, "n_gpus": world, "pairs_per_gpu": pairs_rank, "pairs_total": pairs_rank * world, "chunk_pairs": chunk,
               "is_baseline_config_in_full": world == named_n}
        try:
            t0 = time.perf_counter()
            fr = synth.particle_frames_torch(pairs_rank + 1, h, w, dev, dtype="uint8", first_frame=rank * pairs_rank)
            torch.cuda.synchronize(dev)
            row["render_s"] = time.perf_counter() - t0
            nr, nc = eng.plan((h, w), ws, ov, np.uint8)
            total = pairs_rank * world
            table = parallel.shard_pairs(total, world)
            peer = parallel.PeerGather(eng, total, table) if world > 1 else None
            bounds = [(a, min(a + chunk, pairs_rank)) for a in range(0, pairs_rank, chunk)]
            local = torch.empty((4, pairs_rank, nr, nc), dtype=torch.float32, device=dev)

            def job():
                if peer is not None:
                    peer.begin()
                for a, b in bounds:
                    res = eng.pairs(fr[a : b + 1], ws, ov)
                    if peer is not None:      # this chunk's pairs start at (rank offset + a) of the gathered time axis
                        peer.push(res, int(table[rank, 0]) + a)
                    else:
                        for k in range(4):
                            local[k, a:b] = res[k]
                if peer is not None:
                    return peer.end()
                return local, None

            def sync():
                torch.cuda.synchronize(dev)
                if world > 1:
                    peer.drain()
                    dist.barrier()
                    torch.cuda.synchronize(dev)

            for _ in range(2):
                job()
            sync()
            reps = 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                out, ready = job()
            if ready is not None:
                torch.cuda.current_stream(dev).wait_event(ready)
            e1.record()
            sync()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            nwin = total * nr * nc
            b_alg, f_alg = _alg(ws, ov)
            u_mean = float(torch.nanmean(out[0]))
            row.update({"windows": nwin, "ms": ms, "windows_per_s": nwin / (ms * 1e-3), "frames_gb_per_gpu": fr.numel() / 1e9,
                        "hbm_frac_per_gpu": b_alg * nwin / world / (ms * 1e-3) / 1e9 / peak,
                        "fp32_tflops_per_gpu": f_alg * nwin / world / (ms * 1e-3) / 1e12,
                        "fp32_frac_of_measured": f_alg * nwin / world / (ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None,
                        "u_mean_px": u_mean, "gather": "fused P2P stores, slot ring" if world > 1 else None})
            if peer is not None:
                peer.close()
            del fr, out, local, peer
            torch.cuda.empty_cache()
        except Exception as exc:
            row["error"] = str(exc)[:300]
        rows.append(row)
    eng.plan((H, W), WS, OV, np.uint8)
    return rows


def ensemble_check(eng, dev, rank, world):
    """Ensemble mode across GPUs (pyorc/velocimetry/ffpiv.py:345-376): the 100 * N frame pairs of ONE continuous
    synthetic sequence, sharded over the ranks (parallel.ensemble_sharded: reduce-scatter of the plane sums over the window
    axis, per-slice peak fit, all-gather), against the same pairs on rank 0 alone with the ranks' ranges as its chunks."""
    import torch
    import torch.distributed as dist

    from pyorc_b200 import parallel, synth

    total = N_PAIRS * world
    table = parallel.shard_pairs(total, world)
    a, b = int(table[rank, 0]), int(table[rank, 1])
    fr = synth.particle_frames_torch(b - a + 1, H, W, dev, dtype="uint8", first_frame=a)
    kw = dict(corr_min=0.2, s2n_min=3.0, count_min=0.2)
    for _ in range(2):
        res = parallel.ensemble_sharded(eng, fr, WS, OV, total, table, n_chunks_total=world, **kw)
    torch.cuda.synchronize(dev)
    dist.barrier()
    t0 = time.perf_counter()
    res = parallel.ensemble_sharded(eng, fr, WS, OV, total, table, n_chunks_total=world, **kw)
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    u, v, cm, sn, cnt = res
    out = None
    # the single-GPU reference must see the SAME frames: the device renderer scatters the particles with atomics (index_add_),
    # so two renderings of one frame can differ in a last bit - every rank's stack, halo frame included, is gathered instead
    stacks = torch.empty((world,) + tuple(fr.shape), dtype=fr.dtype, device=dev)
    dist.all_gather_into_tensor(stacks.view(world * fr.shape[0], H, W), fr)
    if rank == 0:
        eng.ens_begin((H, W), WS, OV, np.uint8, device_ordered=True)
        cms, sns = [], []
        for r_ in range(world):
            c_, s_ = eng.ens_add(stacks[r_], WS, OV, corr_min=kw["corr_min"], s2n_min=kw["s2n_min"])
            cms.append(c_.cpu().numpy())
            sns.append(s_.cpu().numpy())
        u1, v1, cnt1 = eng.ens_finish(kw["count_min"] * world)
        cm1, sn1 = parallel.aggregate_ensemble(np.concatenate(cms), np.concatenate(sns), cnt1, kw["count_min"] * world, NR, NC)
        ok = np.isfinite(u1)
        out = {"pairs_total": total, "windows": total * NR * NC, "ms": 1e3 * float(dt.item()), "windows_per_s": total * NR * NC / float(dt.item()),
               "counts_bit_equal": bool(np.array_equal(cnt, cnt1)), "nan_mask_equal": bool(np.array_equal(np.isnan(u.reshape(-1)), np.isnan(u1))),
               "max_abs_du_px": float(np.abs(u.reshape(-1)[ok] - u1[ok]).max()), "max_abs_dv_px": float(np.abs(v.reshape(-1)[ok] - v1[ok]).max()),
               "corr_s2n_means_equal": bool(np.array_equal(cm, cm1, equal_nan=True) and np.array_equal(sn, sn1, equal_nan=True)),
               "valid_windows": int(ok.sum()), "timed": "ens_begin + accumulate (1 launch) + reduce-scatter + peak fit of the slice + all-gather + host means, wall clock"}
        out["equals_single_gpu"] = bool(out["counts_bit_equal"] and out["nan_mask_equal"] and out["max_abs_du_px"] <= 2e-3 and out["max_abs_dv_px"] <= 2e-3
                                        and out["corr_s2n_means_equal"])
    del fr, stacks
    torch.cuda.empty_cache()
    eng.plan((H, W), WS, OV, np.uint8)
    return out


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner / NCCL_DEBUG output) are sent to
    stderr by pointing fd 1 at fd 2 for the whole run; emit() writes the line to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nccl-gather", action="store_true", help="N > 1: gather results with NCCL instead of our own P2P stores")
    ap.add_argument("--fused-gather", action="store_true", help="N > 1: P2P stores from the PIV kernel's epilogue instead of the push kernel on the side stream")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the runs of BASELINE.json configs[0], [2], [3], [4] and the N > 1 ensemble check")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="frame pairs of the workload the CPU baseline times per step (0: sized from the host's speed)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    # the CPU baseline's worker processes are forked BEFORE this process touches CUDA (they only ever run numpy)
    pool = None
    if world == 1:
        from oracle import ffpiv_oracle as O

        pool = O.make_pool(os.cpu_count() or 1)

    import torch
    import torch.distributed as dist

    from pyorc_b200 import _xr, parallel, synth, velocimetry
    from pyorc_b200.engine import get_engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = get_engine(local_rank)          # the engine get_b2piv(device=local_rank) uses
    host_cores = os.cpu_count() or 1
    try:
        host_cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    stage_threads = max(2, min(8, host_cores // world))
    eng.set_option("stage_threads", stage_threads)    # the ranks share one host: split its cores between their staging pools
    n_frames = N_PAIRS + 1
    frames = synth.particle_frames_torch(n_frames, H, W, dev, dtype="uint8", seed=synth.SEED + rank)
    nr, nc = eng.plan((H, W), WS, OV, np.uint8)
    assert (nr, nc) == (NR, NC)
    nwin_rank = N_PAIRS * nr * nc
    table = parallel.shard_pairs(N_PAIRS * world, world)

    # N > 1: the gather of the 16 B / window results is P2P stores of our own kernels into every rank's symmetric-memory slot
    # ring - by a push kernel on the consumer stream, overlapped with the next step (default), or from the PIV kernel's epilogue
    # (--fused-gather: its end-of-kernel wait for NVLink costs 2 .. 4 % of the step) - and a completion barrier on the consumer
    # stream; NCCL all-gather when peer memory is missing
    peer = None
    gather_how = None
    if world > 1 and not args.nccl_gather:
        try:
            peer = parallel.PeerGather(eng, N_PAIRS * world, table, mode="fused" if args.fused_gather else "push")
            gather_how = ("P2P stores from the PIV kernel's epilogue" if args.fused_gather else
                          "P2P push kernel (16-byte stores to every rank) on the consumer stream, overlapped with the next step's PIV kernel")
            gather_how += "; ring of 3 symmetric-memory slots, completion barrier on the consumer stream"
        except Exception as exc:   # no peer access / symmetric memory: keep the collective
            peer = None
            gather_how = f"nccl all_gather_into_tensor per field (symmetric memory unavailable: {str(exc)[:80]})"
    elif world > 1:
        gather_how = "nccl all_gather_into_tensor per field"

    def step_device():
        if peer is not None:
            peer.begin()
            res = eng.pairs(frames, WS, OV)
            return peer.end(res if peer.mode == "push" else None)
        res = eng.pairs(frames, WS, OV)
        if world > 1:
            return parallel.gather_fields(res, N_PAIRS * world, table), None   # four fields, gathered in place
        return res, None

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            if peer is not None:
                peer.drain()
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident throughput (value) ---------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sync_all()
    launches0 = eng.launch_count
    try:
        uuid = torch.cuda.get_device_properties(dev).uuid
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        out, ready = step_device()
    if ready is not None:
        torch.cuda.current_stream(dev).wait_event(ready)      # the last step's results have landed on this rank from everywhere
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    if peer is not None:
        sync_all()
        gathered, ready = step_device()
        ready.synchronize()
        sync_all()
        # every rank must hold every rank's results: compare with an NCCL gather of the same step
        eng.set_peer_outputs(None, 1, 0)
        ref_g = parallel.gather_fields(eng.pairs(frames, WS, OV), N_PAIRS * world, table)
        same = bool(torch.equal(torch.nan_to_num(gathered), torch.nan_to_num(ref_g)))
        gather_how += f"; equals NCCL gather: {same}"
        del ref_g
        sync_all()
    # kernel-only average launch duration (same stream, no gather) for the roofline
    if peer is not None:
        eng.set_peer_outputs(None, 1, 0)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    k0.record()
    for _ in range(args.steps):
        eng.pairs(frames, WS, OV)
    k1.record()
    torch.cuda.synchronize(dev)
    ms_kernel = k0.elapsed_time(k1) / args.steps
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = nwin_rank * world / (ms_step * 1e-3)

    # ---- end to end through the reference-facing binding --------------------------------------------------------------------
    # get_b2piv on an ORDINARY numpy-backed DataArray (pyorc hands `frame_chunk.values` over, ffpiv.py:223,451): memory model +
    # chunk list, staging of the pageable frames, chunked H2D overlapped with the kernel, D2H of the four fields, px -> m/s,
    # Dataset.  N > 1: every rank does that for its shard, then the fields are gathered to every rank (pinned host -> device ->
    # NCCL all-gather per field -> host), i.e. each rank ends with the whole time axis like a single get_b2piv call would.
    host_np = frames.cpu().numpy()                             # pageable
    tcoord = np.arange(n_frames) / 30.0
    da = _xr.DataArray(host_np, ("time", "y", "x"), {"time": tcoord})
    yx = (np.arange(nr), np.arange(nc))
    dt_pairs = np.full(N_PAIRS, 1 / 30.0)
    gbuf = torch.empty((4, N_PAIRS, nr, nc), dtype=torch.float32, device=dev) if world > 1 else None
    ghost = torch.empty((4, N_PAIRS * world, nr, nc), dtype=torch.float32, pin_memory=True) if world > 1 else None

    def step_e2e():
        ds = velocimetry.get_b2piv(da, yx[0], yx[1], dt_pairs, WS, OV, WS, 0.01, 0.01, device=local_rank)
        if world == 1:
            return ds
        for k, name in enumerate(("v_x", "v_y", "corr", "s2n")):
            gbuf[k].copy_(torch.from_numpy(ds[name].values), non_blocking=True)
        g = parallel.gather_fields(gbuf, N_PAIRS * world, table)
        ghost.copy_(g, non_blocking=True)
        torch.cuda.synchronize(dev)
        return ds

    for _ in range(3):
        ds = step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ds = step_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = nwin_rank * world * args.steps / e2e_s
    h2d = int(host_np.nbytes) + (int(gbuf.numel()) * 4 if world > 1 else 0)
    d2h = int(4 * ds["v_x"].values.nbytes) + (int(ghost.numel()) * 4 if world > 1 else 0)
    hu = ds["v_x"].values * (1 / 30.0) / 0.01                 # back to px / frame for the parity figure below
    hv = ds["v_y"].values * (1 / 30.0) / 0.01

    # the same frames from PAGE-LOCKED memory through Engine.pairs (round 1's e2e: no chunk loop, no Dataset), for reference
    host_pin = eng.pinned_empty((n_frames, H, W), np.uint8)
    host_pin[...] = host_np
    for _ in range(2):
        eng.pairs(host_pin, WS, OV)
    sync_all()
    reps = min(args.steps, 20)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.pairs(host_pin, WS, OV)
    tp = torch.tensor([(time.perf_counter() - t0) / reps], device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    pinned = {"value": nwin_rank * world / float(tp.item()), "ms_per_step": 1e3 * float(tp.item()), "api": "Engine.pairs(page-locked numpy), no gather"}

    # the PCIe floor under e2e: the same bytes, pinned host -> device, nothing else - on ALL ranks at the same time (the host's
    # memory / PCIe fabric is shared: 8 concurrent copies do not each get the rate of one)
    hp = torch.empty(frames.numel(), dtype=torch.uint8, pin_memory=True)
    dd = torch.empty_like(frames).view(-1)
    dd.copy_(hp, non_blocking=True)
    sync_all()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(5):
        dd.copy_(hp, non_blocking=True)
    c1.record()
    torch.cuda.synchronize(dev)
    h2d_ms = c0.elapsed_time(c1) / 5
    tt = torch.tensor([h2d_ms, -h2d_ms], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    slow, fast = float(tt[0].item()), -float(tt[1].item())
    pcie = {"concurrent_ranks": world, "h2d_only_ms_slowest_rank": slow, "h2d_only_ms_fastest_rank": fast,
            "h2d_gbs_per_gpu_slowest": hp.numel() / slow / 1e6, "h2d_gbs_per_gpu_fastest": hp.numel() / fast / 1e6,
            "e2e_over_floor": (1e3 * e2e_s / args.steps) / slow}
    del hp, dd

    clocks = sampler.stop() if rank == 0 else None

    # ---- fp32 FMA peak of this GPU, measured now -------------------------------------------------------------------------------
    fp32_peak = None
    try:
        fp32_peak = eng.fp32_peak()
    except Exception:
        pass

    # ---- the other BASELINE.json configurations (reported beside the headline, not part of it: the metric is quoted on
    # configs[1]); N > 1: configs[3] / [4] sharded over the ranks and the ensemble mode across GPUs -----------------------------
    other, sharded, ens = None, None, None
    if not args.no_other_configs:
        if world == 1:
            other = other_configs(eng, dev, fp32_peak)
        sharded = sharded_configs(eng, dev, rank, world, fp32_peak)
        if world > 1:
            try:
                ens = ensemble_check(eng, dev, rank, world)
            except Exception as exc:
                ens = {"error": str(exc)[:300]}

    if rank != 0:
        if peer is not None:
            peer.close()
        eng.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel -----------------------------------------------------------------
    peak, peak_src = measured_peaks()
    achieved = B_ALG * nwin_rank / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "piv_rows_tm_kernel<RCfg<64>, 6, aligned> (row-per-thread register FFTs in packed fp32, TMA tiles, parked spectra in Tensor Memory)", "ms_per_launch": ms_kernel, "alg_bytes_per_window": B_ALG,
                "windows_per_launch": nwin_rank, "peak_source": peak_src,
                "note": "fused kernel is fp32-issue/shared-memory bound by construction (SURVEY.md 8d); see fp32"}
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    fp32 = {"achieved": F_ALG * nwin_rank / (ms_kernel * 1e-3) / 1e12, "unit": "TFLOP/s", "alg_flop_per_window": F_ALG,
            "peak_measured": fp32_peak, "peak_nominal": 148 * 128 * 2 * 1.965e9 / 1e12,
            "peak_how": "b2piv_fp32_peak: 16 independent FFMA chains per thread, 8 x 256 threads per SM, best of 4, CUDA events, this run"}
    fp32["frac"] = fp32["achieved"] / fp32_peak if fp32_peak else None
    fp32["frac_of_nominal"] = fp32["achieved"] / fp32["peak_nominal"]

    # ---- CPU baseline (oracle port, all host cores) on a bounded sample + parity of the e2e result on that sample --------------
    cpu = None
    rmse = None
    if world == 1:
        cores = os.cpu_count() or 1
        cpu_pairs = args.cpu_pairs if args.cpu_pairs > 0 else min(N_PAIRS, max(cores, 16))
        if cpu_pairs >= cores:
            cpu_pairs = min(N_PAIRS, (cpu_pairs // cores) * cores) or cpu_pairs
        cpu_reference(host_np[:5], pool)
        (u, v, c, s), dtc = cpu_reference(host_np[: cpu_pairs + 1], pool)
        nw_s = cpu_pairs * nr * nc
        cpu = {"value": nw_s / dtc, "unit": "windows/s", "cores": pool.n_workers, "kind": "port", "sample": cpu_sample_note(cpu_pairs, pool),
               "host_cores": cores}
        ok = np.isfinite(u) & np.isfinite(hu[:cpu_pairs])
        rmse = {"u_px": float(np.sqrt(np.mean((hu[:cpu_pairs][ok] - u[ok]) ** 2))),
                "v_px": float(np.sqrt(np.mean((hv[:cpu_pairs][ok] - v[ok]) ** 2))),
                "windows": int(ok.sum()), "nan_mask_equal": bool(np.array_equal(np.isnan(u), np.isnan(hu[:cpu_pairs]))),
                "of": "the get_b2piv (e2e) result, converted back to px / frame"}
        pool.shutdown()

    cfg = config_dict(world)
    line = {
        "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "gather": gather_how,
        "roofline": roofline, "fp32": fp32, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "pyorc_b200.velocimetry.get_b2piv(pageable numpy DataArray) -> Dataset" + ("; then NCCL gather of the four fields to every rank's host memory" if world > 1 else ""),
                "pcie": pcie, "pinned_engine_pairs": pinned, "host_cores": host_cores, "stage_threads_per_rank": stage_threads},
        "gpu_launches": int(launches), "clocks": clocks, "rmse_vs_oracle": rmse, "other_configs": other, "sharded_configs": sharded,
        "ensemble_multi_gpu": ens,
    }
    emit(line)
    if peer is not None:
        peer.close()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
