#!/usr/bin/env python
"""bench.py - interrogation windows / s of the fused LSPIV engine (BASELINE.json metric).

Workload (BASELINE.json configs[1]): synthetic 1080p, 100 frame pairs, 64x64 windows, 50 % overlap, single pass,
uint8 frames.  A "step" is one pass of the hot path over that batch (188 800 windows) on every GPU (weak scaling:
each rank owns its own 100-pair shard, frame pairs shard with no data-path collective; for N > 1 the 16 B/window
results are all-gathered over NCCL inside the timed region).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port, host cores)

Prints ONE JSON line (rank 0).  `value` = windows/s with frames resident in HBM; `e2e` = the same metric through
the host API (pinned host frames, H2D + D2H inside the timed region); `roofline` = algorithmic HBM bytes of the
fused kernel against the measured copy bandwidth; `cpu_baseline` = the oracle port on the host cores.
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
WORKLOAD = "synthetic 1080p, 100 frame pairs, 64x64 windows 50% overlap, single-pass, 1xB200 (per GPU)"
# compulsory HBM traffic and FFT flops per window
B_ALG = 2 * (WS[0] - OV[0]) * (WS[1] - OV[1]) * 1 + 16                      # 2064 B (uint8)
F_ALG = 3 * 2.5 * (WS[0] * WS[1]) * np.log2(WS[0] * WS[1]) + 6 * WS[0] * (WS[1] // 2 + 1)  # 381 312 flop


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


def cpu_reference(frames_np, workers):
    """The reference's CPU path (oracle port: float64 pocketfft, all passes over memory pyorc makes) on `frames_np`."""
    from oracle import ffpiv_oracle as O

    t0 = time.perf_counter()
    u, v, c, s = O.cpu_reference_pairs(frames_np, WS, OV, workers=workers)
    dt = time.perf_counter() - t0
    return (u, v, c, s), dt


def run_reference(args, rank, world):
    """--impl reference: time the CPU path on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from pyorc_b200 import synth

    cores = os.cpu_count() or 1
    sample_pairs = args.cpu_pairs
    frames = synth.particle_frames(sample_pairs + 1, H, W, dtype=np.uint8)
    nwin = sample_pairs * 32 * 59
    cpu_reference(frames[:2], cores)  # warm-up (imports, pocketfft plans)
    ts = []
    for _ in range(args.steps):
        _, dt = cpu_reference(frames, cores)
        ts.append(dt)
    total = float(np.sum(ts))
    val = nwin * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frame": [H, W], "window": list(WS), "overlap": list(OV), "input_dtype": "uint8",
                   "sample": f"{sample_pairs} frame pairs of the workload per step ({nwin} windows)"},
        "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_pairs} of the 100 frame pairs per step; restated ffpiv CPU path (upstream ffpiv/rocket-fft unavailable offline), float64 pocketfft, one thread per frame pair on {cores} cores"},
        "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def other_configs(eng, dev):
    """windows/s of BASELINE.json configs[0], [2], [3], [4] (their geometry, a shard of frame pairs each, frames resident in
    HBM, CUDA events, median of 5 after 2 warm-ups) with the algorithmic HBM bytes / fp32 flops of SURVEY.md §8(d)."""
    import torch

    from pyorc_b200 import synth

    peak, _ = measured_peaks()

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), out

    rows = []
    cases = [("configs[0] geometry: 475x371, 32x32 / 50 %", 475, 371, (32, 32), (16, 16), 3, False),
             ("configs[2] single pass: 1080p, 32x32 / 75 %", 1080, 1920, (32, 32), (24, 24), 21, False),
             ("configs[2] two-pass (64x64 / 75 % -> 32x32 / 75 %, discrete window offset)", 1080, 1920, (32, 32), (24, 24), 21, True),
             ("configs[3] geometry: 4K, 64x64 / 50 %", 2160, 3840, (64, 64), (32, 32), 21, False),
             ("configs[4] geometry: 8K, 128x128 / 50 %", 4320, 7680, (128, 128), (64, 64), 11, False)]
    for name, h, w, ws, ov, n, two_pass in cases:
        try:
            fr = synth.particle_frames_torch(n, h, w, dev, dtype="uint8")
            if two_pass:
                ms, out = timed(lambda: eng.pairs_two_pass(fr, ((64, 64), (48, 48)), (ws, ov)))
            else:
                ms, out = timed(lambda: eng.pairs(fr, ws, ov))
            nwin = int(out[0].numel())
            b_alg = 2 * (ws[0] - ov[0]) * (ws[1] - ov[1]) + 16
            f_alg = 3 * 2.5 * ws[0] * ws[1] * np.log2(ws[0] * ws[1]) + 6 * ws[0] * (ws[1] // 2 + 1)
            rows.append({"config": name, "pairs": n - 1, "windows": nwin, "ms": ms, "windows_per_s": nwin / (ms * 1e-3),
                         "hbm_frac": b_alg * nwin / (ms * 1e-3) / 1e9 / peak,
                         "fp32_tflops": f_alg * nwin / (ms * 1e-3) / 1e12})
            del fr, out
        except Exception as exc:   # a shard that does not fit must not take the headline down with it
            rows.append({"config": name, "error": str(exc)[:200]})
    eng.plan((H, W), WS, OV, np.uint8)
    return rows


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
    ap.add_argument("--nccl-gather", action="store_true", help="N > 1: gather results with NCCL instead of the fused P2P stores")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of BASELINE.json configs[0], [2], [3], [4]")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="frame pairs of the workload the CPU baseline times (0: min(max(4, cores), 16))")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.cpu_pairs <= 0:
        args.cpu_pairs = min(max(4, os.cpu_count() or 1), 16)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    from pyorc_b200 import parallel, synth
    from pyorc_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local_rank)
    n_frames = N_PAIRS + 1
    frames = synth.particle_frames_torch(n_frames, H, W, dev, dtype="uint8", seed=synth.SEED + rank)
    nr, nc = eng.plan((H, W), WS, OV, np.uint8)
    nwin_rank = N_PAIRS * nr * nc
    table = parallel.shard_pairs(N_PAIRS * world, world)

    # N > 1: the gather of the 16 B / window results is fused into the kernel epilogue (P2P stores into every rank's
    # symmetric-memory buffer, then a device-side barrier); NCCL all-gather when peer memory is not available
    peer = None
    gather_how = None
    if world > 1 and not args.nccl_gather:
        try:
            peer = parallel.PeerGather(eng, N_PAIRS * world, table)
            gather_how = "fused: P2P stores from the kernel epilogue into symmetric memory + barrier"
        except Exception as exc:   # no peer access / symmetric memory: keep the collective
            peer = None
            gather_how = f"nccl all_gather_into_tensor per field (symmetric memory unavailable: {str(exc)[:80]})"
    elif world > 1:
        gather_how = "nccl all_gather_into_tensor per field"

    def step_device():
        res = eng.pairs(frames, WS, OV)
        if peer is not None:
            return peer.wait()
        if world > 1:
            return parallel.gather_fields(res, N_PAIRS * world, table)   # four fields, gathered in place
        return res

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
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
        step_device()
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    if peer is not None:
        sync_all()
        gathered = step_device()
        sync_all()
        # every rank must hold every rank's results: compare with an NCCL gather of the same step
        ref_g = parallel.gather_fields(eng.pairs(frames, WS, OV), N_PAIRS * world, table)
        same = bool(torch.equal(torch.nan_to_num(gathered), torch.nan_to_num(ref_g)))
        gather_how += f"; equals NCCL gather: {same}"
        sync_all()
        peer.close()
    # kernel-only average launch duration (same stream, no gather) for the roofline
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    k0.record()
    for _ in range(args.steps):
        eng.pairs(frames, WS, OV)
    k1.record()
    torch.cuda.synchronize(dev)
    ms_kernel = k0.elapsed_time(k1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = nwin_rank * world / (ms_step * 1e-3)

    # ---- end to end through the host API (pinned host frames -> H2D -> kernel -> D2H) ---------------------------
    host = eng.pinned_empty((n_frames, H, W), np.uint8)
    host[...] = frames.cpu().numpy()
    for _ in range(2):
        eng.pairs(host, WS, OV)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hu, hv, hc, hs = eng.pairs(host, WS, OV)   # synchronous: returns after the D2H of the four fields
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = nwin_rank * world * args.steps / e2e_s
    h2d = int(host.nbytes)
    d2h = int(4 * hu.nbytes)
    # the same call on ORDINARY numpy memory (what pyorc hands over: frame_chunk.values): the engine stages it through its
    # page-locked ring with a few copy threads (N = 1 only; reported beside the pinned number, not instead of it)
    pageable = None
    if world == 1:
        host_pg = np.array(host, copy=True)
        for _ in range(2):
            eng.pairs(host_pg, WS, OV)
        reps = min(args.steps, 10)
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.pairs(host_pg, WS, OV)
        dt = (time.perf_counter() - t0) / reps
        pageable = {"value": nwin_rank / dt, "ms_per_step": 1e3 * dt}
        del host_pg

    # the PCIe floor under e2e: the same number of bytes, pinned host -> device, nothing else (N = 1 only)
    pcie = None
    if world == 1:
        hp = torch.empty(frames.numel(), dtype=torch.uint8, pin_memory=True)
        dd = torch.empty_like(frames).view(-1)
        dd.copy_(hp, non_blocking=True)
        torch.cuda.synchronize(dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            dd.copy_(hp, non_blocking=True)
        c1.record()
        torch.cuda.synchronize(dev)
        h2d_ms = c0.elapsed_time(c1) / 3
        pcie = {"h2d_only_ms": h2d_ms, "h2d_gbs": hp.numel() / h2d_ms / 1e6}
        del hp, dd

    if rank != 0:
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel -----------------------------------------------------------------
    peak, peak_src = measured_peaks()
    achieved = B_ALG * nwin_rank / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "piv_rows_kernel<64> (row-per-thread, TMA)", "ms_per_launch": ms_kernel, "alg_bytes_per_window": B_ALG,
                "windows_per_launch": nwin_rank, "peak_source": peak_src,
                "note": "fused kernel is fp32-issue/shared-memory bound by construction (SURVEY.md §8d); see fp32"}
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    fp32 = {"achieved": F_ALG * nwin_rank / (ms_kernel * 1e-3) / 1e12, "unit": "TFLOP/s", "alg_flop_per_window": F_ALG,
            "peak_nominal": 148 * 128 * 2 * 1.965e9 / 1e12}
    fp32["frac_of_nominal"] = fp32["achieved"] / fp32["peak_nominal"]

    # ---- CPU baseline (oracle port) on a bounded sample + parity of the GPU result on that sample ---------------
    cpu = None
    rmse = None
    if world == 1:
        cores = os.cpu_count() or 1
        sample = host[: args.cpu_pairs + 1]
        (u, v, c, s), dt = cpu_reference(sample, cores)
        nw_s = args.cpu_pairs * nr * nc
        cpu = {"value": nw_s / dt, "unit": "windows/s", "cores": cores, "kind": "port",
               "sample": f"first {args.cpu_pairs} of the 100 frame pairs ({nw_s} windows); restated ffpiv CPU path "
                         f"(float64 pocketfft, one thread per frame pair on {cores} cores); upstream ffpiv/rocket-fft not installable offline"}
        ok = np.isfinite(u) & np.isfinite(hu[: args.cpu_pairs])
        rmse = {"u_px": float(np.sqrt(np.mean((hu[: args.cpu_pairs][ok] - u[ok]) ** 2))),
                "v_px": float(np.sqrt(np.mean((hv[: args.cpu_pairs][ok] - v[ok]) ** 2))),
                "windows": int(ok.sum()), "nan_mask_equal": bool(np.array_equal(np.isnan(u), np.isnan(hu[: args.cpu_pairs])))}

    # ---- the other BASELINE.json configurations, device resident, one shard of frames each (N = 1 only; reported beside
    # the headline, not part of it: the metric is quoted on configs[1]) ------------------------------------------------------
    other = None
    if world == 1 and not args.no_other_configs:
        other = other_configs(eng, dev)

    line = {
        "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frame": [H, W], "pairs_per_gpu": N_PAIRS, "window": list(WS), "overlap": list(OV),
                   "input_dtype": "uint8", "windows_per_step": nwin_rank * world,
                   "l2": f"inputs {frames.numel() / 1e6:.0f} MB per GPU > 126 MB L2 (no flush needed)",
                   "parallelism": f"frame-pair shard x{world}", "gather": gather_how},
        "roofline": roofline, "fp32": fp32, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "ms_per_step": 1e3 * e2e_s / args.steps, "api": "pyorc_b200.engine.Engine.pairs(numpy pinned)", "pcie": pcie,
                "pageable_numpy": pageable},
        "gpu_launches": int(launches), "clocks": clocks, "rmse_vs_oracle": rmse, "other_configs": other,
    }
    emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
