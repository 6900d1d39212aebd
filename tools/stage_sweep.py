"""Staging of pageable host frames (csrc/stager.h) against its knobs, on 1 .. N ranks AT THE SAME TIME (torchrun): ms per
100-pair 1080p Engine.pairs call from ordinary numpy memory, max over the ranks, beside the page-locked call and the bare H2D.

    python tools/stage_sweep.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/stage_sweep.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pyorc_b200 import synth
from pyorc_b200.engine import Engine

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
e = Engine(local)
H, W, n = 1080, 1920, 101
WS, OV = (64, 64), (32, 32)
fr = synth.particle_frames_torch(n, H, W, dev, dtype="uint8", seed=synth.SEED + rank)
pinned = e.pinned_empty((n, H, W), np.uint8)
pinned[...] = fr.cpu().numpy()
pageable = np.array(pinned, copy=True)
cores = len(os.sched_getaffinity(0))
reps = int(os.environ.get("REPS", 8))


def sync():
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)


def run(host):
    for _ in range(2):
        e.pairs(host, WS, OV)
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        e.pairs(host, WS, OV)
    t = torch.tensor([(time.perf_counter() - t0) / reps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return 1e3 * float(t.item())


def say(*a):
    if rank == 0:
        print(*a, flush=True)


say(f"ranks {world}, host cores {cores}, {reps} calls per row; ms per 100-pair 1080p Engine.pairs call (max over ranks)")
say(f"page-locked frames: {run(pinned):.3f}")
dd = torch.empty_like(fr)
hp = torch.from_numpy(pinned)
sync()
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0.record()
for _ in range(5):
    dd.copy_(hp, non_blocking=True)
c1.record()
torch.cuda.synchronize(dev)
t = torch.tensor([c0.elapsed_time(c1) / 5], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
say(f"bare H2D of the same bytes: {float(t.item()):.3f}")

if os.environ.get("AB"):
    # same-box A/B of a few settings, alternating, three rounds: (threads, mode, slice KB, groups, non-temporal)
    cfgs = [(8, 0, 0, 0, 1), (8, 1, 1024, 3, 1), (8, 1, 2048, 3, 1), (8, 1, 512, 6, 1), (12, 1, 1024, 3, 1), (6, 1, 1024, 3, 1), (8, 1, 1024, 3, 0)]
    for rnd in range(3):
        row = []
        for th, mode, kb, groups, nt in cfgs:
            e.set_option("stage_threads", th)
            e.set_option("stage_mode", mode)
            if mode:
                e.set_option("stage_slice_kb", kb)
                e.set_option("stage_groups", groups)
                e.set_option("stage_nt", nt)
            row.append(f"[t{th} m{mode} {kb}KB g{groups} nt{nt}] {run(pageable):.3f}")
        say(f"round {rnd}: " + "  ".join(row))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0)
per_rank = max(2, min(8, (cores - world) // world))
thread_list = sorted({per_rank, max(2, per_rank // 2)}, reverse=True)
for th in thread_list:
    e.set_option("stage_threads", th)
    e.set_option("stage_mode", 0)
    say(f"threads {th} | round-1 pool (3 x 6 MB, non-temporal, condition variables): {run(pageable):.3f}")
    e.set_option("stage_mode", 1)
    for nt in (0, 1):
        e.set_option("stage_nt", nt)
        for kb in (64, 128, 256, 512, 1024):
            row = []
            for groups in (3, 4, 6, 10):
                e.set_option("stage_slice_kb", kb)
                e.set_option("stage_groups", groups)
                row.append(f"g{groups} {run(pageable):.3f}")
            say(f"threads {th} | stager nt={nt} slice {kb:4d} KB (H2D of {kb * th / 1024:.2f} MB): " + "  ".join(row))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
