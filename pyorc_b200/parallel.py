"""Frame-pair sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Per-time-step PIV is embarrassingly parallel over frame pairs (pyorc/velocimetry/ffpiv.py:399-440: chunks only share
a 1-frame halo, :140), so rank *r* of *R* owns the contiguous pair range ``[r*P/R, (r+1)*P/R)`` and frames
``[start, end]`` inclusive.  The only collectives are the broadcast of the pair-range table and the gather of the
16 B/window results (NCCL on GPUs; the same code runs over gloo for CPU tests with an injected compute function).
Ensemble mode reduce-scatters the plane sums over the window axis, each rank peak-fits its slice, and the 8 B/window results
are all-gathered (:func:`ensemble_sharded`).
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_pairs(n_pairs: int, world_size: int, rank: Optional[int] = None):
    """Contiguous, balanced pair ranges.  Returns the ``[world_size, 2]`` (start, stop) table or one row."""
    if n_pairs < 0 or world_size < 1:
        raise ValueError("n_pairs >= 0 and world_size >= 1 required")
    base, rem = divmod(n_pairs, world_size)
    starts = [r * base + min(r, rem) for r in range(world_size)]
    table = np.array([[s, s + base + (1 if r < rem else 0)] for r, s in enumerate(starts)], dtype=np.int64)
    return table if rank is None else tuple(int(v) for v in table[rank])


def frame_range(pair_range: Tuple[int, int]) -> Tuple[int, int]:
    """Frames (inclusive halo) a pair range needs: pairs [a, b) -> frames [a, b] i.e. slice a : b+1."""
    a, b = pair_range
    return a, (b + 1 if b > a else a)


def scatter_pair_table(n_pairs: int, group=None, device=None):
    """Rank 0 computes the pair-range table and broadcasts it (the 'scatter of the frame-pair index list')."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = torch.zeros((world, 2), dtype=torch.int64, device=device)
    if rank == 0:
        t.copy_(torch.from_numpy(shard_pairs(n_pairs, world)))
    dist.broadcast(t, src=0, group=group)
    return t.cpu().numpy()


def gather_fields(local, n_pairs_total: int, table: np.ndarray, group=None):
    """All-gather per-rank result stacks ``[4, pairs_r, rows, cols]`` into ``[4, n_pairs_total, rows, cols]``
    (ragged shards are padded to the largest shard for the collective, then trimmed)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    counts = (table[:, 1] - table[:, 0]).astype(int)
    pmax = int(counts.max()) if len(counts) else 0
    # equal shards (the usual case): every field is gathered straight into its final place - rank r's pairs are the r-th
    # contiguous block of out[k] - with one collective per field and no staging copies (stack / pad / trim cost 10 % of a
    # 100-pair 1080p step on 8 GPUs).  `local` may be the engine's tuple of four tensors or a stacked tensor.
    if len(counts) and int(counts.min()) == pmax and pmax > 0 and all(local[k].is_contiguous() for k in range(len(local))):
        nf = len(local)
        _, rows, cols = local[0].shape
        out = torch.empty((nf, n_pairs_total, rows, cols), dtype=local[0].dtype, device=local[0].device)
        for k in range(nf):
            dist.all_gather_into_tensor(out[k], local[k], group=group)
        return out
    if isinstance(local, (tuple, list)):
        local = torch.stack(list(local))
    nf, _, rows, cols = local.shape
    pad = torch.zeros((nf, pmax, rows, cols), dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.empty((nf, n_pairs_total, rows, cols), dtype=local.dtype, device=local.device)
    for r in range(world):
        a, b = int(table[r, 0]), int(table[r, 1])
        out[:, a:b] = bufs[r][:, : b - a]
    return out


def _result_block(res):
    """The contiguous ``[4, n_pairs, rows, cols]`` block behind the four fields ``Engine.pairs`` returned (they are views of one
    allocation); any other 4-tuple of tensors is stacked (one copy)."""
    import torch

    if isinstance(res, torch.Tensor):
        return res
    f0 = res[0]
    n = f0.numel()
    if (len(res) == 4 and f0.dim() == 3 and n > 0 and all(r.is_contiguous() and r.shape == f0.shape and r.dtype == f0.dtype for r in res)
            and all(res[k].data_ptr() == f0.data_ptr() + k * n * f0.element_size() for k in range(4))
            and all(res[k].untyped_storage().data_ptr() == f0.untyped_storage().data_ptr() for k in range(4))):
        return f0.as_strided((4, *f0.shape), (n, *f0.stride()), f0.storage_offset())
    return torch.stack(list(res))


class PeerGather:
    """The result gather as P2P stores of our own kernels: every rank writes its 16 B / window straight into the gather buffer
    of EVERY rank (torch symmetric memory = peer-mapped HBM over NVLink / NVSwitch), so no collective follows the compute - only
    a cross-rank completion barrier before a buffer is read.  Two modes:

    * ``"push"`` (default): the PIV kernel writes this rank's own block; a small copy kernel on the consumer stream forwards it
      to every rank's slot (16-byte stores) while the compute stream is already in the next step.
    * ``"fused"``: the PIV kernel's epilogue stores each result to all ranks itself.  One launch less, but a kernel that writes
      peer memory waits for NVLink's acknowledgements when it ends, on the compute stream: +25 .. 30 us per launch at N = 2 and
      +64 us at N = 8 on a 1.63 ms step, however the stores are issued (per frame by the fitting threads, spread over a warp's
      lanes, or batched per work unit - tools/scale_probe.py), against +3 us for the push.

    The buffers form a ring of ``depth`` slots and the barrier runs on a separate high-priority CONSUMER stream, so the compute
    stream never waits for the other ranks in steady state (round 1 had one buffer and a barrier on the compute stream after
    every step: 6 % of an 8-GPU step):

        pg = PeerGather(engine, n_pairs_total, table)        # after engine.plan(...); collective (rendezvous)
        for step in ...:
            pg.begin()                                        # next slot; waits only if that slot is still being read anywhere
            res = engine.pairs(d_frames_of_this_rank, ws, ov) # "fused": results land in the slot on every rank
            fields, ready = pg.end(res)                       # consumer stream: kernel done -> ("push": copy to all) -> barrier -> `ready`
            ...                                               # whoever reads `fields` [4, n_pairs_total, rows, cols] first waits
                                                              # for `ready` (stream.wait_event / ready.synchronize())
        pg.wait()                                             # or: make the current stream wait for the last slot (blocking use)

    Slot reuse: step s + depth writes the slot of step s.  It may start once every rank has stopped reading that slot, which a
    rank has (in consumer-stream order) when it reaches the barrier of step s + 1; so ``begin`` makes the compute stream wait
    for the local ``ready`` event of step s + 1 - with ``depth = 3`` that is two steps behind the compute.  A consumer reads a slot
    on the consumer stream (``with torch.cuda.stream(pg.consumer)``) or makes the consumer stream wait for its read before the
    next ``end()``.  The NCCL path (:func:`gather_fields`) remains for ragged use and for backends without peer access."""

    def __init__(self, engine, n_pairs_total: int, table: np.ndarray, group=None, depth: int = 3, mode: str = "push"):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        if engine._plan is None:
            raise RuntimeError("plan the engine before creating a PeerGather")
        if depth < 2:
            raise ValueError("depth >= 2 (a slot is rewritten only after the barrier of the NEXT step, see the class docstring)")
        rows, cols = engine._plan[1]
        group = group if group is not None else dist.group.WORLD
        rank = dist.get_rank(group)
        dev = torch.device("cuda", engine.device)
        if mode not in ("fused", "push"):
            raise ValueError("mode must be 'fused' or 'push'")
        self.mode = mode
        self.engine, self.depth, self.n_pairs_total, self.pair_offset = engine, int(depth), int(n_pairs_total), int(table[rank, 0])
        self.ring = symm_mem.empty((self.depth, 4, self.n_pairs_total, rows, cols), dtype=torch.float32, device=dev)
        self.handle = symm_mem.rendezvous(self.ring, group)
        self._ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(self._ptrs) > 8:
            raise NotImplementedError("at most 8 peers (one box)")
        self._slot_bytes = self.ring[0].numel() * 4
        self.consumer = torch.cuda.Stream(device=dev, priority=-1)
        self._ready = [None] * self.depth      # event per slot: every rank's results of the step that used it have landed here
        self._step = self._ended = -1
        self._device = dev
        self.out = self.ring[0]

    def begin(self):
        """Start the next step: pick its slot, wait (compute stream) until no rank still reads it, point the kernels at it."""
        import torch

        self._step += 1
        slot = self._step % self.depth
        nxt = self._ready[(slot + 1) % self.depth]
        if self._step >= self.depth and nxt is not None:
            torch.cuda.current_stream(self._device).wait_event(nxt)     # barrier of step (s - depth + 1) has been passed here
        if self.mode == "fused":
            self.engine.set_peer_outputs([p + slot * self._slot_bytes for p in self._ptrs], self.n_pairs_total, self.pair_offset)
        self.out = self.ring[slot]
        return slot

    def push(self, local, pair_offset=None):
        """``mode="push"``: forward ``local`` - what ``engine.pairs`` returned (its four fields are views of one contiguous
        ``[4, n_pairs, rows, cols]`` block) or such a block itself - to every rank's current slot at ``pair_offset`` of the gathered time axis (default:
        this rank's first pair), on the consumer stream, after the work queued so far on the current stream.  A step made of several
        chunk launches pushes each chunk with its own offset and then calls :meth:`end` without a block."""
        import torch

        if self.mode != "push":
            raise RuntimeError("PeerGather(mode='fused') stores from the kernel epilogue; push() is for mode='push'")
        if self._step < 0:
            raise RuntimeError("begin() first")
        slot = self._step % self.depth
        local = _result_block(local)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self._device))
        self.consumer.wait_event(done)
        local.record_stream(self.consumer)
        self.engine.peer_push(local, [p + slot * self._slot_bytes for p in self._ptrs], self.n_pairs_total,
                              self.pair_offset if pair_offset is None else int(pair_offset), stream=self.consumer)
        self._pushed = self._step

    def end(self, local=None, pair_offset=None):
        """After ``engine.pairs``: completion barrier of this step on the consumer stream.  Returns the slot's tensor and the event
        after which it holds every rank's results.  ``mode="push"``: pass the result block here (or to :meth:`push` before)."""
        import torch

        if self._step < 0:
            raise RuntimeError("begin() first")
        slot = self._step % self.depth
        if local is not None:
            self.push(local, pair_offset)
        elif self.mode == "push" and getattr(self, "_pushed", -1) != self._step:
            raise ValueError("mode='push': pass the local result block to end() or push() first")
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self._device))
        with torch.cuda.stream(self.consumer):
            self.consumer.wait_event(done)
            self.handle.barrier()
            ready = torch.cuda.Event()
            ready.record(self.consumer)
        self._ready[slot] = ready
        self._ended = self._step
        return self.ring[slot], ready

    def wait(self):
        """Blocking use (one step at a time): ``end()`` if it has not been called for this step, then the CURRENT stream waits until
        all ranks' results of the step are in ``self.out``; returns it."""
        import torch

        if self._step < 0:
            raise RuntimeError("begin() / engine.pairs() first")
        slot = self._step % self.depth
        if self._ended != self._step:
            self.end()      # mode="push": the block must have been push()ed
        torch.cuda.current_stream(self._device).wait_event(self._ready[slot])
        return self.ring[slot]

    def drain(self):
        """Host-side: wait for the consumer stream (all outstanding barriers)."""
        self.consumer.synchronize()

    def close(self):
        self.drain()
        self.engine.set_peer_outputs(None, 1, 0)


def piv_pairs_sharded(frames_for_rank: Callable[[int, int], object], n_pairs_total: int,
                      compute: Callable[[object], Tuple], group=None, device=None):
    """Distributed per-time-step PIV.

    ``frames_for_rank(f0, f1)`` returns this rank's frames ``[f0, f1)`` (device tensor or ndarray);
    ``compute(frames)`` returns ``(u, v, corr_max, s2n)`` stacks ``[pairs, rows, cols]`` (the engine's ``pairs``).
    Returns the gathered ``[4, n_pairs_total, rows, cols]`` tensor on every rank.
    """
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    table = scatter_pair_table(n_pairs_total, group=group, device=device)
    a, b = int(table[rank, 0]), int(table[rank, 1])
    f0, f1 = frame_range((a, b))
    if b > a:
        res = compute(frames_for_rank(f0, f1))
        local = torch.stack([torch.as_tensor(r) for r in res]).to(device if device is not None else "cpu")
    else:
        local = None
    # shapes must agree across ranks even when a rank has no pair: share (rows, cols)
    shp = torch.zeros(2, dtype=torch.int64, device=device)
    if local is not None:
        shp[0], shp[1] = local.shape[2], local.shape[3]
    dist.all_reduce(shp, op=dist.ReduceOp.MAX, group=group)
    if local is None:
        local = torch.zeros((4, 0, int(shp[0]), int(shp[1])), dtype=torch.float32, device=device)
    return gather_fields(local.float(), n_pairs_total, table, group=group)


def window_slices(n_windows: int, world_size: int) -> np.ndarray:
    """``[world_size, 2]`` (first, stop) window ranges of the ensemble peak fit: ``n_windows // world_size`` windows per rank -
    the equal blocks a reduce-scatter needs - and the remainder (fewer than ``world_size`` windows) on the last rank."""
    base = n_windows // world_size
    t = np.array([[r * base, (r + 1) * base] for r in range(world_size)], dtype=np.int64)
    t[-1, 1] = n_windows
    return t


def ensemble_reduce_finish(plane, count, finish: Callable[[int, int], Tuple], group=None):
    """Ensemble mode across ranks: every rank holds partial plane sums ``plane [n_windows, wy * wx]`` and
    valid counts ``count [n_windows]`` of ITS frame pairs (pyorc/velocimetry/ffpiv.py:361-363 accumulates them over chunks).

    1. ``count`` is all-reduced (4 B / window);
    2. ``plane`` is REDUCE-SCATTERED over the window axis: rank r receives the global sums of its window slice
       (:func:`window_slices`) - each rank moves 1/R of the planes instead of all of them - written back into its rows of
       ``plane``; the remainder rows (fewer than R windows) are all-reduced and belong to the last rank;
    3. ``finish(first, n) -> (u, v)`` ([n] tensors: count filter, mean plane, peak fit of those rows - ffpiv.py:280-282, :324)
       runs on the slice;
    4. the 8 B / window results are all-gathered.

    Returns ``u, v`` ``[n_windows]`` (the same on every rank) and the reduced ``count``.  Works on any backend (NCCL with the
    engine's accumulators; gloo in the CPU tests with a numpy ``finish``)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nw = plane.shape[0]
    dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    table = window_slices(nw, world)
    base = nw // world
    if world > 1:
        if base > 0:
            mine = torch.empty((base,) + tuple(plane.shape[1:]), dtype=plane.dtype, device=plane.device)
            dist.reduce_scatter_tensor(mine, plane[: base * world], op=dist.ReduceOp.SUM, group=group)
            plane[rank * base : (rank + 1) * base].copy_(mine)
            del mine
        if nw > base * world:
            tail = plane[base * world :]
            dist.all_reduce(tail, op=dist.ReduceOp.SUM, group=group)
    first, stop = int(table[rank, 0]), int(table[rank, 1])
    u, v = finish(first, stop - first)
    smax = int((table[:, 1] - table[:, 0]).max())
    send = torch.zeros((2, smax), dtype=torch.float32, device=plane.device)
    send[0, : stop - first] = torch.as_tensor(u, dtype=torch.float32, device=plane.device)
    send[1, : stop - first] = torch.as_tensor(v, dtype=torch.float32, device=plane.device)
    recv = torch.empty((world * 2, smax), dtype=torch.float32, device=plane.device)   # concatenation along dim 0 (all backends)
    if world > 1:
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        recv.copy_(send)
    recv = recv.view(world, 2, smax)
    out = torch.empty((2, nw), dtype=torch.float32, device=plane.device)
    for r in range(world):
        a, b = int(table[r, 0]), int(table[r, 1])
        out[:, a:b] = recv[r, :, : b - a]
    return out[0], out[1], count


def aggregate_ensemble(corr_max_concat: np.ndarray, s2n_concat: np.ndarray, corr_count: np.ndarray, min_count: float, n_rows: int,
                       n_cols: int):
    """The host half of ``aggregate_results`` (ffpiv.py:263-286): count filter on the per-pair maxima, time means of the
    per-pair ``corr_max`` and ``s2n`` ``[pairs, n_windows]`` -> ``[1, n_rows, n_cols]`` each."""
    import warnings

    corr_max_concat = np.array(corr_max_concat, dtype=np.float32, copy=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        corr_max_concat[:, np.asarray(corr_count).reshape(-1) < min_count] = np.nan
        corr_max_mean = np.nanmean(corr_max_concat, axis=0).reshape(-1, n_rows, n_cols)
        s2n_mean = np.nanmean(s2n_concat, axis=0).reshape(-1, n_rows, n_cols)
    return corr_max_mean, s2n_mean


def ensemble_sharded(engine, frames, window_size, overlap, n_pairs_total: int, table: np.ndarray, corr_min=0.2, s2n_min=3.0,
                     count_min=0.2, n_chunks_total: Optional[int] = None, signal_threshold=None, group=None):
    """Ensemble-correlation PIV with the frame pairs sharded over the ranks (one process per GPU).

    ``frames``: this rank's device-resident frames (its pair range of ``table`` plus the 1-frame halo; None for an empty
    shard).  Every rank accumulates the planes of its pairs in its engine (one launch), the plane sums are reduce-scattered
    over the window axis, each rank peak-fits its slice, and the results plus the per-pair ``corr_max`` / ``s2n`` statistics
    are all-gathered (:func:`ensemble_reduce_finish`, :func:`gather_fields`).  ``n_chunks_total``: the reference's
    ``n_frames`` of the count filter - its number of CHUNKS (ffpiv.py:373), default one chunk per rank.

    Returns ``u, v`` [px / frame], ``corr_max_mean``, ``s2n_mean`` as numpy ``[1, n_rows, n_cols]`` (identical on every rank)
    and the valid count ``[n_windows]``."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", engine.device)
    a, b = int(table[rank, 0]), int(table[rank, 1])
    n_chunks_total = world if n_chunks_total is None else int(n_chunks_total)
    min_count = float(count_min) * n_chunks_total
    if frames is not None:
        dims, dtype = tuple(frames.shape[-2:]), (np.uint8 if frames.dtype == torch.uint8 else np.float32)
    else:
        dims, dtype = engine._plan[0][0], (np.uint8 if engine._plan[0][3] == 0 else np.float32)
    nr, nc = engine.ens_begin(dims, window_size, overlap, dtype, device_ordered=True)
    if b > a:
        cm, sn = engine.ens_add(frames, window_size, overlap, corr_min=corr_min, s2n_min=s2n_min, signal_threshold=signal_threshold)
        local = torch.stack([cm, sn]).view(2, b - a, nr, nc)
    else:
        local = torch.zeros((2, 0, nr, nc), dtype=torch.float32, device=dev)
    plane, count = engine.ens_accumulators()
    u, v, count = ensemble_reduce_finish(plane, count, lambda first, n: engine.ens_finish_device(min_count, first, n), group=group)
    stats = gather_fields(local, n_pairs_total, table, group=group)          # [2, pairs_total, rows, cols]
    count_np = count.cpu().numpy()
    cm_mean, sn_mean = aggregate_ensemble(stats[0].reshape(n_pairs_total, -1).cpu().numpy(), stats[1].reshape(n_pairs_total, -1).cpu().numpy(),
                                          count_np, min_count, nr, nc)
    return u.cpu().numpy().reshape(1, nr, nc), v.cpu().numpy().reshape(1, nr, nc), cm_mean, sn_mean, count_np
