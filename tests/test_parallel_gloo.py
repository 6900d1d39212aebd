"""CPU, world_size 2 over gloo: frame-pair sharding, pair-table broadcast and result gather (pyorc_b200.parallel) with
the compute step injected (oracle) - the same code path runs over NCCL with the CUDA engine in bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyorc_b200 import parallel


def test_shard_pairs_properties():
    for n in (0, 1, 7, 100, 101):
        for w in (1, 2, 3, 8):
            t = parallel.shard_pairs(n, w)
            assert t.shape == (w, 2) and t[0, 0] == 0 and t[-1, 1] == n
            assert np.all(t[1:, 0] == t[:-1, 1])
            sizes = t[:, 1] - t[:, 0]
            assert sizes.max() - sizes.min() <= 1
    assert parallel.shard_pairs(100, 8, rank=3) == (38, 51) or parallel.shard_pairs(100, 8, rank=3)[1] - parallel.shard_pairs(100, 8, rank=3)[0] in (12, 13)
    assert parallel.frame_range((4, 9)) == (4, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ffpiv_oracle as O
    from pyorc_b200 import synth

    O.CLIP_NORMALIZED = True
    frames = synth.particle_frames(n_pairs + 1, 80, 112, dtype=np.uint8)
    ws, ov = (32, 32), (16, 16)
    nr, nc = O.get_array_shape((80, 112), ws, ov)

    def compute(fr):
        u, v, c, s = O.uv_timestep(fr, nc, nr, ws, ov)
        return u.astype(np.float32), v.astype(np.float32), c, s

    out = parallel.piv_pairs_sharded(lambda f0, f1: frames[f0:f1], n_pairs, compute)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), out.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [5, 1, 4])   # ragged shards, an empty shard, equal shards (in-place gather)
def test_sharded_piv_over_gloo_equals_single_process(tmp_path, n_pairs):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_pairs, str(tmp_path)), nprocs=world, join=True)
    from oracle import ffpiv_oracle as O
    from pyorc_b200 import synth

    O.CLIP_NORMALIZED = True
    frames = synth.particle_frames(n_pairs + 1, 80, 112, dtype=np.uint8)
    nr, nc = O.get_array_shape((80, 112), (32, 32), (16, 16))
    u, v, c, s = O.uv_timestep(frames, nc, nr, (32, 32), (16, 16))
    ref = np.stack([u.astype(np.float32), v.astype(np.float32), c, s])
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.shape == ref.shape
        assert np.array_equal(got, ref, equal_nan=True)


# ---- ensemble mode across ranks: reduce-scatter of the plane sums over the window axis, peak fit per slice, all-gather ----------
def _ens_partial(O, frames, ws, ov, corr_min, s2n_min):
    """One rank's share of `_get_ffpiv_mean` (ffpiv.py:200-243, :359-365) with the oracle: plane sums, counts, per-pair stats."""
    nr, nc = O.get_array_shape(frames.shape[-2:], ws, ov)
    ens = O.Ensemble(nr, nc, ws, ov, corr_min=corr_min, s2n_min=s2n_min, count_min=0.0)
    ens.add_chunk(frames)
    return ens


def _ens_worker(rank, world, port, n_pairs, shape, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ffpiv_oracle as O
    from pyorc_b200 import synth

    O.CLIP_NORMALIZED = False
    ws, ov = (16, 16), (8, 8)
    frames = synth.particle_frames(n_pairs + 1, *shape, dtype=np.uint8)
    nr, nc = O.get_array_shape(shape, ws, ov)
    nw, npx = nr * nc, ws[0] * ws[1]
    table = parallel.scatter_pair_table(n_pairs)
    a, b = int(table[rank, 0]), int(table[rank, 1])
    f0, f1 = parallel.frame_range((a, b))
    if b > a:
        ens = _ens_partial(O, frames[f0:f1], ws, ov, 0.2, 1.5)
        plane = torch.from_numpy(np.array(ens.corr_sum, dtype=np.float32).reshape(nw, npx).copy())
        count = torch.from_numpy(np.asarray(ens.corr_count, dtype=np.float32).reshape(nw).copy())
        local = torch.from_numpy(np.stack([ens.corr_chunks[0], ens.s2n_chunks[0]]).reshape(2, b - a, nr, nc))
    else:
        plane, count = torch.zeros((nw, npx)), torch.zeros(nw)
        local = torch.zeros((2, 0, nr, nc))
    n_chunks = int((table[:, 1] > table[:, 0]).sum())
    min_count = 0.2 * n_chunks

    def finish(first, n):   # count filter + mean plane + peak fit of a window slice (ffpiv.py:280-282, :324)
        cs = plane[first : first + n].numpy().reshape(n, ws[0], ws[1]).copy()
        cc = count[first : first + n].numpy()
        with np.errstate(all="ignore"):
            cs[cc < min_count] = np.nan
            mean = np.divide(cs, cc.astype(np.int64)[:, None, None])   # float32 / int64 -> float64, as in the reference
        u, v = O.u_v_displacement(mean[None], 1, n)
        return torch.from_numpy(u.reshape(n).astype(np.float32)), torch.from_numpy(v.reshape(n).astype(np.float32))

    u, v, cnt = parallel.ensemble_reduce_finish(plane, count, finish)
    stats = parallel.gather_fields(local, n_pairs, table)
    cm, sn = parallel.aggregate_ensemble(stats[0].reshape(n_pairs, -1).numpy(), stats[1].reshape(n_pairs, -1).numpy(), cnt.numpy(), min_count, nr, nc)
    np.savez(os.path.join(out_dir, f"ens{rank}.npz"), u=u.numpy().reshape(1, nr, nc), v=v.numpy().reshape(1, nr, nc), cm=cm, sn=sn, cnt=cnt.numpy(),
             table=table)
    dist.destroy_process_group()


def test_window_slices():
    for nw in (0, 1, 7, 616, 1888):
        for w in (1, 2, 3, 8):
            t = parallel.window_slices(nw, w)
            assert t[0, 0] == 0 and t[-1, 1] == nw and np.all(t[1:, 0] == t[:-1, 1])
            assert np.all((t[:-1, 1] - t[:-1, 0]) == nw // w)


@pytest.mark.parametrize("n_pairs,shape", [(5, (40, 56)), (4, (48, 64)), (1, (40, 56))])   # odd window counts, equal shards, an empty shard
def test_ensemble_over_gloo_equals_single_process(tmp_path, n_pairs, shape):
    world = 2
    port = _free_port()
    mp.spawn(_ens_worker, args=(world, port, n_pairs, shape, str(tmp_path)), nprocs=world, join=True)
    from oracle import ffpiv_oracle as O
    from pyorc_b200 import synth

    O.CLIP_NORMALIZED = False
    ws, ov = (16, 16), (8, 8)
    frames = synth.particle_frames(n_pairs + 1, *shape, dtype=np.uint8)
    nr, nc = O.get_array_shape(shape, ws, ov)
    got0 = np.load(tmp_path / "ens0.npz")
    table = got0["table"]
    # single process, the reference's chunk loop with the ranks' pair ranges as chunks (same sums in the same order)
    ref = O.Ensemble(nr, nc, ws, ov, corr_min=0.2, s2n_min=1.5, count_min=0.2)
    for a, b in table:
        if b > a:
            ref.add_chunk(frames[a : b + 1])
    u, v, cm, sn = ref.finalize()
    assert np.isfinite(u).sum() > 0
    for r in range(world):
        got = np.load(tmp_path / f"ens{r}.npz")
        assert np.array_equal(got["u"], u.astype(np.float32), equal_nan=True)
        assert np.array_equal(got["v"], v.astype(np.float32), equal_nan=True)
        assert np.array_equal(got["cm"], cm, equal_nan=True) and np.array_equal(got["sn"], sn, equal_nan=True)
        assert np.array_equal(got["cnt"], np.asarray(ref.corr_count).reshape(-1))
