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
