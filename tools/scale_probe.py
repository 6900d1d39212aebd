"""Where does the multi-GPU step lose time against the kernel alone?  torchrun --nproc-per-node N tools/scale_probe.py
Times 40 steps each of: kernel alone | + P2P stores to every rank (no barrier) | + completion barrier on the consumer stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pyorc_b200 import parallel, synth
from pyorc_b200.engine import Engine

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(lr)
H, W, WS, OV, P = 1080, 1920, (64, 64), (32, 32), 100
fr = synth.particle_frames_torch(P + 1, H, W, dev, dtype="uint8", seed=synth.SEED + rank)
eng.plan((H, W), WS, OV, np.uint8)
table = parallel.shard_pairs(P * world, world)
pg = parallel.PeerGather(eng, P * world, table, mode="fused")
pp = parallel.PeerGather(eng, P * world, table, mode="push")

def timed(step, n=40):
    for _ in range(5):
        step()
    torch.cuda.synchronize(); pg.drain(); pp.drain(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        step()
    b.record()
    torch.cuda.synchronize(); pg.drain(); pp.drain()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

def alone():
    eng.set_peer_outputs(None, 1, 0)
    eng.pairs(fr, WS, OV)
def stores_only():
    pg.begin(); eng.pairs(fr, WS, OV)
def full():
    pg.begin(); eng.pairs(fr, WS, OV); pg.end()
def push():
    eng.set_peer_outputs(None, 1, 0)
    pp.begin(); res = eng.pairs(fr, WS, OV); pp.end(res)
for name, fn in (("kernel alone", alone), ("fused: + P2P stores", stores_only), ("fused: + barrier (consumer)", full), ("push on the consumer stream", push), ("kernel alone", alone),
                 ("fused: + barrier (consumer)", full), ("push on the consumer stream", push)):
    ms = timed(fn)
    if rank == 0:
        print(f"N={world} {name:28s} {ms:.4f} ms per step", flush=True)
pp.drain(); pg.close(); eng.close(); dist.barrier(); dist.destroy_process_group()
