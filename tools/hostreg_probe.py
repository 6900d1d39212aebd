"""What does pinning pyorc's numpy memory IN PLACE cost?  cudaHostRegister / cudaHostUnregister of a pageable numpy array (the
alternative to staging it through a page-locked ring: one crossing of the host memory bus instead of three) - development aid."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

rt = torch.cuda.cudart()
torch.cuda.init()
dev = torch.device("cuda", 0)
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists("/sys/kernel/mm/transparent_hugepage/enabled") else "?", flush=True)

def reg(ptr, n, flags=0):
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(ptr, n, flags)
    return int(rc), (time.perf_counter() - t0) * 1e3

def unreg(ptr):
    t0 = time.perf_counter()
    rc = rt.cudaHostUnregister(ptr)
    return int(rc), (time.perf_counter() - t0) * 1e3

for mb in (4, 16, 64, 209):
    n = mb << 20
    for trial in range(3):
        a = np.empty(n, np.uint8)
        a[:] = 7                                  # touched: the pages exist
        ptr = a.ctypes.data
        rc, t_reg = reg(ptr, n)
        d = torch.empty(n, dtype=torch.uint8, device=dev)
        h = torch.from_numpy(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
        t_h2d = e0.elapsed_time(e1)
        rc2, t_un = unreg(ptr)
        rc3, t_reg2 = reg(ptr, n)               # the same pages again
        rc4, t_un2 = unreg(ptr)
        rc5, t_ro = reg(ptr, n, 8)              # cudaHostRegisterReadOnly
        rc6, _ = unreg(ptr) if rc5 == 0 else (0, 0)
        print(f"{mb:4d} MB trial {trial}: register {t_reg:7.3f} ms ({n / t_reg / 1e6:6.1f} GB/s) rc {rc} | H2D {t_h2d:6.3f} ms ({n / t_h2d / 1e6:5.1f} GB/s) | unregister {t_un:6.3f} ms | "
              f"again: register {t_reg2:7.3f} unregister {t_un2:6.3f} | read-only flag: {t_ro:7.3f} rc {rc5}", flush=True)
        del a, d, h

# four threads registering four quarters of one 209 MB array at the same time
n = 209 << 20
a = np.empty(n, np.uint8); a[:] = 3
q = (n // 4) & ~4095
base = (a.ctypes.data + 4095) & ~4095
times = [0.0] * 4
def work(k):
    times[k] = reg(base + k * q, q - 4096)[1]
t0 = time.perf_counter()
ths = [threading.Thread(target=work, args=(k,)) for k in range(4)]
[t.start() for t in ths]; [t.join() for t in ths]
wall = (time.perf_counter() - t0) * 1e3
print(f"4 threads x {q >> 20} MB: wall {wall:.3f} ms, per thread {['%.2f' % t for t in times]}", flush=True)
for k in range(4):
    unreg(base + k * q)
