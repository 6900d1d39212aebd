"""Throughput of the window sizes between the FFT sizes (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import quick_bench as q
for n, nf in ((50, 5), (66, 3), (80, 3), (100, 3), (126, 3)):
    q.run(1080, 1920, (n, n), (n // 2, n // 2), nf, reps=3)
q.run_ens(1080, 1920, (100, 100), (50, 50), 3, reps=2)
