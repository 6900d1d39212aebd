"""One 128x128 / 50 % launch on 8K frames (development aid for ncu captures of piv_rows128_kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    run(4320, 7680, (128, 128), (64, 64), n, reps=int(os.environ.get("B2_REPS", "3")))
