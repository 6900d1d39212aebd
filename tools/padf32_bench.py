"""float32 frames with pyorc's non power-of-two windows (its recipe: normalize -> edge_detect -> minmax -> get_piv(window_size=25)):
padded mode of the row-per-thread kernel (auto) against the shared-memory kernel (variant 1) and the uint8 padded mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
if __name__ == "__main__":
    for ws, ov in (((26, 26), (12, 12)), ((20, 20), (10, 10)), ((10, 10), (5, 5)), ((30, 18), (15, 9))):
        for variant in (0, 1):
            run(1080, 1920, ws, ov, 21, dtype="float32", variant=variant)
        run(1080, 1920, ws, ov, 21, dtype="uint8", variant=0)
