"""33 .. 64 px windows: padded 128-plane polyphase kernel (auto) against the shared-memory kernel (variant 1) - development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
if __name__ == "__main__":
    for ws, ov in (((50, 50), (25, 25)), ((40, 40), (20, 20)), ((64, 48), (32, 24)), ((34, 34), (17, 17))):
        for variant in (0, 1):
            run(1080, 1920, ws, ov, 21, variant=variant)
    run(2160, 3840, (50, 50), (25, 25), 21, variant=0)
    for variant in (0, 1):   # float32 frames at 128x128: polyphase kernel against the shared-memory kernel
        run(2160, 3840, (128, 128), (64, 64), 11, dtype="float32", variant=variant)
