"""float32 frames, even windows of 34 .. 64 px: padded 128-plane polyphase kernel (auto) against the shared-memory kernel
(variant 1), per time step and ensemble - development aid (profiles/r02/pad128_f32_bench.log)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run, run_ens
if __name__ == "__main__":
    for ws, ov in (((50, 50), (25, 25)), ((40, 40), (20, 20)), ((64, 48), (32, 24)), ((34, 34), (17, 17))):
        for variant in (0, 1):
            run(1080, 1920, ws, ov, 11, dtype="float32", variant=variant)
    run(1080, 1920, (50, 50), (25, 25), 11, dtype="uint8", variant=0)      # the uint8 padded mode beside it
    for variant in (0, 1):
        run_ens(1080, 1920, (50, 50), (25, 25), 11, dtype="float32", variant=variant)
