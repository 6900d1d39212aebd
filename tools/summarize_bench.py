"""Markdown table of the committed bench lines (profiles/rNN/bench_n{1,2,4,8}_final.json) - the numbers DESIGN.md quotes."""
import json, os, sys
d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02")
rows = {}
for n in (1, 2, 4, 8):
    p = os.path.join(d, f"bench_n{n}_final.json")
    if os.path.exists(p):
        rows[n] = json.load(open(p))
ns = sorted(rows)
def cell(fn):
    out = []
    for n in ns:
        try:
            out.append(fn(rows[n]))
        except Exception:
            out.append("-")
    return " | ".join(out)
def shard(r, k, key="windows_per_s"):
    s = [x for x in r["sharded_configs"] if x["config"].startswith(f"configs[{k}]")][0]
    return s
v1 = rows[1]["value"] if 1 in rows else None
print("| | " + " | ".join(f"N = {n}" for n in ns) + " |")
print("|---|" + "---|" * len(ns))
print("| `value`, M windows/s (ms per step) | " + cell(lambda r: f"{r['value'] / 1e6:.1f} ({r['ms_per_step']:.3f})") + " |")
print("| weak-scaling efficiency of `value` | " + cell(lambda r: f"{r['value'] / (r['n_gpus'] * v1):.3f}") + " |")
print("| kernel alone, ms per launch | " + cell(lambda r: f"{r['roofline']['ms_per_launch']:.3f}") + " |")
print("| `e2e` get_b2piv (pageable) + gather, ms per step (M windows/s) | " + cell(lambda r: f"{r['e2e']['ms_per_step']:.2f} ({r['e2e']['value'] / 1e6:.1f})") + " |")
print("| page-locked `Engine.pairs`, ms | " + cell(lambda r: f"{r['e2e']['pinned_engine_pairs']['ms_per_step']:.2f}") + " |")
print("| concurrent H2D floor, ms slowest rank (GB/s per GPU) | " + cell(lambda r: f"{r['e2e']['pcie']['h2d_only_ms_slowest_rank']:.2f} ({r['e2e']['pcie']['h2d_gbs_per_gpu_slowest']:.1f})") + " |")
print("| `configs[3]` 4K shard, 500 pairs / GPU, M windows/s | " + cell(lambda r: f"{shard(r, 3)['windows_per_s'] / 1e6:.1f}" + (" (in full)" if shard(r, 3)["is_baseline_config_in_full"] else "")) + " |")
print("| `configs[4]` 8K shard, 625 pairs / GPU, M windows/s | " + cell(lambda r: f"{shard(r, 4)['windows_per_s'] / 1e6:.2f}" + (" (in full)" if shard(r, 4)["is_baseline_config_in_full"] else "")) + " |")
print("| ensemble over the ranks = one GPU (max abs du, dv px) | " + cell(lambda r: f"{r['ensemble_multi_gpu']['equals_single_gpu']} ({r['ensemble_multi_gpu']['max_abs_du_px']:.1e}, {r['ensemble_multi_gpu']['max_abs_dv_px']:.1e})") + " |")
print("| ensemble over the ranks, ms (M windows/s) | " + cell(lambda r: f"{r['ensemble_multi_gpu']['ms']:.2f} ({r['ensemble_multi_gpu']['windows_per_s'] / 1e6:.0f})") + " |")
if 1 in rows:
    r = rows[1]
    print(f"\nN = 1: fp32 {r['fp32']['achieved']:.1f} TFLOP/s of {r['fp32']['peak_measured']:.1f} measured = {r['fp32']['frac']:.3f}; roofline.frac {r['roofline']['frac']:.4f}; cpu_baseline {r['cpu_baseline']['value']:.0f} windows/s on {r['cpu_baseline']['cores']} cores; rmse u {r['rmse_vs_oracle']['u_px']:.2e} v {r['rmse_vs_oracle']['v_px']:.2e} px")
    for o in r["other_configs"]:
        print(f"  {o['config']}: {o.get('windows_per_s', 0) / 1e6:.1f} M windows/s" + (f", rmse vs imposed field {o['rmse_vs_imposed_field_px']:.4f} px" if "rmse_vs_imposed_field_px" in o else ""))
