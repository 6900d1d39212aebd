"""Warp-state samples of an ncu report binned along the kernel's instruction stream (250 SASS instructions per bin), with the
dominant opcodes per bin - shows which phase of a long straight-line kernel the time goes to.  usage: ncu_phases.py rep [bin]"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
B = int(sys.argv[2]) if len(sys.argv) > 2 else 250
rows = list(csv.reader(io.StringIO(out)))
print("#", rows[0][1][:120])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, name):
    try:
        return float(r[ix[name]])
    except Exception:
        return 0.0
def op(r):
    t = r[ix["Source"]].split()
    if not t:
        return ""
    return (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
tot = sum(f(r, "# Samples") for r in data)
stalls = ["stall_selected", "stall_not_selected", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_barrier", "stall_mio", "stall_dispatch", "stall_math", "stall_long_sb"]
print(f"# {len(data)} SASS instructions, {int(tot)} samples; per bin: share of all samples, warp-instructions executed (M), % of the bin's samples per warp state, top opcodes")
print("bin,instr_from,instr_to,samples_pct,exec_M," + ",".join(s.replace("stall_", "") for s in stalls) + ",top_ops")
for b in range(0, len(data), B):
    seg = data[b:b + B]
    smp = sum(f(r, "# Samples") for r in seg)
    ex = sum(f(r, "Instructions Executed") for r in seg)
    ops = collections.Counter(op(r) for r in seg)
    st = [sum(f(r, s) for r in seg) for s in stalls]
    print(f"{b // B},{b},{b + len(seg)},{100 * smp / max(tot, 1):.2f},{ex / 1e6:.1f}," + ",".join(f"{100 * x / max(smp, 1):.1f}" for x in st) + "," + " ".join(f"{k}:{v}" for k, v in ops.most_common(4)))
