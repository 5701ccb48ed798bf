"""Top stall contributors of one kernel from an .ncu-rep source page (SASS view).
    python scripts/ncu_source_top.py rep.ncu-rep regex:k2_merge [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# find header row
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
col = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break
        continue
    data.append(r)
tot = sum(int(r[col["# Samples"]]) for r in data)
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[col[n]] or 0) for r in data) for n in stall_cols}
print("total samples", tot, "instructions", len(data))
print("stall mix:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
# opcode mix by executed instructions
ops = {}
for r in data:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[col["Instructions Executed"]] or 0)
ti = sum(ops.values())
print("opcode mix (warp instrs):", {k: round(100 * v / ti, 1) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]}, "total", ti)
print("--- top instructions by samples")
for idx, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][col["# Samples"]]))[:topn]:
    st = {n[6:]: int(r[col[n]] or 0) for n in stall_cols if int(r[col[n]] or 0) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{idx:5d} {int(r[col['# Samples']]):6d} {r[col['Source']].strip()[:70]:70s} {top}")
