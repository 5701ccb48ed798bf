"""Samples of an .ncu-rep source page grouped into the code between block-wide barriers, with the stall mix and
the hottest instructions of the segments executed most often (the recurrence loop).
    python scripts/ncu_segments.py rep.ncu-rep regex:mega_kernel [min_exec]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
min_exec = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; col = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break
        continue
    data.append(r)
S = [int(r[col["# Samples"]]) for r in data]
tot = sum(S)
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
seg_start = 0; acc = 0; segs = []
for i, r in enumerate(data):
    acc += S[i]
    if "BAR.SYNC" in r[col["Source"]] or "EXIT" in r[col["Source"]]:
        if acc > 20:
            print(f"{seg_start:5d}-{i:5d} samples {acc:5d} {100 * acc / tot:5.1f}%  exec {r[col['Instructions Executed']]}")
            segs.append((seg_start, i, int(r[col['Instructions Executed']])))
        seg_start = i + 1; acc = 0
print("total samples", tot)
for lo, hi, ex in segs:
    if ex < min_exec:
        continue
    agg = {n: 0 for n in stall_cols}; t = 0
    for r in data[lo:hi + 1]:
        t += int(r[col["# Samples"]])
        for n in stall_cols:
            agg[n] += int(r[col[n]] or 0)
    print(lo, hi, t, {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for i in sorted(sorted(range(lo, hi + 1), key=lambda i: -S[i])[:16]):
        r = data[i]
        st = {n[6:]: int(r[col[n]] or 0) for n in stall_cols if int(r[col[n]] or 0) > 0}
        print(f"   {i:5d} {S[i]:5d} {r[col['Source']].strip()[:64]:64s} {sorted(st.items(), key=lambda kv: -kv[1])[:3]}")
