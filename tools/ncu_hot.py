"""Top stall sites of a kernel from an .ncu-rep (source page, SASS view): python tools/ncu_hot.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    try:
        s = int(r[isamp])
    except Exception:
        continue
    tot += s
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stall), reverse=True)[:2]
    data.append((s, r[ia], r[isrc], top))
print(f"# {rows[0][1][:120]}  total samples {tot}")
for idx, (s, a, src, top) in enumerate(data):
    pass
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:N]
for i in sorted(order):
    s, a, src, top = data[i]
    print(f"{100.0*s/tot:5.1f}%  {src[:70]:70s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
