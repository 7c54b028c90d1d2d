"""Top SASS instructions by stall samples from `ncu --page source --csv` (python tools/ncu_hot.py rep [n])."""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:]))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {100*s/tot:5.1f}% exec={r[ix['Instructions Executed']]:>10s} {r[ix['Source']].strip()[:90]:90s} {st}")
