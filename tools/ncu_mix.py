"""Dynamic instruction mix (warp-level executed instructions by opcode) and stall samples by opcode from an .ncu-rep
captured with --import-source on:  python tools/ncu_mix.py rep [n]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:]))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); sm = collections.Counter()
for r in rows[1:]:
    src = r[ix["Source"]].strip().split()
    if not src:
        continue
    op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
    op = op.split(".")[0].rstrip(";")
    ex[op] += int(r[ix["Instructions Executed"]] or 0)
    sm[op] += int(r[ix["# Samples"]] or 0)
te, ts = sum(ex.values()), sum(sm.values())
print(f"executed warp instructions {te}, samples {ts}")
for op, c in ex.most_common(n):
    print(f"  {op:10s} {c:12d} {100*c/te:5.1f}% of instructions   {100*sm[op]/max(ts,1):5.1f}% of stall samples")
