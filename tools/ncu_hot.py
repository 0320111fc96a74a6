"""Top stalled SASS instructions of one kernel launch in an .ncu-rep (source page).
    python tools/ncu_hot.py report.ncu-rep <launch-index> [top-n]"""
import csv, subprocess, sys
rep, idx = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", idx, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[ix['# Samples']].isdigit():
        data.append(r)
print(rows[0][:2])
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for n, r in enumerate(data):
    r.append(n)
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:topn]:
    s = int(r[ix['# Samples']])
    st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{r[-1]:5d} {s:6d} {100*s/max(tot,1):5.1f}%  {r[ix['Source']].strip()[:72]:72s} {st}")
