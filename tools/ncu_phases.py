"""Stall samples of one kernel aggregated by barrier-delimited phase (ncu source page)."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
S = "Warp Stall Sampling (All Samples)"
cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix[S]] or 0) for r in data)
toti = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print(f"{kern}: {len(data)} SASS, {toti} warp instr, {tot} samples")
cur = [0, 0, 0, collections.Counter()]
for i, r in enumerate(data):
    src = r[ix["Source"]].strip()
    cur[0] += int(r[ix[S]] or 0)
    cur[1] += int(r[ix["Instructions Executed"]] or 0)
    cur[2] += 1
    for h in cols:
        cur[3][h[6:]] += int(r[ix[h]] or 0)
    if "BAR.SYNC" in src or i == len(data) - 1:
        t = sum(cur[3].values()) or 1
        if cur[0] > 0.005 * tot:
            print(f"  ..{i:5d} samples {100 * cur[0] / tot:5.1f}%  instr {100 * cur[1] / toti:5.1f}%  sass {cur[2]:4d}  "
                  + ", ".join(f"{k}={100 * v / t:.0f}%" for k, v in cur[3].most_common(5)))
        cur = [0, 0, 0, collections.Counter()]
