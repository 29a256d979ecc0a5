"""Top stall locations of one kernel from an ncu report's source page (SASS + sampled stalls)."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break  # first kernel instance only
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
S = "Warp Stall Sampling (All Samples)"
tot = sum(int(r[ix[S]] or 0) for r in data)
toti = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print(f"kernel {kern}: {len(data)} SASS lines, {toti} warp instructions, {tot} stall samples")
stall_cols = [h for h in hdr if h.startswith("stall_")]
agg = collections.Counter()
for r in data:
    for h in stall_cols:
        agg[h] += int(r[ix[h]] or 0)
print("stall reasons:", ", ".join(f"{k[6:]}={v}" for k, v in agg.most_common(8)))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix[S]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {100 * int(r[ix[S]]) / tot:5.1f}% exec={int(r[ix['Instructions Executed']]):9d} {r[ix['Source']].strip()[:70]:70s} {why}")
