"""Summarise an `ncu --csv --metrics ...` launch list: per kernel time, DRAM bytes, L2 hit rate."""
import csv, collections, sys
f = sys.argv[1]
n = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(open(f, errors="replace")))
hdr = None
data = collections.OrderedDict()
for r in rows:
    if len(r) > 10 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        key = (int(d["ID"]), d["Kernel Name"].split("(")[0][-60:])
        try:
            data.setdefault(key, {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            pass
tot = sum(v.get("gpu__time_duration.sum", 0) for v in data.values())
print(f"{'id':>4} {'kernel':<50} {'ms':>9} {'share':>6} {'rdGB':>8} {'wrGB':>8} {'B/q rd':>7} {'B/q wr':>7} {'GB/s':>8} {'L2hit':>6}")
for (i, name), v in data.items():
    t = v.get("gpu__time_duration.sum", 0) / 1e6
    rd, wr = v.get("dram__bytes_read.sum", 0), v.get("dram__bytes_write.sum", 0)
    bq = (f"{rd / n:7.1f} {wr / n:7.1f}" if n else f"{'':7} {'':7}")
    print(f"{i:>4} {name:<50} {t:9.3f} {100 * t * 1e6 / tot:5.1f}% {rd / 1e9:8.3f} {wr / 1e9:8.3f} {bq} {(rd + wr) / t / 1e6 if t else 0:8.1f} {v.get('lts__t_sector_hit_rate.pct', 0):6.1f}")
print(f"total {tot / 1e6:.3f} ms")
