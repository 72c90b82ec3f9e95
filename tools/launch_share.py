"""Kernel shares of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).  usage: launch_share.py launches.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] == "ns" else v * 1e3 if r[iu] == "ms" else v
    n, t = tot.get(r[ik], (0, 0.0))
    tot[r[ik]] = (n + 1, t + v)
allt = sum(t for _, t in tot.values())
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:100]:100s} launches {n:4d}  total {t:9.1f} us  mean {t / n:8.2f} us  share {100 * t / allt:5.1f} %")
mas = {k: v for k, v in tot.items() if "mas::" in k}
mt = sum(t for _, t in mas.values())
print("\nshare inside the maximum_path chain:")
for k, (n, t) in sorted(mas.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k[:60]:60s} {100 * t / mt:5.1f} %   mean {t / n:8.2f} us")
