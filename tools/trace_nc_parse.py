"""Summarise a neg_cent role trace (MAS_NC_DEBUG=16 MAS_NC_TRACE_DUMP=1 output of tools/trace_nc.py)."""
import sys, collections, statistics
ev = collections.defaultdict(list)
for line in open(sys.argv[1]):
    f = line.split()
    if len(f) == 9 and f[0] == "TRACE":
        ev[int(f[2])].append((int(f[4]), int(f[6]), int(f[8])))
names = {0: "B loader", 1: "MMA issuer", 2: "epilogue", 3: "converter group 0"}
for role in sorted(ev):
    e = ev[role]
    gaps = collections.defaultdict(list)
    for (a, ai, at), (b, bi, bt) in zip(e, e[1:]):
        gaps[(a, b)].append(bt - at)
    print(f"{names.get(role, role)}: {len(e)} events, span {e[-1][2] - e[0][2]} cycles")
    for k in sorted(gaps):
        v = gaps[k]
        print(f"   ev{k[0]}->ev{k[1]}: n={len(v)} median {statistics.median(v):.0f} mean {statistics.mean(v):.0f} max {max(v)}")
