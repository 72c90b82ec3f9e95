"""Summarise an ncu report's source page: stall totals + hottest SASS lines.  usage: ncu_src.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
for k in ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']:
    if k in h: print(k, v[h.index(k)], u[h.index(k)])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; idx = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
tot = {k: 0 for k in stalls}; data = []
for r in rows[2:]:
    if len(r) < len(h): continue
    s = int(r[idx['# Samples']] or 0)
    st = {k: int(r[idx[k]] or 0) for k in stalls}
    for k in stalls: tot[k] += st[k]
    data.append((s, r[idx['Source']].strip(), int(r[idx['Instructions Executed']] or 0), st))
print('total samples', sum(d[0] for d in data))
print({k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
top = sorted(range(len(data)), key=lambda i: -data[i][0])[:topn]
for i in sorted(top):
    s, src, n, st = data[i]
    print(f"{i:5d} {s:5d} {n:8d} {src[:64]:64s}", {k[6:]: v for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v})
