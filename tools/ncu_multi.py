"""Per-kernel summary (+ stall hot spots) of an ncu report with several kernels.  usage: ncu_multi.py rep [pattern] [topN]"""
import csv, subprocess, io, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ''; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active']
for v in rows[2:]:
    name = v[h.index('Kernel Name')]
    if pat not in name: continue
    print('==', name)
    for k in keys:
        if k in h: print('  ', k, v[h.index(k)], u[h.index(k)])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; secs.append(cur)
    elif cur is not None: cur['rows'].append(r)
for sec in secs:
    if pat not in sec['name'] or not sec['rows']: continue
    h = sec['rows'][0]; idx = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
    data = []
    for r in sec['rows'][1:]:
        if len(r) < len(h): continue
        data.append((int(r[idx['# Samples']] or 0), r[idx['Source']].strip(), int(r[idx['Instructions Executed']] or 0), {k[6:]: int(r[idx[k]] or 0) for k in stalls}))
    tot = {}
    for d in data:
        for k, v in d[3].items(): tot[k] = tot.get(k, 0) + v
    print('--', sec['name'], 'samples', sum(d[0] for d in data), {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
    for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][0])[:topn]):
        s, srcl, n, st = data[i]
        print(f"{i:5d} {s:5d} {n:8d} {srcl[:66]:66s}", {k: v for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v})
