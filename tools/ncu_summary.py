"""Summarise an .ncu-rep: headline metrics + stall samples by code region (developer tool)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'sm__inst_executed.sum', 'launch__grid_size',
        'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__block_size',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warps_issue_stalled', 'sm__throughput.avg.pct',
        'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct', 'smsp__inst_executed.avg.per_cycle_active',
        'smsp__issue_active.avg.pct', 'launch__shared_mem_per_block', 'sm__maximum_warps_per_active_cycle_pct',
        'sm__inst_executed_pipe_lsu', 'smsp__inst_executed_pipe']
for r in rows[2:]:
    print("== kernel:", r[hdr.index('Kernel Name')][:90])
    for i, h in enumerate(hdr):
        if any(h.startswith(w) for w in want) and not h.endswith('.pct') :
            try:
                v = float(r[i])
            except ValueError:
                continue
            if 'stalled' in h and v < 0.3:
                continue
            print(f"  {h} [{units[i]}] = {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if his:
    hi = his[0]
    h = rows[hi]
    ia, isamp, iex = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
    end = his[1] if len(his) > 1 else len(rows)
    data = []
    for r in rows[hi + 1:end]:
        try:
            data.append((int(r[isamp] or 0), int(r[iex] or 0), r[ia]))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    print("total samples", tot, "instructions", len(data))
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    for s in range(0, len(data), step):
        blk = data[s:s + step]
        ops = {}
        for d in blk:
            t = d[2].split()
            op = t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')
            ops[op] = ops.get(op, 0) + 1
        keyops = {k: v for k, v in ops.items() if any(x in k for x in ('BAR', 'MATCH', 'LDG', 'STG', 'ATOM', 'LDS', 'STS', 'WARPSYNC', 'SHFL', 'VOTE', 'RED', 'LDL', 'STL'))}
        print(f"  [{s:4d}] {100*sum(d[0] for d in blk)/tot:5.1f}%  exec/warp-item~{sum(d[1] for d in blk)}  {keyops}")
    print("top instructions:")
    for i, d in sorted(enumerate(data), key=lambda x: -x[1][0])[:25]:
        print(f"  {i:4d} {100*d[0]/tot:5.1f}%  {d[2][:100]}")
