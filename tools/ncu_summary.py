"""summarise an .ncu-rep (first kernel): key metrics + SASS opcode mix + top stall lines. usage: ncu_summary.py rep [npx]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; npx = float(sys.argv[2]) if len(sys.argv) > 2 else 805306368.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_barrier.pct','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
for i, h in enumerate(hdr):
    if h in keep or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
        print(f"{h:90s} {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); h = rows[1]
ia, ie, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
c = collections.Counter(); samp = collections.Counter(); tot = 0; lines = []
for r in rows[2:]:
    try: n = int(r[ie])
    except Exception: continue
    t = r[ia].strip().split()
    if not t: continue
    op = (t[1] if t[0].startswith('@') else t[0]).rstrip(';')
    c[op] += n; tot += n
    try: s = int(r[isamp])
    except Exception: s = 0
    samp[op] += s; lines.append((s, n, r[ia].strip()[:90]))
print(f"\nwarp instructions executed: {tot}  = {tot*32/npx:.1f} thread-instructions per input pixel")
for op, n in c.most_common(32): print(f"  {op:24s} {n:12d} {100*n/tot:5.1f}%  {n*32/npx:6.2f}/px  samples {samp[op]}")
print("\ntop sampled SASS lines:")
for s, n, t in sorted(lines, reverse=True)[:25]: print(f"  {s:7d} {n:10d}  {t}")
