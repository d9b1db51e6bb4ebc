"""Summarise an .ncu-rep (read in the build container): per-kernel headline metrics and the top stall sites.
usage: ncu_extract.py report.ncu-rep [kernel-index-for-source-page]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__average_t_sector_hit_rate_realtime.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'l1tex__lsu_writeback_active_mem_lg.sum','sm__memory_throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    idx = [i for i, h in enumerate(hdr) if h == w]
    if idx:
        i = idx[0]
        print("%-72s %-10s %s" % (w, units[i], [r[i] for r in data]))
if len(sys.argv) > 2:
    k = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    print(rows[0][:2], len(data), "rows")
    iS, iSrc, iA, iE = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Address'), hdr.index('Instructions Executed')
    stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    I = lambda x: int(x) if x.strip().lstrip('-').isdigit() else 0
    seen, uniq = set(), []
    for r in data:
        if r[iA] not in seen:
            seen.add(r[iA]); uniq.append(r)
    tot = sum(I(r[iS]) for r in uniq)
    print("total samples", tot)
    agg = {}
    for r in uniq:
        for i in stalls:
            agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + I(r[i])
    print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for r in sorted(uniq, key=lambda r: -I(r[iS]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
        st = {hdr[i][6:]: I(r[i]) for i in stalls if I(r[i]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(r[iA][-5:], "%5s %8s" % (r[iS], r[iE]), r[iSrc][:84], st)
