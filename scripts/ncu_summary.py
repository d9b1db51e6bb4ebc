"""Headline counters of every launch in an `ncu --set full` report, as a markdown table (read here, no GPU needed).
usage: ncu_summary.py report.ncu-rep > profiles/<name>.md"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [('Kernel Name', 'kernel'), ('Grid Size', 'grid'), ('Block Size', 'block'), ('gpu__time_duration.sum', 'time'),
        ('launch__registers_per_thread', 'regs'), ('sm__cycles_active.avg', 'SM cycles active'), ('sm__cycles_elapsed.max', 'cycles elapsed'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % of active'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % of peak'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 % of peak'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
        ('smsp__inst_executed.sum', 'warp instructions')]
cols = [(hdr.index(k), n) for k, n in want if k in hdr]
print("| " + " | ".join(n for _, n in cols) + " |")
print("|" + "---|" * len(cols))
for r in data:
    cells = []
    for i, n in cols:
        v = r[i]
        if n == 'kernel':
            v = v.split('(')[0][:48]
        elif units[i]:
            v = "%s %s" % (v, units[i])
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
