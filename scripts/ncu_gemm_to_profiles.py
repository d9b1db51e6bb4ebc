"""Turn the `ncu --set full` capture of scripts/gemm_prof.py (read here, no GPU needed) into the committed evidence:
profiles/<tag>_gemm_tcp_ncu_full_extract.csv (headline metrics per launch) and profiles/<rN>_gemm_tcp_ncu.json (DRAM bytes
per launch, keyed by the GEMM shape string bench.py uses -- bench.py reads `traffic` from it).
usage: ncu_gemm_to_profiles.py report.ncu-rep tag"""
import csv, io, json, os, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
keys = ["nt 40960x128x400", "nt 40960x400x128", "tn 128x400x40960 acc", "nn 40960x400x384"]       # scripts/gemm_prof.py order
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'smsp__inst_executed.sum']
idx = {w: hdr.index(w) for w in want if w in hdr}
out = [["shape"] + [w + (" [" + units[i] + "]" if units[i] else "") for w, i in idx.items()]]
js = {}
def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
for k, r in zip(keys, data):
    out.append([k] + [r[i] for i in idx.values()])
    ir, iw = idx['dram__bytes_read.sum'], idx['dram__bytes_write.sum']
    js[k] = dict(dram_bytes=int(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])),
                 dram_read_bytes=int(to_bytes(r[ir], units[ir])), dram_write_bytes=int(to_bytes(r[iw], units[iw])),
                 ncu_duration_us=float(r[idx['gpu__time_duration.sum']]), capture=tag,
                 tensor_pipe_pct=float(r[idx['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]),
                 dram_pct=float(r[idx['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]))
with open(os.path.join(ROOT, "profiles", tag + "_gemm_tcp_ncu_full_extract.csv"), "w", newline="") as f:
    csv.writer(f).writerows(zip(*out))          # one column per launch
json.dump(js, open(os.path.join(ROOT, "profiles", tag[:2] + "_gemm_tcp_ncu.json"), "w"), indent=1)
print(json.dumps(js, indent=1))
