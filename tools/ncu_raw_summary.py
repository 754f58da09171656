"""Key metrics of an `ncu -i X.ncu-rep --page raw --csv` dump, one per line (name, unit, value):
    ncu -i prof.ncu-rep --page raw --csv > raw.csv;  python tools/ncu_raw_summary.py raw.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = re.compile(r'^(dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum$|launch__(block_size|grid_size|registers_per_thread|'
                  r'shared_mem_per_block_dynamic|occupancy_limit_(registers|shared_mem))$|l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.sum$|'
                  r'l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|'
                  r'l1tex__t_sector_hit_rate\.pct$|lts__t_sector_hit_rate\.pct$|lts__t_bytes\.sum$|'
                  r'sm__inst_executed_pipe_(fp64|alu|lsu|fma)\.avg\.pct_of_peak_sustained_active$|'
                  r'sm__inst_executed_pipe_tensor.*avg\.pct_of_peak_sustained_active$|sm__pipe_tensor.*(avg|sum)\.pct_of_peak_sustained_(active|elapsed)$|'
                  r'sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active$|sm__throughput\.avg\.pct_of_peak_sustained_elapsed$|'
                  r'sm__warps_active\.avg\.pct_of_peak_sustained_active$|smsp__cycles_active\.avg$|smsp__inst_executed\.sum$|'
                  r'smsp__issue_active\.avg\.pct_of_peak_sustained_active$|smsp__pcsamp_warps_issue_stalled_[a-z_]+$)')
for h, u, v in sorted(zip(hdr, units, vals)):
    name = h.split('TriageCompute.')[-1]
    if keep.match(name) and not name.endswith('_not_issued'):
        print('%-90s %-12s %s' % (name, u, v))
