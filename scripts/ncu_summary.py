"""Print the roofline-relevant metrics of one .ncu-rep raw CSV (ncu -i X.ncu-rep --page raw --csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct']
for vals in rows[2:]:
    d = dict(zip(hdr, zip(units, vals)))
    for k in want:
        if k in d:
            print(f"{k:75s} {d[k][0]:16s} {d[k][1][:110]}")
    st = sorted(((float(v[1]), h) for h, v in d.items() if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and v[1]), reverse=True)
    print("top stalls (warps stalled per issue-active cycle):", ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]}={x:.2f}" for x, h in st[:6]))
    print()
