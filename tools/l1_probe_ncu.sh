#!/bin/bash
# ncu's view of the L1 probe patterns: what "l1tex__data_pipe_lsu_wavefronts % of peak" reads when the pattern runs flat out
M="l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct"
for size in 0 1; do for mode in 0 1 3 4; do
  ncu --metrics $M --clock-control none -k regex:probe -s 1 -c 1 --csv --log-file gpurun_out/l1ncu_${size}_${mode}.csv build/l1_probe $size $mode > /dev/null 2>&1
done; done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob("gpurun_out/l1ncu_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    if len(rows)<2: print(f,"no data"); continue
    h=rows[0]; i_n=h.index("Metric Name"); i_v=h.index("Metric Value")
    print(f, {r[i_n].replace("l1tex__",""):r[i_v] for r in rows[1:]})
PY
