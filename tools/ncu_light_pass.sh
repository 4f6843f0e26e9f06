O=gpurun_out
for C in C4 C3; do
  c=$(echo $C | tr A-Z a-z)
  S=21; [ $C = C3 ] && S=17
  ncu --set full --clock-control none --import-source on --kernel-name regex:'k_trace' --launch-skip $S --launch-count 1 -f -o /tmp/light_$c python tools/ncu_wave.py $C > $O/r2s_ncu_light_$c.log 2>&1
  python tools/ncu_src.py /tmp/light_$c.ncu-rep "k_trace:k_traceILb1E" 0 60 > $O/r2s_klight_bounce1_source_lines_$c.txt 2>&1
  ncu -i /tmp/light_$c.ncu-rep --page raw --csv > /tmp/raw_$c.csv 2>/dev/null
  python - /tmp/raw_$c.csv > $O/r2s_klight_bounce1_metrics_$c.txt <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2:]
keep = ("Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg")
for v in vals:
    for h, u, x in zip(hdr, units, v):
        if h in keep or "stall" in h.lower() and "pct" not in h and "warps_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
            print(f"{h} [{u}] = {x}")
PY
done
tail -3 $O/r2s_ncu_light_c4.log; head -40 $O/r2s_klight_bounce1_source_lines_c4.txt
