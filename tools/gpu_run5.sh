set -x
O=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_trace' --launch-skip 19 --launch-count 2 -f -o /tmp/prof_r2e_c4 python tools/ncu_wave.py C4 > $O/r2e_ncu_c4.log 2>&1
python tools/ncu_summary.py full /tmp/prof_r2e_c4.ncu-rep > $O/r2e_ncu_full_summary_c4.txt 2>&1
python tools/ncu_src.py /tmp/prof_r2e_c4.ncu-rep "k_trace:k_traceILb0E" 0 60 > $O/r2e_ktrace_primary_source_lines_c4.txt 2>&1
ncu -i /tmp/prof_r2e_c4.ncu-rep --page details --csv 2>/dev/null | grep -i "stall\|L2\|dram\|hit rate\|Issue\|Eligible\|No Eligible\|One or More" | head -80 > $O/r2e_details_c4.txt
cat $O/r2e_ncu_full_summary_c4.txt; head -70 $O/r2e_ktrace_primary_source_lines_c4.txt
