O=gpurun_out
TAG=r2z
python bench.py --config C5 --bvh device --steps 3 --warmup 3 --spp 64 > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err
python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > $O/${TAG}_bench_for_ncu.json 2> $O/${TAG}_bench_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_ncu_launches.csv python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > $O/${TAG}_ncu_list.log 2>&1
python tools/ncu_summary.py launches $O/${TAG}_ncu_launches.csv > $O/${TAG}_ncu_launch_summary.txt 2>&1
C=C4; c=c4; K=29
python tools/ncu_wave.py $C > $O/${TAG}_wave_$c.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_trace|k_shade' --launch-skip $K --launch-count $K -f -o /tmp/prof_${TAG}_$c python tools/ncu_wave.py $C > $O/${TAG}_ncu_$c.log 2>&1
python tools/ncu_summary.py full /tmp/prof_${TAG}_$c.ncu-rep $O/ncu_traffic_$c.json $C 16 > $O/${TAG}_ncu_full_summary_$c.txt 2>&1
python tools/ncu_src.py /tmp/prof_${TAG}_$c.ncu-rep "k_trace:k_traceILb0E" 1 45 > $O/${TAG}_ktrace_bounce1_source_lines_$c.txt 2>&1
python tools/ncu_src.py /tmp/prof_${TAG}_$c.ncu-rep "k_trace:k_traceILb1E" 0 45 > $O/${TAG}_klight_bounce1_source_lines_$c.txt 2>&1
cut -c1-200 $O/${TAG}_bench_c5.json; head -6 $O/${TAG}_ncu_launch_summary.txt; head -6 $O/${TAG}_ncu_full_summary_c4.txt | cut -c1-220
