set -x
O=gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -q -k "c4_window" 2>&1 | tail -5 > $O/r2d_tests.log
L=$PWD/raytracer-odin_b200/csrc
python tools/trace_bench.py C4 "" $L/libodinrt_b200_pf.so $L/libodinrt_b200_na.so $L/libodinrt_b200_pfna.so > $O/r2d_trace_bench_c4.log 2>&1
python tools/trace_bench.py C2 "" $L/libodinrt_b200_pf.so $L/libodinrt_b200_na.so $L/libodinrt_b200_pfna.so > $O/r2d_trace_bench_c2.log 2>&1
for v in "" _pf _na _pfna; do ORT_LIB=$L/libodinrt_b200$v.so python tools/tune.py C4 64 ORT_NONE 0 >> $O/r2d_tune_c4.log 2>&1; done
cat $O/r2d_tests.log $O/r2d_trace_bench_c4.log $O/r2d_trace_bench_c2.log $O/r2d_tune_c4.log
# ---- ncu: launch list of the bench command, then one full wave of C4 and of C5 (summaries made here: the reports are too big to ship)
python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > $O/r2d_bench_for_ncu.json 2> $O/r2d_bench_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2d_ncu_launches.csv python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > $O/r2d_ncu_list.log 2>&1
python tools/ncu_summary.py launches $O/r2d_ncu_launches.csv > $O/r2d_ncu_launch_summary.txt 2>&1
for C in C4 C5; do
  D=10; [ $C = C5 ] && D=12
  K=$((3*D-1))
  c=$(echo $C | tr A-Z a-z)
  python tools/ncu_wave.py $C > $O/r2d_wave_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name regex:'k_trace|k_shade' --launch-skip $K --launch-count $K -f -o /tmp/prof_r2d_$c python tools/ncu_wave.py $C > $O/r2d_ncu_$c.log 2>&1
  SPW=16; [ $C = C5 ] && SPW=4
  python tools/ncu_summary.py full /tmp/prof_r2d_$c.ncu-rep $O/ncu_traffic_$c.json $C $SPW > $O/r2d_ncu_full_summary_$c.txt 2>&1
  python tools/ncu_src.py /tmp/prof_r2d_$c.ncu-rep "k_trace:k_traceILb0E" 1 45 > $O/r2d_ktrace_closest_source_lines_$c.txt 2>&1
  python tools/ncu_src.py /tmp/prof_r2d_$c.ncu-rep "k_shade:k_shade" 1 30 > $O/r2d_kshade_source_lines_$c.txt 2>&1
  ls -la /tmp/prof_r2d_$c.ncu-rep
done
tail -n 4 $O/r2d_wave_c4.log $O/r2d_wave_c5.log $O/r2d_ncu_c4.log $O/r2d_ncu_c5.log
head -30 $O/r2d_ncu_full_summary_c4.txt
du -sh $O
