O=gpurun_out
# everything measured at the end of round 2, in one GPU call (outputs in gpurun_out/, copied to profiles/)
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2j_tests.log
python bench.py --steps 20 --warmup 5 > $O/r2j_bench_c4.json 2> $O/r2j_bench_c4.err
python bench.py --config C5 --bvh device --steps 3 --warmup 3 --spp 64 > $O/r2j_bench_c5.json 2> $O/r2j_bench_c5.err
python bench.py --config C2 --steps 3 --warmup 3 > $O/r2j_bench_c2.json 2> $O/r2j_bench_c2.err
python bench.py --config C3 --steps 3 --warmup 3 --spp 256 > $O/r2j_bench_c3.json 2> $O/r2j_bench_c3.err
cat $O/r2j_tests.log; for c in c4 c5 c2 c3; do tail -2 $O/r2j_bench_$c.err; cut -c1-200 $O/r2j_bench_$c.json; echo; done
# ---- ncu: launch list of the bench command (short form), then one full wave of C4 and of C5
python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > $O/r2j_bench_for_ncu.json 2> $O/r2j_bench_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2j_ncu_launches.csv python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > $O/r2j_ncu_list.log 2>&1
python tools/ncu_summary.py launches $O/r2j_ncu_launches.csv > $O/r2j_ncu_launch_summary.txt 2>&1
for C in C4 C5; do
  D=10; [ $C = C5 ] && D=12
  K=$((3*D-1))
  c=$(echo $C | tr A-Z a-z)
  python tools/ncu_wave.py $C > $O/r2j_wave_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name regex:'k_trace|k_shade' --launch-skip $K --launch-count $K -f -o /tmp/prof_r2j_$c python tools/ncu_wave.py $C > $O/r2j_ncu_$c.log 2>&1
  SPW=16; [ $C = C5 ] && SPW=4
  python tools/ncu_summary.py full /tmp/prof_r2j_$c.ncu-rep $O/ncu_traffic_$c.json $C $SPW > $O/r2j_ncu_full_summary_$c.txt 2>&1
  python tools/ncu_src.py /tmp/prof_r2j_$c.ncu-rep "k_trace:k_traceILb0E" 0 45 > $O/r2j_ktrace_primary_source_lines_$c.txt 2>&1
  python tools/ncu_src.py /tmp/prof_r2j_$c.ncu-rep "k_trace:k_traceILb0E" 1 45 > $O/r2j_ktrace_bounce1_source_lines_$c.txt 2>&1
  python tools/ncu_src.py /tmp/prof_r2j_$c.ncu-rep "k_shade:k_shade" 1 30 > $O/r2j_kshade_source_lines_$c.txt 2>&1
done
head -12 $O/r2j_ncu_full_summary_c4.txt | cut -c1-200; head -8 $O/r2j_ncu_full_summary_c5.txt | cut -c1-200
du -sh $O
