O=gpurun_out
TAG=${TAG:-r2w}
# everything measured at the end of round 2, in one GPU call (outputs in gpurun_out/, copied to profiles/)
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err
python bench.py --config C5 --bvh device --steps 3 --warmup 3 --spp 64 > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err
python bench.py --config C2 --steps 3 --warmup 3 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
python bench.py --config C3 --steps 3 --warmup 3 --spp 256 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err
cat $O/${TAG}_tests.log; for c in c4 c5 c2 c3; do tail -2 $O/${TAG}_bench_$c.err; cut -c1-200 $O/${TAG}_bench_$c.json; echo; done
# ---- ncu: launch list of the bench command (short form), then one full wave of C4 and of C5
python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > $O/${TAG}_bench_for_ncu.json 2> $O/${TAG}_bench_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_ncu_launches.csv python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > $O/${TAG}_ncu_list.log 2>&1
python tools/ncu_summary.py launches $O/${TAG}_ncu_launches.csv > $O/${TAG}_ncu_launch_summary.txt 2>&1
for C in C4 C5; do
  D=10; [ $C = C5 ] && D=12
  K=$((3*D-1))
  c=$(echo $C | tr A-Z a-z)
  python tools/ncu_wave.py $C > $O/${TAG}_wave_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name regex:'k_trace|k_shade' --launch-skip $K --launch-count $K -f -o /tmp/prof_${TAG}_$c python tools/ncu_wave.py $C > $O/${TAG}_ncu_$c.log 2>&1
  SPW=16; [ $C = C5 ] && SPW=4
  python tools/ncu_summary.py full /tmp/prof_${TAG}_$c.ncu-rep $O/ncu_traffic_$c.json $C $SPW > $O/${TAG}_ncu_full_summary_$c.txt 2>&1
  python tools/ncu_src.py /tmp/prof_${TAG}_$c.ncu-rep "k_trace:k_traceILb0E" 0 45 > $O/${TAG}_ktrace_primary_source_lines_$c.txt 2>&1
  python tools/ncu_src.py /tmp/prof_${TAG}_$c.ncu-rep "k_trace:k_traceILb0E" 1 45 > $O/${TAG}_ktrace_bounce1_source_lines_$c.txt 2>&1
  python tools/ncu_src.py /tmp/prof_${TAG}_$c.ncu-rep "k_shade:k_shade" 1 30 > $O/${TAG}_kshade_source_lines_$c.txt 2>&1
  python tools/ncu_src.py /tmp/prof_${TAG}_$c.ncu-rep "k_trace:k_traceILb1E" 0 45 > $O/${TAG}_klight_bounce1_source_lines_$c.txt 2>&1
done
head -12 $O/${TAG}_ncu_full_summary_c4.txt | cut -c1-200; head -8 $O/${TAG}_ncu_full_summary_c5.txt | cut -c1-200
du -sh $O
