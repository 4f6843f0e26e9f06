set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2b_tests.log
cat gpurun_out/r2b_tests.log
python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > gpurun_out/r2b_bench_for_ncu.json 2> gpurun_out/r2b_bench_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_ncu_launches.csv python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu > gpurun_out/r2b_ncu_list.log 2>&1
python tools/ncu_wave.py C4 > gpurun_out/r2b_wave_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_trace|k_shade' --launch-skip 29 --launch-count 29 -f -o gpurun_out/prof_r2b_c4 python tools/ncu_wave.py C4 > gpurun_out/r2b_ncu_c4.log 2>&1
python tools/ncu_wave.py C5 > gpurun_out/r2b_wave_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_trace|k_shade' --launch-skip 35 --launch-count 35 -f -o gpurun_out/prof_r2b_c5 python tools/ncu_wave.py C5 > gpurun_out/r2b_ncu_c5.log 2>&1
tail -3 gpurun_out/r2b_wave_c4.log gpurun_out/r2b_wave_c5.log gpurun_out/r2b_ncu_c4.log gpurun_out/r2b_ncu_c5.log
ls -la gpurun_out | tail -12
