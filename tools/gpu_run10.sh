O=gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -k "multi" -s 2>&1 | tail -8 > $O/r2h_tests_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --spp 512 > $O/r2h_bench_2gpu_c4_512.json 2> $O/r2h_bench_2gpu.err
python bench.py --gpus 1 --steps 3 --warmup 3 --spp 512 --no-cpu > $O/r2h_bench_1gpu_c4_512.json 2>> $O/r2h_bench_2gpu.err
python tools/sustained.py C4 0,1 25 64 > $O/r2h_sustained_c4_2gpu.log 2>&1
python tools/sustained.py C4 0 25 64 > $O/r2h_sustained_c4_1gpu.log 2>&1
cat $O/r2h_tests_2gpu.log; tail -3 $O/r2h_bench_2gpu.err; cut -c1-700 $O/r2h_bench_2gpu_c4_512.json; echo; cut -c1-300 $O/r2h_bench_1gpu_c4_512.json; echo; cat $O/r2h_sustained_c4_2gpu.log $O/r2h_sustained_c4_1gpu.log
rm -f $O/*.png
