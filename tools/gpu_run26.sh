O=gpurun_out
python -m pytest tests -m gpu -q -k "multi" 2>&1 | tail -4 > $O/r2ab_multi_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --spp 512 --no-cpu > $O/r2ab_bench_2gpu_c4_512spp.json 2> $O/r2ab_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 3 > $O/r2ab_ref_2gpu.json 2> $O/r2ab_ref_2gpu.err
cat $O/r2ab_multi_tests.log; tail -3 $O/r2ab_bench_2gpu.err; cut -c1-300 $O/r2ab_bench_2gpu_c4_512spp.json; cut -c1-200 $O/r2ab_ref_2gpu.json
