O=gpurun_out
nvidia-smi -L | wc -l; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2i_bench_8gpu_c4.json 2> $O/r2i_bench_8gpu.err
tail -2 $O/r2i_bench_8gpu.err; cut -c1-400 $O/r2i_bench_8gpu_c4.json; echo
python tools/sustained.py C5 0,1,2,3,4,5,6,7 125 64 > $O/r2i_sustained_c5_8gpu.log 2>&1
cat $O/r2i_sustained_c5_8gpu.log
rm -f $O/*.png
