set -x
python -m pytest tests/test_gpu_round2.py tests/test_cpp_host.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2c_tests.log
POOL=$PWD/raytracer-odin_b200/csrc/libodinrt_b200_pool.so
ORT_LIB=$POOL python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c_tests_pool.log
python tools/trace_bench.py C4 "" $POOL > gpurun_out/r2c_trace_bench_c4.log 2>&1
python tools/trace_bench.py C2 "" $POOL > gpurun_out/r2c_trace_bench_c2.log 2>&1
ORT_LIB=$POOL python tools/tune.py C4 64 ORT_POOL 0,1 > gpurun_out/r2c_tune_c4.log 2>&1
ORT_LIB=$POOL ORT_POOL=1 python tools/tune.py C4 64 ORT_POOL_INNER_MIN 8,12,16,20,24,28 >> gpurun_out/r2c_tune_c4.log 2>&1
ORT_LIB=$POOL ORT_POOL=1 python tools/tune.py C4 64 ORT_POOL_NODE_MIN 8,16,24,32 >> gpurun_out/r2c_tune_c4.log 2>&1
ORT_LIB=$POOL ORT_POOL=1 python tools/tune.py C4 64 ORT_POOL_REFILL 4,8,16,24,32 >> gpurun_out/r2c_tune_c4.log 2>&1
ORT_LIB=$POOL python tools/tune.py C2 64 ORT_POOL 0,1 > gpurun_out/r2c_tune_c2.log 2>&1
cat gpurun_out/r2c_tests.log gpurun_out/r2c_tests_pool.log gpurun_out/r2c_trace_bench_c4.log gpurun_out/r2c_trace_bench_c2.log gpurun_out/r2c_tune_c4.log gpurun_out/r2c_tune_c2.log
