L=$PWD/raytracer-odin_b200/csrc
ORT_LIB=$L/libodinrt_b200_tos.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -3
for C in C4:64 C2:64; do
  for V in tune tos tune tos; do echo $C $V; ORT_LIB=$L/libodinrt_b200_$V.so python tools/tune.py ${C%:*} ${C#*:} ORT_NONE 0 | cut -c1-200; done
done
python tools/trace_bench.py C4 $L/libodinrt_b200_tune.so $L/libodinrt_b200_tos.so 2>&1 | tail -8 | cut -c1-400
