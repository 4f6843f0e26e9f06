L=$PWD/raytracer-odin_b200/csrc
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -5
for C in C4 C2 C3; do
  for V in tune pair; do echo $C $V; ORT_LIB=$L/libodinrt_b200_$V.so python tools/tune.py $C 64 ORT_LIGHT_PREFILTER 2,2; done
done
echo C5; ORT_LIB=$L/libodinrt_b200_tune.so python tools/tune.py C5 16 ORT_LIGHT_PREFILTER 2
