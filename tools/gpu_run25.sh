L=$PWD/raytracer-odin_b200/csrc
for C in C4:64 C2:64; do
  for V in tune stk8 stk12 tune stk8 stk12; do echo $C $V; ORT_LIB=$L/libodinrt_b200_$V.so python tools/tune.py ${C%:*} ${C#*:} ORT_NONE 0 | cut -c1-200; done
done
