L=$PWD/raytracer-odin_b200/csrc
for C in C4:64 C2:64 C3:64 C5:16; do
  for V in l8 l9 l10 l8 l9; do echo $C $V; ORT_LIB=$L/libodinrt_b200_$V.so python tools/tune.py ${C%:*} ${C#*:} ORT_LIGHT_PREFILTER 2 | cut -c1-200; done
done
