L=$PWD/raytracer-odin_b200/csrc
O=gpurun_out
for C in C4:64 C2:64 C3:64; do
  for V in tune s32 tune s32; do echo $C $V; ORT_LIB=$L/libodinrt_b200_$V.so python tools/tune.py ${C%:*} ${C#*:} ORT_NONE 0 | cut -c1-200; done
done
# all raw metrics + ncu's own rule output for ONE closest-hit launch on bounce-1 rays of C4 (k_trace launches of a
# wave: closest b0, closest b1, light b1, ...; the warm-up wave has 19)
ncu --set full --clock-control none --kernel-name regex:'k_trace' --launch-skip 20 --launch-count 1 -f -o /tmp/cb1 python tools/ncu_wave.py C4 > $O/r2y_ncu.log 2>&1
ncu -i /tmp/cb1.ncu-rep --page raw --csv > $O/r2y_closest_b1_raw_c4.csv 2>/dev/null
ncu -i /tmp/cb1.ncu-rep --page details > $O/r2y_closest_b1_details_c4.txt 2>/dev/null
wc -c $O/r2y_closest_b1_raw_c4.csv $O/r2y_closest_b1_details_c4.txt
