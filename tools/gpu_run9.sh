O=gpurun_out
python tools/tune.py C4 64 ORT_OVERLAP 1,2,3,4,6 > $O/r2g_tune.log 2>&1
python tools/tune.py C4 64 ORT_WAVE_PATHS 16777216,33554432,67108864 >> $O/r2g_tune.log 2>&1
python tools/tune.py C2 64 ORT_OVERLAP 1,2,4 >> $O/r2g_tune.log 2>&1
python tools/tune.py C5 16 ORT_OVERLAP 1,2,4 >> $O/r2g_tune.log 2>&1
cat $O/r2g_tune.log
