set -x
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -s 2>&1 | tail -30 > $O/r2f_tests.log
python tools/primary_chunks.py C4 > $O/r2f_primary_chunks_c4.log 2>&1
python tools/tune.py C4 64 ORT_NONE 0 > $O/r2f_tune.log 2>&1
python tools/tune.py C2 64 ORT_NONE 0 >> $O/r2f_tune.log 2>&1
python tools/tune.py C3 64 ORT_NONE 0 >> $O/r2f_tune.log 2>&1
python tools/tune.py C5 16 ORT_NONE 0 >> $O/r2f_tune.log 2>&1
cat $O/r2f_tests.log $O/r2f_primary_chunks_c4.log $O/r2f_tune.log
