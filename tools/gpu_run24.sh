O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z_smoke.log 2>&1
python bench.py --steps 6 --warmup 3 > $O/r2z_bench_c4.json 2> $O/r2z_bench_c4.err
python bench.py --config C3 --steps 3 --warmup 3 --spp 256 > $O/r2z_bench_c3.json 2> $O/r2z_bench_c3.err
python bench.py --config C2 --steps 3 --warmup 3 > $O/r2z_bench_c2.json 2> $O/r2z_bench_c2.err
cat $O/r2z_tests.log; tail -1 $O/r2z_smoke.log; for c in c4 c3 c2; do tail -2 $O/r2z_bench_$c.err; cut -c1-160 $O/r2z_bench_$c.json; echo; done
