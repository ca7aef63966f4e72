python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2i_tests.log
B="python bench.py --steps 20 --warmup 5"
$B > gpurun_out/r2i_bench_full.json 2> gpurun_out/r2i_bench_full.err
$B --no-cpu-baseline --no-comparators --no-native > gpurun_out/r2i_bench_dense.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_adam_sh|k_composite_fwd2|k_composite_bwd5|k_preprocess_bwd" --launch-skip 80 --launch-count 4 -f -o gpurun_out/r2i_ncu python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-comparators > gpurun_out/r2i_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_ncu_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-comparators > gpurun_out/r2i_ncu_launches_bench.log 2>&1
tail -12 gpurun_out/r2i_tests.log; tail -c 600 gpurun_out/r2i_bench_full.err
