python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2f_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-comparators"
$B > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
$B --opt bwd_group=3 --opt bwd_minblocks=6 > gpurun_out/r2f_bench_g3m6.json 2>/dev/null
$B --opt bwd_group=3 --opt bwd_minblocks=5 > gpurun_out/r2f_bench_g3m5.json 2>/dev/null
$B --opt bwd_group=3 --opt bwd_minblocks=4 > gpurun_out/r2f_bench_g3m4.json 2>/dev/null
$B --opt bwd_group=1 --opt bwd_minblocks=5 > gpurun_out/r2f_bench_g1m5.json 2>/dev/null
$B --opt bin_impl=0 > gpurun_out/r2f_bench_bin0.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_bin_tiles|k_adam_sh" --launch-skip 40 --launch-count 2 -f -o gpurun_out/r2f_ncu python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-comparators > gpurun_out/r2f_ncu_bench.log 2>&1
tail -25 gpurun_out/r2f_tests.log
