python -m pytest tests -m gpu -q 2>&1 | tail -70 > gpurun_out/r2g_tests.log
python -m pytest tests/test_gpu_reference_render.py -m gpu -q -x -k checkpoint 2>&1 | tail -40 > gpurun_out/r2g_ckpt.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-comparators"
$B > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
$B --opt bin_impl=1 > gpurun_out/r2g_bench_bin1.json 2>/dev/null
$B --sync-frame > gpurun_out/r2g_bench_sync.json 2>/dev/null
$B --workload gs_mesh_100k_800 > gpurun_out/r2g_bench_cfg2.json 2>/dev/null
$B --workload gs_multi_mesh_2M_1080p > gpurun_out/r2g_bench_cfg4.json 2>/dev/null
python bench.py --steps 200 --warmup 5 --workload gs_mesh_500k_1080p --mode render_animated > gpurun_out/r2g_bench_cfg5.json 2>/dev/null
python bench.py --steps 200 --warmup 5 --workload gs_mesh_500k_1080p --mode render_animated --save-images gpurun_out/tmp_frames --image-format png > gpurun_out/r2g_bench_cfg5_png.json 2>/dev/null
rm -rf gpurun_out/tmp_frames
python bench.py --steps 50 --warmup 5 --workload gs_flat_10k_256 > gpurun_out/r2g_bench_cfg1.json 2>/dev/null
tail -30 gpurun_out/r2g_tests.log; tail -15 gpurun_out/r2g_ckpt.log
