python -m pytest tests/test_gpu_dist.py tests/test_gpu_step.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2h_dist_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-comparators > gpurun_out/r2h_bench_n1.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-comparators --no-native > gpurun_out/r2h_bench_n1_dense.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
tail -12 gpurun_out/r2h_dist_tests.log; tail -3 gpurun_out/r2h_bench_n2.err; head -c 300 gpurun_out/r2h_bench_n2.json
