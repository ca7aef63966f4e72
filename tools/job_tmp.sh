python -m pytest tests/test_gpu_expansion.py tests/test_gpu_step.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-comparators --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2w_bench.json')); k=d['kernels']
print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['details']['N_mean'], {n:round(v['ms']*v['launches_per_step'],3) for n,v in k.items()})
PY
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py tests/test_io_image.py -m gpu -q -x -k "survivor_list_backward_edge_cases or nosync or denormal or sync_free or image" 2>&1 | tail -6 > gpurun_out/r2w_memcheck.log
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "survivor_list_backward_edge_cases" 2>&1 | tail -6 > gpurun_out/r2w_racecheck.log
cat gpurun_out/r2w_memcheck.log gpurun_out/r2w_racecheck.log
