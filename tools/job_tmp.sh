python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-comparators --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2x_bench.json')); k=d['kernels']
print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['details']['N_mean'], {n:round(v['ms']*v['launches_per_step'],3) for n,v in k.items()})
PY
