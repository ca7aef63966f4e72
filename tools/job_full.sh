python -m pytest tests/test_gpu_step.py -m gpu -q -x -s -k "denormal" 2>&1 | grep -v "^$" | tail -6
for o in "adam_sh_ieee=0" "adam_sh_ieee=1"; do
python bench.py --steps 20 --warmup 5 --no-comparators --no-cpu-baseline --opt $o > gpurun_out/r2s_$o.json 2> gpurun_out/r2s_$o.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s_*.json')):
    d=json.load(open(f)); k=d['kernels']
    print(f, round(d['value'],1), round(d['ms_per_step'],3), d['details']['N_mean'], round(k['adam']['ms']*k['adam']['launches_per_step'],3))
PY
