python -m pytest tests/test_gpu_step.py -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-comparators > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python bench.py --steps 20 --warmup 5 --no-comparators --opt bwd_minb=8 > gpurun_out/r2n_bench_minb8.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-comparators --opt bwd_minb=5 > gpurun_out/r2n_bench_minb5.json 2>/dev/null
python - <<'PY'
import json
for f in ('r2n_bench','r2n_bench_minb8','r2n_bench_minb5'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); k=d['kernels']
        print(f, round(d['value'],1), round(d['ms_per_step'],3), {n:round(v['ms'],3) for n,v in k.items() if n in ('ssim_stats','ssim_grad','composite_bwd','composite_fwd','adam')})
    except Exception as e: print(f, 'failed', e)
PY
