B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-comparators"
$B > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1100 -c 330 --csv --log-file gpurun_out/r2j_ncu_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-comparators > gpurun_out/r2j_ncu_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_adam_sh" --launch-skip 20 --launch-count 1 -f -o gpurun_out/r2j_ncu_adam python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-comparators > /dev/null 2>&1
python -c "
import json; d=json.load(open('gpurun_out/r2j_bench.json')); k=d['kernels']
print(d['value'], d['ms_per_step'], {n:(round(v['ms'],4),v['launches_per_step']) for n,v in k.items()})"
