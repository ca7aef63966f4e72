# Round-end evidence on ONE B200 (run through tools/gpurun_retry.sh): tests, smoke, the default bench line, the reference
# arm, the other BASELINE configs, the ncu launch list and one --set full capture of the top kernels.
T=r2t
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${T}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1_full.json 2> gpurun_out/${T}_bench_n1_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
B="python bench.py --steps 20 --warmup 5 --no-comparators"
$B --no-optimizer --no-cpu-baseline > gpurun_out/${T}_bench_n1_no_optimizer.json 2>/dev/null
$B --workload gs_mesh_100k_800 > gpurun_out/${T}_bench_cfg2_n1.json 2>/dev/null
$B --workload gs_multi_mesh_2M_1080p --no-cpu-baseline > gpurun_out/${T}_bench_cfg4_n1.json 2>/dev/null
python bench.py --steps 200 --warmup 5 --workload gs_mesh_500k_1080p --mode render_animated > gpurun_out/${T}_bench_cfg5_n1.json 2>/dev/null
python bench.py --steps 200 --warmup 5 --workload gs_mesh_500k_1080p --mode render_animated --save-images gpurun_out/tmp_frames --image-format png > gpurun_out/${T}_bench_cfg5_png_n1.json 2>/dev/null
rm -rf gpurun_out/tmp_frames
python bench.py --steps 50 --warmup 5 --workload gs_flat_10k_256 > gpurun_out/${T}_bench_cfg1_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1100 -c 330 --csv --log-file gpurun_out/${T}_ncu_launches_n1.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-comparators > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_adam_sh|k_composite_fwd2|k_composite_bwd5|k_preprocess_bwd|k_ssim" --launch-skip 100 --launch-count 6 -f -o gpurun_out/${T}_ncu python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-comparators > gpurun_out/${T}_ncu_bench.log 2>&1
tail -3 gpurun_out/${T}_smoke.log; tail -4 gpurun_out/${T}_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2t_bench_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], round(d.get('value',0),2), d.get('unit'), 'e2e', (d.get('e2e') or {}).get('value'))
    except Exception as e: print(f, 'failed', e)
PY
