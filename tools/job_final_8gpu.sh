# Round-end evidence on an 8-GPU box: the 2-rank distribution test, the scaling run and BASELINE configs 4 / 5 at N = 8.
T=r2t
python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -5 > gpurun_out/${T}_dist_test.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-comparators > gpurun_out/${T}_scale_n1.json 2>/dev/null
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port $((29800+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_scale_n$N.json 2> gpurun_out/${T}_scale_n$N.err
done
$TR --nproc-per-node 8 --master-port 29820 bench.py --gpus 8 --steps 20 --warmup 5 --workload gs_multi_mesh_2M_1080p > gpurun_out/${T}_bench_cfg4_n8.json 2>/dev/null
$TR --nproc-per-node 8 --master-port 29821 bench.py --gpus 8 --steps 100 --warmup 5 --workload gs_mesh_500k_1080p --mode render_animated > gpurun_out/${T}_bench_cfg5_n8.json 2>/dev/null
cat gpurun_out/${T}_dist_test.log
python - <<'PY'
import json
for n in (1,2,4,8):
    d=json.load(open('gpurun_out/r2t_scale_n%d.json'%n)); print(n, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
for c in ('cfg4','cfg5'):
    d=json.load(open('gpurun_out/r2t_bench_%s_n8.json'%c)); print(c, round(d['value'],1))
PY
