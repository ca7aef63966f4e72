TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-comparators > gpurun_out/r2k_n1.json 2> gpurun_out/r2k_n1.err
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port $((29700+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2k_n$N.json 2> gpurun_out/r2k_n$N.err
done
$TR --nproc-per-node 8 --master-port 29720 bench.py --gpus 8 --steps 20 --warmup 5 --workload gs_multi_mesh_2M_1080p > gpurun_out/r2k_cfg4_n8.json 2> gpurun_out/r2k_cfg4_n8.err
$TR --nproc-per-node 8 --master-port 29721 bench.py --gpus 8 --steps 100 --warmup 5 --workload gs_mesh_500k_1080p --mode render_animated > gpurun_out/r2k_cfg5_n8.json 2> gpurun_out/r2k_cfg5_n8.err
NCCL_DEBUG=INFO $TR --nproc-per-node 4 --master-port 29722 bench.py --gpus 4 --steps 5 --warmup 3 2>&1 | grep -i "algo\|proto\|channels\|NVLS\|Connected" | head -30 > gpurun_out/r2k_nccl_n4.log
python -c "
import json
for n in (1,2,4,8):
    d=json.load(open('gpurun_out/r2k_n%d.json'%n)); print(n, round(d['value'],1), round(d['ms_per_step'],3), round(d['kernels']['adam']['ms']*d['kernels']['adam']['launches_per_step'],3))
"
