python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1
python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2l_dist.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port $((29800+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2l_n$N.json 2> gpurun_out/r2l_n$N.err
done
cat gpurun_out/r2l_smoke.log | tail -2; cat gpurun_out/r2l_dist.log
python -c "
import json
for n in (2,4,8):
    d=json.load(open('gpurun_out/r2l_n%d.json'%n)); print(n, round(d['value'],1), round(d['ms_per_step'],3), round(d['kernels']['adam']['ms']*d['kernels']['adam']['launches_per_step'],3), 'e2e', round(d['e2e']['value'],1))
"
