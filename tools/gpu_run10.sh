timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json'))
for k in ('value','ms_per_step','cluster','strong','joint','parity_sample'): print(k, d.get(k))
print('e2e', d['e2e']['value'], d['e2e'].get('h2d_gb_per_s_per_gpu'), 'frac', d['roofline']['frac'])
"
tail -2 gpurun_out/bench_n8.err | cut -c1-300
