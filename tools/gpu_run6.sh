python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -5 gpurun_out/bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json'))
for k in ('value','ms_per_step','cluster','strong','joint','parity_sample'): print(k, d.get(k))
print('e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
"
