timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_n4.json'))
print('value',d['value'],'ms',d['ms_per_step'],'cluster_ms',d['cluster']['ms_per_step'],d['cluster']['sharded_equals_single_gpu'],'strong',d['strong']['value'],d['strong']['ms_per_step'],'joint',d['joint']['equals_single_gpu_merge'],'e2e',d['e2e']['value'],'parity',d['parity_sample']['mismatches'])"
tail -2 gpurun_out/bench_n4.err | cut -c1-200
