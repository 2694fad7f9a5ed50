timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29555 tools/sharded_check.py 4000 2>&1 | grep -E "merge_mode|sharded cluster|differs|rror" | tail -4
echo rc=$?
timeout 220 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 --parity-sample 0 > gpurun_out/t8.out 2> gpurun_out/t8.err
echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/t8.out'))
print('value',d['value'],'ms',d['ms_per_step'],'cluster_ms',d['cluster']['ms_per_step'],d['cluster']['sharded_equals_single_gpu'],'strong',d.get('strong',{}).get('value'), d.get('strong',{}).get('ms_per_step'), 'joint', d.get('joint',{}).get('equals_single_gpu_merge'), 'e2e', d['e2e']['value'])"
tail -3 gpurun_out/t8.err | cut -c1-300
