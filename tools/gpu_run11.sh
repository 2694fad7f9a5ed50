for n in 1 8; do timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 29581 tools/h2d_ceiling.py 2>/dev/null | tail -1; done | tee gpurun_out/h2d_ceiling.jsonl
STRGPU_COMM_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 2 --warmup 2 --no-joint --parity-sample 0 > gpurun_out/t11.out 2> gpurun_out/t11.err
echo rc=$?
grep -o "\[strgpu rank [0-7]\] sharded cluster:.*" gpurun_out/t11.err | grep -v "pair_cap 60" | sort | uniq | head -40
