STRGPU_COMM_TIMING=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 2 --warmup 3 --no-joint --no-strong --parity-sample 0 > gpurun_out/t7.out 2> gpurun_out/t7.err
echo rc=$?
grep "strgpu rank 0" gpurun_out/t7.err | tail -3
tail -5 gpurun_out/t7.err | cut -c1-300
