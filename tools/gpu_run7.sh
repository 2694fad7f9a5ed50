timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 2 --warmup 3 --no-joint --parity-sample 0 --reads-per-gpu 24000000 > gpurun_out/t7.out 2> gpurun_out/t7.err
echo rc=$?
tail -4 gpurun_out/t7.err | cut -c1-400
head -c 600 gpurun_out/t7.out
