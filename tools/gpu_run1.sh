set -x
python -m pytest tests/test_scan_gpu.py -m gpu -x -q 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads-per-gpu 48000000 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | head -c 600
