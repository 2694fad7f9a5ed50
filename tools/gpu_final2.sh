ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cli --parity-sample 0 > gpurun_out/ncu_final.log 2>&1
tail -1 gpurun_out/ncu_final.log | head -c 200
wc -l gpurun_out/launches_final.csv
