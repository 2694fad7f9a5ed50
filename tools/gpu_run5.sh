ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"radix|scan_|make_sort|sort_plan|gather|cluster_|piece_|bucket_|compact|set_ncur" -c 200 --csv --log-file gpurun_out/launches_cluster_v2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
tail -1 gpurun_out/ncu_bench2.log | head -c 300
