timeout 900 python -m pytest tests/test_cluster_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -6
ncu --set full --clock-control none --import-source on -k regex:"repeat_prefilter|ladder_stage" -s 8 -c 3 -o gpurun_out/k1_v12 -f python tools/profile_scan.py --calls 2 --regions 2 > gpurun_out/ncu_v12.log 2>&1
tail -2 gpurun_out/ncu_v12.log
