ncu --set full --clock-control none --import-source on -k regex:"ladder_stage" -s 4 -c 2 -o gpurun_out/k1_v11 -f python tools/profile_scan.py --calls 2 --regions 2 > gpurun_out/ncu_v11.log 2>&1
tail -2 gpurun_out/ncu_v11.log
