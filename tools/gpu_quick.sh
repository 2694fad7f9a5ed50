#!/bin/bash
# One short GPU-box check of the extract command line: host inflate vs --gpu-inflate on synthetic BAMs (same .bin?), with the
# binary's stage reports.  Everything is written to gpurun_out/ step by step so that a cut-off call still leaves results.
set +e
mkdir -p gpurun_out
B=strling_b200/bin/strling
{ nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; } > gpurun_out/q_host.txt 2>&1
$B debug synth-bam /tmp/q.bam 400000 > gpurun_out/q_synth.txt 2>&1
timeout 60 $B extract -v /tmp/q.bam /tmp/a.bin 2> gpurun_out/q_small_host.txt; echo "rc $?" >> gpurun_out/q_small_host.txt
timeout 60 $B extract -v --gpu-inflate /tmp/q.bam /tmp/b.bin 2> gpurun_out/q_small_gpuinflate.txt; echo "rc $?" >> gpurun_out/q_small_gpuinflate.txt
cmp /tmp/a.bin /tmp/b.bin > gpurun_out/q_small_cmp.txt 2>&1; echo "cmp rc $?" >> gpurun_out/q_small_cmp.txt
$B debug synth-bam /tmp/big.bam 3000000 >> gpurun_out/q_synth.txt 2>&1
for i in 1 2; do
timeout 90 $B extract -v /tmp/big.bam /tmp/c.bin 2>&1 | grep perf >> gpurun_out/q_big_host.txt
timeout 90 $B extract -v --gpu-inflate /tmp/big.bam /tmp/d.bin 2>&1 | grep -E "perf|gpu:" >> gpurun_out/q_big_gpuinflate.txt
done
cmp /tmp/c.bin /tmp/d.bin > gpurun_out/q_big_cmp.txt 2>&1; echo "cmp rc $?" >> gpurun_out/q_big_cmp.txt
timeout 100 python -m pytest tests/test_cli_gpu.py tests/test_decode_gpu.py -x -q -k "config1 or config4 or inflate" > gpurun_out/q_pytest.txt 2>&1
tail -3 gpurun_out/q_pytest.txt
cat gpurun_out/q_small_cmp.txt gpurun_out/q_big_cmp.txt gpurun_out/q_big_host.txt gpurun_out/q_big_gpuinflate.txt
