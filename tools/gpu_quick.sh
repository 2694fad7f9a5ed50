#!/bin/bash
# One short GPU-box check of the extract command line: host inflate vs --gpu-inflate (STRGPU_INFLATE_KERNEL=1|2|3) on a
# synthetic BAM of 6x10^6 reads -- same .bin? -- with the binary's stage reports.  Everything is written to gpurun_out/ step
# by step so that a cut-off call still leaves results.
set +e
mkdir -p gpurun_out
B=strling_b200/bin/strling
O=gpurun_out/q3
$B debug synth-bam /tmp/big.bam 3000000 > ${O}_synth.txt 2>&1
perf() { grep -E "perf|gpu:|rror" | sed 's/.*perf: //'; }
echo "host" > ${O}_runs.txt;          timeout 30 $B extract -v /tmp/big.bam /tmp/c.bin 2>&1 | perf >> ${O}_runs.txt
echo "k3" >> ${O}_runs.txt;           STRGPU_INFLATE_KERNEL=3 timeout 30 $B extract -v --gpu-inflate /tmp/big.bam /tmp/d3.bin 2>&1 | perf >> ${O}_runs.txt
cmp /tmp/c.bin /tmp/d3.bin > /dev/null 2>&1; echo "cmp c.bin d3.bin rc $?" >> ${O}_cmp.txt
echo "k3 shards8" >> ${O}_runs.txt;   STRGPU_INFLATE_KERNEL=3 timeout 30 $B extract -v --gpu-inflate --replay-shards 8 /tmp/big.bam /tmp/d3b.bin 2>&1 | perf >> ${O}_runs.txt
cmp /tmp/c.bin /tmp/d3b.bin > /dev/null 2>&1; echo "cmp c.bin d3b.bin rc $?" >> ${O}_cmp.txt
STRGPU_INFLATE_KERNEL=3 timeout 40 python -m pytest tests/test_decode_gpu.py -x -q -k zlib > ${O}_pytest_k3.txt 2>&1
echo "k3 shards8 batch1M" >> ${O}_runs.txt; STRGPU_INFLATE_KERNEL=3 timeout 30 $B extract -v --gpu-inflate --replay-shards 8 --batch-reads 1048576 /tmp/big.bam /tmp/d3c.bin 2>&1 | perf >> ${O}_runs.txt
cmp /tmp/c.bin /tmp/d3c.bin > /dev/null 2>&1; echo "cmp c.bin d3c.bin rc $?" >> ${O}_cmp.txt
tail -2 ${O}_pytest_k3.txt
cat ${O}_cmp.txt ${O}_runs.txt
