#!/bin/bash
# One short GPU-box check of the command line: `strling extract` with host inflate and with --gpu-inflate (STRGPU_INFLATE_KERNEL=1|2|3)
# on a synthetic BAM of 6x10^6 reads -- same .bin? --, then `strling call`, with the binaries' stage reports.  Everything is
# written to gpurun_out/ step by step so that a cut-off call still leaves results.  (profiles/r2_cli_gpu_box_v*.txt are runs of
# earlier versions of this script.)
set +e
mkdir -p gpurun_out
B=strling_b200/bin/strling
O=gpurun_out/q
$B debug synth-bam /tmp/big.bam 3000000 > ${O}_synth.txt 2>&1
perf() { grep -E "perf|gpu:|rror" | sed 's/.*perf: //'; }
echo "host" > ${O}_runs.txt; timeout 60 $B extract -v /tmp/big.bam /tmp/c.bin 2>&1 | perf >> ${O}_runs.txt
# the oracle pipeline's .bin for this BAM (tools/cli_oracle_md5.py 3000000 2): md5 ab5b3864dcff18faea27cd62b6b20932
md5sum /tmp/c.bin > ${O}_cmp.txt
for k in 1 2 3; do
  echo "gpu-inflate kernel $k" >> ${O}_runs.txt
  STRGPU_INFLATE_KERNEL=$k timeout 60 $B extract -v --gpu-inflate /tmp/big.bam /tmp/d$k.bin 2>&1 | perf >> ${O}_runs.txt
  cmp /tmp/c.bin /tmp/d$k.bin > /dev/null 2>&1; echo "cmp c.bin d$k.bin rc $?" >> ${O}_cmp.txt
done
( time STRLING_CALL_TIMING=1 $B call -o /tmp/c /tmp/big.bam /tmp/c.bin ) > ${O}_call.txt 2>&1; echo "rc $?" >> ${O}_call.txt
wc -l /tmp/c-bounds.txt /tmp/c-genotype.txt /tmp/c-unplaced.txt >> ${O}_call.txt 2>&1
cat ${O}_cmp.txt ${O}_runs.txt; tail -12 ${O}_call.txt
