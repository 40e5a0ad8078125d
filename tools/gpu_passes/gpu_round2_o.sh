#!/bin/bash
O=gpurun_out/r02o
mkdir -p $O
for b in 11 10 9 8; do for c in 1 2; do echo "bits=$b chunk_tiles=$c"; RDX_BITS=$b RDX_CHUNK=$c RDX_BIG_ONLY=1 timeout 120 ./tools/radix_test 400000000; done; done > $O/radix_sweep.txt 2>&1
cat $O/radix_sweep.txt
RDX_CHUNK=2 RDX_BIG_ONLY=1 timeout 120 ./tools/radix_test 400000000 > $O/plain.txt 2>&1 && \
RDX_CHUNK=2 RDX_BIG_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rdx_scatter -s 12 -c 1 -o $O/prof_rdx2 -f ./tools/radix_test 400000000 > $O/ncu_rdx.log 2>&1
tail -3 $O/ncu_rdx.log
