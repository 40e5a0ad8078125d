#!/bin/bash
O=gpurun_out/r02p
mkdir -p $O
for c in 1 2 3 4; do echo "chunk_tiles=$c"; RDX_CHUNK=$c RDX_BIG_ONLY=1 timeout 120 ./tools/radix_test 400000000; done > $O/radix_sweep.txt 2>&1
cat $O/radix_sweep.txt
timeout 300 ./tools/radix_test 100000000 > $O/radix_test.txt 2>&1; echo "rc=$?" >> $O/radix_test.txt
tail -4 $O/radix_test.txt
