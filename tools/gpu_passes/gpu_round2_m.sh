#!/bin/bash
O=gpurun_out/r02m
mkdir -p $O
./tools/microbench3 > $O/microbench3.txt 2>&1
cat $O/microbench3.txt
for b in 11 10 9 8; do echo "RDX_BITS=$b"; RDX_BITS=$b RDX_BIG_ONLY=1 ./tools/radix_test 400000000; done > $O/radix_bits.txt 2>&1
cat $O/radix_bits.txt
