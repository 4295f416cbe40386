#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/probe5.log; : > $L
P=tools/umma_probe
run() { timeout 60 $P "$@" >> $L 2>&1; rc=$?; [ $rc -ne 0 ] && echo "EXIT rc=$rc args: $*" >> $L; }
# lean n mt iters whole_warp
for n in 16 32 64 96 128 192 256; do run lean $n 2 4000 0; done
for n in 64 128; do run lean $n 2 4000 1; run lean $n 1 8000 0; run lean $n 4 2000 0; done
cat $L
