#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/probe5.log; : > $L
P=tools/umma_probe
run() { timeout 60 $P "$@" >> $L 2>&1; rc=$?; [ $rc -ne 0 ] && echo "EXIT rc=$rc args: $*" >> $L; }
# lean n mt iters flags(1 whole-warp, 2 commit/8 MMAs, 4 wait/8 MMAs, 8 fence/8 MMAs)
for n in 64 128; do for f in 1 3 5 9 7 15; do run lean $n 2 4000 $f; done; done
for f in 1 3 7 15; do run lean 128 1 8000 $f; done
cat $L
