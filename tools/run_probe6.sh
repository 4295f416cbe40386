#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/probe6.log; : > $L
P=tools/umma_probe
run() { timeout 60 $P "$@" >> $L 2>&1; rc=$?; [ $rc -ne 0 ] && echo "EXIT rc=$rc args: $*" >> $L; }
# dual n iters flags warps
for n in 64 128; do
  for nw in 1 2 4; do
    for f in 0 2 6 14 18; do run dual $n 4000 $f $nw; done
  done
done
cat $L
