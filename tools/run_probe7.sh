#!/bin/bash
# cta_group::2 probe (correctness for both N, row-shifted A view, issue rate) next to the single-CTA rates
mkdir -p gpurun_out; L=gpurun_out/probe7.log; : > $L
P=tools/umma_probe
run() { timeout 60 $P "$@" >> $L 2>&1; rc=$?; [ $rc -ne 0 ] && echo "EXIT rc=$rc args: $*" >> $L; }
run cta2 256 2 0 4000
run cta2 256 2 5 4000
run cta2 128 2 0 4000
run cta2 128 4 33 4000
run cta2 64 2 0 4000
run dual 128 4000 0 2
run dual 256 4000 0 2
run dual 128 4000 0 1
run dual 256 4000 0 1
cat $L
