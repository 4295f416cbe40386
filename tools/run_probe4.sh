#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/probe4.log; : > $L
P=tools/umma_probe
run() { timeout 60 $P "$@" >> $L 2>&1; rc=$?; [ $rc -ne 0 ] && echo "EXIT rc=$rc args: $*" >> $L; }
# mix n mtiles taps iters nloads chunk stages gMB same commit_every tap_rows pollers fence_every
run mix 128 2 9 400 0 16384 8 64 0  0 0 0 0
run mix 128 2 9 400 0 16384 8 64 0  2 0 0 0
run mix 128 2 9 400 0 16384 8 64 0  1 0 0 0
run mix 128 2 9 400 0 16384 8 64 0  0 33 0 0
run mix 128 2 9 400 0 16384 8 64 0  0 65 0 0
run mix 128 2 9 400 0 16384 8 64 0  0 0 3 0
run mix 128 2 9 400 0 16384 8 64 0  0 0 0 1
run mix 128 2 9 400 0 16384 8 64 0  2 65 3 1
run mix 128 2 9 400 5000 16384 8 64 0  2 65 3 1
run mix 128 2 9 400 5000 16384 8 64 0  0 65 0 0
cat $L
