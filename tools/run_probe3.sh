#!/bin/bash
# smem-port / ingest interference probe (see "mix" in umma_probe.cu)
mkdir -p gpurun_out; L=gpurun_out/probe3.log; : > $L
P=tools/umma_probe
run() { timeout 60 $P "$@" >> $L 2>&1; rc=$?; [ $rc -ne 0 ] && echo "EXIT rc=$rc args: $*" >> $L; }
# mix n mtiles taps iters nloads chunk stages gMB same
run mix 128 2 9 400 0
run mix 256 1 9 400 0
run mix 64 2 9 400 0
for st in 4 8; do
  run mix 128 2 9 0 5000 16384 $st 64 0
  run mix 128 2 9 0 5000 16384 $st 0 1
done
run mix 128 2 9 0 2500 32768 4 64 0
for nl in 1250 2500 3750 5000 7500; do
  run mix 128 2 9 400 $nl 16384 8 64 0
done
run mix 128 2 9 400 5000 16384 8 0 1
run mix 128 2 9 400 2500 16384 8 0 1
for nl in 2500 5000; do
  run mix 256 1 9 400 $nl 16384 8 64 0
  run mix 64 2 9 400 $nl 16384 8 64 0
done
cat $L
