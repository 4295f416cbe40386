#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/probe2.log; : > $L
P=tools/umma_probe
run() { timeout 30 $P "$@" >> $L 2>&1; rc=$?; [ $rc -ne 0 ] && echo "EXIT rc=$rc args: $*" >> $L; }
# mnmajor m n kr ksteps sa sb tf32 sbo layout tmasw
run mnmajor 128 64 128 4 0 0 0 1024 2 3
for sbo in 512 1024 256; do
  for sh in "0 0" "4 4" "3 5" "1 0"; do
    run mnmajor 128 64 128 8 $sh 1 $sbo 1 4
  done
done
run mnmajor 128 64 128 8 0 0 1 512 1 3
run mnmajor 128 64 128 8 0 0 1 512 2 4
run mnmajor 64 32 128 8 3 5 1 512 1 4
run mnmajor 128 128 128 8 3 5 1 512 1 4
cat $L
