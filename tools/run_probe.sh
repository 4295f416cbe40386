#!/bin/bash
# Runs every umma_probe case in its own process (a trap in one cannot poison the next).
mkdir -p gpurun_out
L=gpurun_out/probe.log
: > $L
P=tools/umma_probe
run() { timeout 30 $P "$@" >> $L 2>&1; rc=$?; [ $rc -ne 0 ] && echo "EXIT rc=$rc args: $*" >> $L; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $L 2>&1
# kmajor layout row_bytes tf32 n kblocks ra mtiles shift bo_mode
run kmajor 2 128 0 64 2 256 1 0 0
for n in 16 32 128 256; do run kmajor 2 128 0 $n 2 256 1 0 0; done
for s in 1 2 3 7 8 9 33 34 100; do run kmajor 2 128 0 64 2 256 1 $s 0; run kmajor 2 128 0 64 2 256 1 $s 1; done
run kmajor 2 128 0 128 4 256 1 35 0
for s in 0 1 3 8 33; do run kmajor 4 64 0 64 2 256 1 $s 0; run kmajor 4 64 0 64 2 256 1 $s 1; done
for s in 0 1 5 33; do run kmajor 0 128 0 64 2 256 1 $s 0; done
for s in 0 1 5 33; do run kmajor 2 128 1 64 2 256 1 $s 0; done
# mnmajor m n kr ksteps sa sb
run mnmajor 128 64 128 4 0 0
run mnmajor 128 64 128 4 1 0
run mnmajor 128 64 128 4 0 1
run mnmajor 128 64 128 4 3 5
run mnmajor 128 64 128 4 8 8
run mnmajor 128 64 128 4 33 35
run mnmajor 64 64 128 4 0 0
run mnmajor 64 64 128 4 3 5
run mnmajor 128 128 128 4 3 5
run mnmajor 128 256 128 4 3 5
run mnmajor 64 256 128 4 3 5
run halo
# perf n mtiles taps tap_rows iters layout
for lay in 2 0; do
  for n in 32 64 128 256; do
    run perf $n 1 9 0 300 $lay
    run perf $n 1 9 1 300 $lay
  done
  run perf 64 4 9 1 100 $lay
  run perf 64 4 9 33 100 $lay
  run perf 128 4 9 33 100 $lay
  run perf 128 2 9 33 100 $lay
  run perf 256 2 9 33 100 $lay
done
cat $L
