#!/bin/bash
# Runs every diagnostic group in its own process (a trapped kernel poisons the CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for g in "$@"; do
  echo "##### $g" | tee -a gpurun_out/diag.log
  timeout 300 python tests/gpu_diag.py $g >> gpurun_out/diag.log 2>&1
  echo "exit=$?" | tee -a gpurun_out/diag.log
done
grep -E "^\[(FAIL|PERF)\]|=====|exit=" gpurun_out/diag.log | tail -60
