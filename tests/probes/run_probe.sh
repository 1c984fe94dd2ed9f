#!/bin/bash
# Runs each probe group in its own process (a trapped kernel poisons the CUDA context).
mkdir -p gpurun_out
: > gpurun_out/probe.txt
for g in "$@"; do
  timeout 120 tests/probes/umma_probe $g >> gpurun_out/probe.txt 2>&1
  echo "exit($g)=$?" >> gpurun_out/probe.txt
done
cat gpurun_out/probe.txt
