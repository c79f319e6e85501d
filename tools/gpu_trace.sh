#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/k4_trace.py > gpurun_out/k4_trace.log 2>&1
echo "exit=$?" >> gpurun_out/k4_trace.log
