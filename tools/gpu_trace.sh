#!/bin/bash
mkdir -p gpurun_out
for v in g3 g4; do
OVN_TRACE_LIB=libovn_b200_trace_$v.so timeout 300 python tools/k4_trace.py > gpurun_out/k4_trace_$v.log 2>&1
echo "exit=$?" >> gpurun_out/k4_trace_$v.log
done
