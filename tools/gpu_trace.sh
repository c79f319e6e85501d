#!/bin/bash
# clock64 timeline of k_delta_conv1_tc: needs the trace build
#   cd overlapnet_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
#      -DOVN_K4_TRACE -shared -o ../libovn_b200_trace.so api.cu projection.cu gt_overlap.cu network_fp32.cu network_tc.cu
mkdir -p gpurun_out
timeout 300 python tools/k4_trace.py > gpurun_out/k4_trace.log 2>&1
echo "exit=$?" >> gpurun_out/k4_trace.log
