#!/bin/bash
# development aid: rebuild the product library and the trace-instrumented copy
set -e
cd "$(dirname "$0")/../mppi_gpu_b200/csrc"
make 2>&1 | grep -v "^nvcc" | head -20
grep -A3 "step_kernelILi3ELb0" ptxas.log | grep -E "registers|spill" || true
mkdir -p ../../tools/_build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DMPPI_STEP_TRACE -shared \
  -o ../../tools/_build/libmppi_trace.so kernels.cu step.cu tile.cu controller.cu comm.cpp -ldl 2>&1 | grep -i -E "error" || true
