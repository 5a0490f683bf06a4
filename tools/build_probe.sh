#!/bin/bash
# builds tools/bin/tc_probe (B200 microbenchmarks / descriptor validation; not product code)
set -e
cd "$(dirname "$0")"
mkdir -p bin
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -o bin/tc_probe tc_probe.cu
echo built tools/bin/tc_probe
