#!/bin/bash
mkdir -p gpurun_out/r2cpu
g++ -O3 -std=c++17 -mavx512f -pthread -o /tmp/c2c scripts/c2c_probe.cpp
for mb in 1 2 8 32; do /tmp/c2c $mb; done | tee gpurun_out/r2cpu/c2c.txt
