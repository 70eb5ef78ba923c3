#!/bin/bash
# fused-kernel iteration: parity of the fused pass, then the kernel table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "fused or cgs" > gpurun_out/pytest_fused.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fused.log
tail -4 gpurun_out/pytest_fused.log | cut -c1-300
timeout 300 python scripts/cgs_bench.py f32 2,4,6,8,12,16,24,25,32,48,49,56,64,75,88,100 > gpurun_out/cgs_bench_f32.log 2>&1
timeout 300 python scripts/cgs_bench.py bf16 4,8,16,25,50,100,200 > gpurun_out/cgs_bench_bf16.log 2>&1
grep torch gpurun_out/cgs_bench_f32.log gpurun_out/cgs_bench_bf16.log | cut -c1-330
