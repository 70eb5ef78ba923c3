#!/bin/bash
# fused-kernel iteration: parity of the fused pass, then the kernel table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "fused or cgs" > gpurun_out/pytest_fused.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fused.log
tail -4 gpurun_out/pytest_fused.log | cut -c1-300
timeout 300 python scripts/cgs_bench.py f32 8,16,25,26,40,50,51,75,100 > gpurun_out/cgs_bench_f32.log 2>&1
timeout 300 python scripts/cgs_bench.py bf16 8,25,50,100,200 > gpurun_out/cgs_bench_bf16.log 2>&1
grep torch gpurun_out/cgs_bench_f32.log gpurun_out/cgs_bench_bf16.log | cut -c1-330
