#!/bin/bash
# first GPU pass: parity tests, smoke, functional bench, HVP probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python bench.py --small --steps 10 --warmup 2 --global-batch 16 > gpurun_out/bench_small.log 2>&1; echo "exit $?" >> gpurun_out/bench_small.log
timeout 600 python scripts/hvp_probe.py > gpurun_out/hvp_probe.log 2>&1; echo "exit $?" >> gpurun_out/hvp_probe.log
timeout 900 python bench.py --steps 20 --warmup 3 --global-batch 8 > gpurun_out/bench_gb8_k20.log 2>&1; echo "exit $?" >> gpurun_out/bench_gb8_k20.log
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_small.log gpurun_out/hvp_probe.log gpurun_out/bench_gb8_k20.log
