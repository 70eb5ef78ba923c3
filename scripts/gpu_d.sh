#!/bin/bash
# 2-GPU pass: NCCL invariance test, 2-rank bench, CUDA-graph probe
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_lanczos.py -m gpu -q -k multi_gpu --timeout 600 > gpurun_out/pytest_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_k20.json 2> gpurun_out/bench_n2_k20.err; echo "exit $?" >> gpurun_out/bench_n2_k20.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 3 --no-e2e > gpurun_out/bench_n2_k100.json 2> gpurun_out/bench_n2_k100.err; echo "exit $?" >> gpurun_out/bench_n2_k100.err
timeout 600 python scripts/hvp_probe.py 1,4 > gpurun_out/hvp_probe2.log 2>&1
tail -3 gpurun_out/pytest_2gpu.log; tail -3 gpurun_out/bench_n2_k20.err; cat gpurun_out/bench_n2_k20.json | cut -c1-3000; tail -2 gpurun_out/bench_n2_k100.err; cat gpurun_out/bench_n2_k100.json | cut -c1-1500; grep -v loss_type gpurun_out/hvp_probe2.log
