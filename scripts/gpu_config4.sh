#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
  scripts/run_configs.py --config 4 > gpurun_out/config4_n$N.json 2> gpurun_out/config4_n$N.err; echo "config 4 exit $?"
tail -c 1500 gpurun_out/config4_n$N.json; echo
grep -v "loss_type\|OMP_NUM\|\*\*\*\*" gpurun_out/config4_n$N.err | tail -5 | cut -c1-300
