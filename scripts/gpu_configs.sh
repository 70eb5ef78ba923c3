#!/bin/bash
mkdir -p gpurun_out
for c in 1 3 5; do
  timeout 900 python scripts/run_configs.py --config $c > gpurun_out/config$c.json 2> gpurun_out/config$c.err; echo "config $c exit $?"
  tail -c 1500 gpurun_out/config$c.json; echo
  grep -v "loss_type" gpurun_out/config$c.err | tail -3 | cut -c1-300
done
