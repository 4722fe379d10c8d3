#!/bin/bash
# GPU tests of the narrow classes + timings of the operator configurations (engine and trainer)
O=gpurun_out; TAG=${1:-small}
timeout 600 python -m pytest tests/test_gpu_tpp.py tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > $O/${TAG}_test.log
for rep in 1 2; do
    timeout 200 python scripts/prof_small.py 2>&1 | cut -c1-240,400- >> $O/${TAG}_small.log
    timeout 200 python scripts/prof_cfg3.py 2>&1 | grep "^epoch" >> $O/${TAG}_cfg3.log
done
cat $O/${TAG}_test.log $O/${TAG}_small.log $O/${TAG}_cfg3.log
