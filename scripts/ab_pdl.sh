#!/bin/bash
# A/B of the programmatic-dependency edges in the k-step graphs of the narrow operator configurations (same box, back to back).
O=gpurun_out; TAG=${1:-pdl}
timeout 600 python -m pytest tests/test_gpu_tpp.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -5 > $O/${TAG}_test.log
for rep in 1 2; do
  for v in 1 0; do
    echo "== VARNET_B200_PDL=$v" >> $O/${TAG}_small.log
    VARNET_B200_PDL=$v timeout 200 python scripts/prof_small.py 2>&1 | cut -c1-120,400- >> $O/${TAG}_small.log
    echo "== VARNET_B200_PDL=$v" >> $O/${TAG}_cfg3.log
    VARNET_B200_PDL=$v timeout 200 python scripts/prof_cfg3.py 2>&1 | grep "^epoch" >> $O/${TAG}_cfg3.log
  done
done
cat $O/${TAG}_test.log $O/${TAG}_small.log $O/${TAG}_cfg3.log
