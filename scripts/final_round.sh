#!/bin/bash
# One GPU call for the end-of-round record: GPU tests, the bench line (both arms), launch list, full ncu capture of the dominant
# kernel, DRAM traffic of one cfg-4 launch, small-config timings.  Everything goes to gpurun_out/${TAG}_*.
TAG=${1:-r2f}
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -60 > $O/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-width-sweep --e2e-steps 1 > $O/${TAG}_l.log 2>&1
AB_NTF=37888 timeout 400 ncu --set full --import-source on --clock-control none -k regex:tc64_var_kernel -s 4 -c 1 -o $O/${TAG}_tc64 -f \
    python scripts/ab_tc64.py --child final > $O/${TAG}_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:tc64_var -s 5 -c 1 --csv --log-file $O/${TAG}_traffic.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-width-sweep --no-operator-configs --e2e-steps 1 > $O/${TAG}_t.log 2>&1
timeout 200 python scripts/prof_small.py > $O/${TAG}_small.log 2>&1
timeout 200 python scripts/prof_cfg3.py 2>&1 | grep "^epoch" > $O/${TAG}_cfg3.log
timeout 100 scripts/micro/tmem_overlap > $O/${TAG}_overlap.log 2>&1
tail -3 $O/${TAG}_pytest.log; head -c 400 $O/${TAG}_bench.json; echo; tail -2 $O/${TAG}_bench.err; cat $O/${TAG}_cfg3.log; cat $O/${TAG}_overlap.log
