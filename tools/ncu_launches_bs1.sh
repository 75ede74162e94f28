#!/bin/bash
export OVLA_GRAPHS=0
CMD1="python bench.py --batch 1 --steps 1 --warmup 1 --lite"
$CMD1 > gpurun_out/r02o_plain_bs1.log 2>&1 || { echo "bs1 plain run failed"; tail -5 gpurun_out/r02o_plain_bs1.log; exit 1; }
N=$(python -c "import json;print(json.load(open('gpurun_out/r02o_plain_bs1.log'))['gpu_launches'])")
echo "launches per step: $N"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ovla:: -s $N -c $N --csv \
    --log-file gpurun_out/r02o_launches_bs1.csv $CMD1 > gpurun_out/r02o_ncu_launches_bs1.log 2>&1
wc -l gpurun_out/r02o_launches_bs1.csv
