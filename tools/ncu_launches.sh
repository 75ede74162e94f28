#!/bin/bash
# per-launch device times of one whole step (after one warm-up step); only kernels of namespace ovla::
R=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --lite"
$CMD > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ovla:: -s 2367 -c 2367 --csv \
    --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
wc -l gpurun_out/${R}_launches.csv
