#!/bin/bash
# ncu evidence of the wave-aligned GEMM build (run under gpurun, ONE GPU): tools/ncu_capture_r02w.sh r02w
#  1) the plain command must exit 0;  2) per-launch device times of one whole bs=256 step (launch list);
#  3) --set full of one Llama layer's four GEMMs (qkv+rope, o_proj, gate_up, down_proj with the fused-norm epilogues).
R=${1:-r02w}
CMD="python bench.py --steps 1 --warmup 1 --lite"
mkdir -p gpurun_out
$CMD > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.log; exit 1; }
N=$(python -c "import json;print(json.load(open('gpurun_out/${R}_plain.log'))['gpu_launches'])")
echo "launches per step: $N"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ovla:: -s $N -c $N --csv \
    --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 209 -c 4 \
    -o gpurun_out/${R}_gemm -f $CMD > gpurun_out/${R}_ncu_gemm.log 2>&1
ls -la gpurun_out/ | grep ${R}_ | tail -20
