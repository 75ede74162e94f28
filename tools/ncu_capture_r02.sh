#!/bin/bash
# ncu evidence of round 2 (run under gpurun, ONE GPU): tools/ncu_capture_r02.sh r02j
#  1) the plain command must exit 0;  2) per-launch device times of one whole bs=256 step (launch list);
#  3) --set full of one Llama layer's four GEMMs (fused-norm epilogues) and of the ViT GEMMs of one block per tower;
#  4) the persistent bs=1 decode-step kernel (CUDA graphs off so that ncu sees the launch).
R=${1:-r02j}
CMD="python bench.py --steps 1 --warmup 1 --lite"
mkdir -p gpurun_out
$CMD > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.log; exit 1; }
N=$(python -c "import json;print(json.load(open('gpurun_out/${R}_plain.log'))['gpu_launches'])")
echo "launches per step: $N"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ovla:: -s $N -c $N --csv \
    --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 209 -c 4 \
    -o gpurun_out/${R}_gemm -f $CMD > gpurun_out/${R}_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 4 \
    -o gpurun_out/${R}_gemm_dino -f $CMD > gpurun_out/${R}_ncu_gemm_dino.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 94 -c 4 \
    -o gpurun_out/${R}_gemm_siglip -f $CMD > gpurun_out/${R}_ncu_gemm_siglip.log 2>&1
export OVLA_GRAPHS=0
CMD1="python bench.py --batch 1 --steps 1 --warmup 1 --lite"
$CMD1 > gpurun_out/${R}_plain_bs1.log 2>&1 || { echo "bs1 plain run failed"; tail -5 gpurun_out/${R}_plain_bs1.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 6 -c 1 \
    -o gpurun_out/${R}_decode_mega -f $CMD1 > gpurun_out/${R}_ncu_mega.log 2>&1
ls -la gpurun_out/ | grep ${R}_ | tail -20
