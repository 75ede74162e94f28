#!/bin/bash
# ncu --set full of the GEMMs of the final round-2 build (run under gpurun, ONE GPU): tools/ncu_capture_r03.sh r03k
#  one Llama layer's four GEMMs (launches 209..212 of a bs=256 step) and one ViT block per tower (1..4, 94..97)
R=${1:-r03k}
CMD="python bench.py --steps 1 --warmup 1 --lite"
mkdir -p gpurun_out
$CMD > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 209 -c 4 \
    -o gpurun_out/${R}_gemm -f $CMD > gpurun_out/${R}_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 4 \
    -o gpurun_out/${R}_gemm_dino -f $CMD > gpurun_out/${R}_ncu_gemm_dino.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 94 -c 4 \
    -o gpurun_out/${R}_gemm_siglip -f $CMD > gpurun_out/${R}_ncu_gemm_siglip.log 2>&1
ls -la gpurun_out/ | grep ${R}_ | tail
