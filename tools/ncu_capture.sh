#!/bin/bash
# ncu evidence for one round (run under gpurun, ONE GPU). Usage: tools/ncu_capture.sh r01 [launches_per_step]
# 1) plain runs must exit 0; 2) per-launch device times of one whole bs=256 step; 3) --set full of the top kernels;
# 4) bs=1 decode kernels (GEMV, fused decode attention) with CUDA graphs disabled so that ncu sees plain launches.
R=${1:-r01}
N=${2:-2367}
CMD="python bench.py --steps 1 --warmup 1 --lite"
mkdir -p gpurun_out
$CMD > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ovla:: -s $N -c $N --csv \
    --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
if [ "$NCU_LIGHT" != "2" ]; then
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 210 -c 4 \
    -o gpurun_out/${R}_gemm -f $CMD > gpurun_out/${R}_ncu_gemm.log 2>&1
fi
if [ -n "$NCU_LIGHT" ]; then
  ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 22 -c 3 \
      -o gpurun_out/${R}_attn_tc -f $CMD > gpurun_out/${R}_ncu_attn_tc.log 2>&1
  ls -la gpurun_out/ | grep ${R}_ | tail -20
  exit 0
fi
ncu --set full --clock-control none --import-source on -k regex:decode_rope_attn -s 4 -c 2 \
    -o gpurun_out/${R}_decode_attn -f $CMD > gpurun_out/${R}_ncu_decode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 10 -c 1 \
    -o gpurun_out/${R}_flash_attn -f $CMD > gpurun_out/${R}_ncu_flash.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 26 -c 2 \
    -o gpurun_out/${R}_attn_tc -f $CMD > gpurun_out/${R}_ncu_attn_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:norm_rows -s 120 -c 2 \
    -o gpurun_out/${R}_norm -f $CMD > gpurun_out/${R}_ncu_norm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pool_tokens -s 3 -c 1 \
    -o gpurun_out/${R}_pool -f $CMD > gpurun_out/${R}_ncu_pool.log 2>&1
export OVLA_GRAPHS=0
CMD1="python bench.py --batch 1 --steps 1 --warmup 1 --lite"
$CMD1 > gpurun_out/${R}_plain_bs1.log 2>&1 || { echo "bs1 plain run failed"; tail -5 gpurun_out/${R}_plain_bs1.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:ovla_wstream_kernel -s 200 -c 5 \
    -o gpurun_out/${R}_gemv_bs1 -f $CMD1 > gpurun_out/${R}_ncu_gemv.log 2>&1
ls -la gpurun_out/ | grep ${R}_ | tail -20
