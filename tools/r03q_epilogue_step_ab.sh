# whole-step effect of the epilogue work: same box, arms alternating; "before" = the library of commit 79c01c5
# (wave-aligned GEMMs, serpentine) built as libovla_b200_wavealign.so, "after" = this tree
for i in 1 2; do
for arm in before after; do
  if [ $arm = before ]; then export OVLA_B200_LIB=openvla_probe_b200/libovla_b200_wavealign.so; else unset OVLA_B200_LIB; fi
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bs1 --no-probe --no-siglip 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({'arm':'$arm','actions_per_s':round(d['value'],2),'ms_per_step':round(d['ms_per_step'],1),'gemm_ms':round(d['kernel_breakdown']['gemm_tcgen05']['ms_per_step'],1),'gemm_frac':round(d['roofline']['frac'],4),'sm_mhz':d['clocks']['sm_mhz']}))"
done; done
