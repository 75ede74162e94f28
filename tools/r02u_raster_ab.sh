set -x
NCU="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --csv"
timeout 200 python tools/gemm_raster_sweep.py time gpurun_out/r02u_llama_time.jsonl > gpurun_out/r02u_llama_time.log 2>&1; tail -2 gpurun_out/r02u_llama_time.log
RASTER_SET=vit timeout 250 python tools/gemm_raster_sweep.py time gpurun_out/r02u_vit_time.jsonl > gpurun_out/r02u_vit_time.log 2>&1; tail -2 gpurun_out/r02u_vit_time.log
timeout 200 $NCU --log-file gpurun_out/r02u_llama_ncu.csv python tools/gemm_raster_sweep.py ncu > gpurun_out/r02u_ncu1.log 2>&1
python tools/gemm_raster_sweep.py join gpurun_out/r02u_llama_ncu.csv > gpurun_out/r02u_llama_traffic.jsonl
RASTER_SET=vit timeout 200 $NCU --log-file gpurun_out/r02u_vit_ncu.csv python tools/gemm_raster_sweep.py ncu > gpurun_out/r02u_ncu2.log 2>&1
RASTER_SET=vit python tools/gemm_raster_sweep.py join gpurun_out/r02u_vit_ncu.csv > gpurun_out/r02u_vit_traffic.jsonl
for i in 1 2; do
OVLA_GEMM_WAVESYNC=0 OVLA_GEMM_GROUP_N=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bs1 --no-probe --no-siglip > gpurun_out/r02u_bench_off$i.json 2> gpurun_out/r02u_bench_off$i.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bs1 --no-probe --no-siglip > gpurun_out/r02u_bench_on$i.json 2> gpurun_out/r02u_bench_on$i.err
done
python - <<'PY'
import json
for n in ("off1","on1","off2","on2"):
    try:
        d=json.load(open(f"gpurun_out/r02u_bench_{n}.json"))
        print(n, round(d["value"],2), round(d["ms_per_step"],1), {k:round(v["ms_per_step"],1) for k,v in d["kernel_breakdown"].items()}, round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"])
    except Exception as e: print(n, "failed", e)
PY
