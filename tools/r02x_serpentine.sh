NCU="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --csv"
timeout 200 python tools/gemm_raster_sweep.py time gpurun_out/r02x_llama_time.jsonl > gpurun_out/r02x_llama_time.log 2>&1; tail -2 gpurun_out/r02x_llama_time.log
timeout 200 $NCU --log-file gpurun_out/r02x_llama_ncu.csv python tools/gemm_raster_sweep.py ncu > gpurun_out/r02x_ncu1.log 2>&1
python tools/gemm_raster_sweep.py join gpurun_out/r02x_llama_ncu.csv > gpurun_out/r02x_llama_traffic.jsonl
