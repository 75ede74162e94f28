# same-box A/B of producer variants: prev = HEAD~ build, new = this tree, lead1 = arrival one K block early
for i in 1 2; do
OVLA_B200_LIB=openvla_probe_b200/libovla_b200_prev.so python tools/gemm_sustained_bench.py child | sed "s/^/prev  /"
python tools/gemm_sustained_bench.py child | sed "s/^/new   /"
[ -f openvla_probe_b200/libovla_b200_lead1.so ] && OVLA_B200_LIB=openvla_probe_b200/libovla_b200_lead1.so python tools/gemm_sustained_bench.py child | sed "s/^/lead1 /"
done
OVLA_B200_LIB=openvla_probe_b200/libovla_b200_prev.so python tools/gemm_epilogue_bench.py | sed "s/^/prev /" | cut -c1-700
python tools/gemm_epilogue_bench.py | sed "s/^/new  /" | cut -c1-700
python -m pytest tests/test_gpu_operators.py -m gpu -x -q 2>&1 | tail -2
