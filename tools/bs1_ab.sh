# bs=1 latency A/B of two library builds on one box (arms alternating): p50 of the bs1_latency block of bench.py
for i in 1 2 3; do
for arm in prev new; do
  if [ $arm = prev ]; then export OVLA_B200_LIB=openvla_probe_b200/libovla_b200_prev.so; else unset OVLA_B200_LIB; fi
  python bench.py --batch 1 --steps 30 --warmup 5 --no-cpu-baseline --no-probe --no-siglip 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$arm', 'ms_per_step', round(d['ms_per_step'],3), 'bs1', d.get('bs1_latency'))"
done; done
