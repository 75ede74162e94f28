for gm in 16 64 16 64; do
OVLA_GRAPH_MAX_BATCH=$gm python bench.py --config siglip-7b --batch 64 --steps 6 --warmup 4 --no-cpu-baseline --no-bs1 --no-probe --no-siglip 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('siglip bs64 graph_max_batch=$gm', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2))"
done
