set -u
for gps in 1 2; do
GRID_PER_SM=$gps VARIANTS=1 SHAPES=dense300,dense600,dense800,dense1200,dense1400,dense1500 timeout 600 python scripts/variant_bench.py 2>&1 | python -c "
import sys, json
rows = [json.loads(l) for l in sys.stdin if l.startswith('{')]
best = {}
for r in rows: best[r['shape']] = min(best.get(r['shape'], 1e9), r['us_per_pivot'])
print('grid_per_sm', $gps, best)
"
done
