set -u
NO_KNAP=1 NO_CPU=1 SPEC=8,16,32,64 REPS=5 timeout 300 python scripts/mip_bench.py 2>/dev/null | grep impl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('K', r['spec_width'], 'gpu_ms %.2f' % r['gpu_ms'], 'wall %.2f' % r['wall_ms'], 'rounds', r['rounds'], 'lps', r['node_lps'], 'eval %.2f commit %.2f kernel %.2f' % (r['host_eval_ms'], r['host_commit_ms'], r['node_kernel_ms']))
"
timeout 900 python -m pytest tests/test_gpu_lp_parity.py -m gpu -q -x -p no:cacheprovider -k "node_sequence or timeout or keep_solutions" 2>&1 | tail -3
