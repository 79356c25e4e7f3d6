#!/bin/bash
# Multi-GPU round (run under `gpurun --gpus 8`): bit-exactness of the sharded frontier over the in-library NCCL
# communicator, then bench.py (LP replicas + MIP block) and the MIP node-throughput report at 1/2/4/8 GPUs.
# The 1-, 2- and 4-GPU runs use disjoint devices and run concurrently; every step has its own timeout.
set -u
mkdir -p gpurun_out/multi
O=gpurun_out/multi
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export NO_CPU=1 REPS=${REPS:-2} SPEC_PER_RANK=${SPEC_PER_RANK:-32} KNAP_NODES=${KNAP_NODES:-1000} SPIN_S=0.2
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/gpus.txt 2>&1
echo "== dist_check N=8 (in-library NCCL)"
timeout 400 $TR --nproc-per-node 8 --master-port 29701 scripts/dist_check.py > $O/dist_check_n8.log 2>&1; echo "exit $?"; tail -n 4 $O/dist_check_n8.log
echo "== bench.py N=1,2,4 concurrently on disjoint GPUs"
( CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err ) &
( CUDA_VISIBLE_DEVICES=1,2 timeout 400 $TR --nproc-per-node 2 --master-port 29702 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > $O/bench_n2.json 2> $O/bench_n2.err ) &
( CUDA_VISIBLE_DEVICES=3,4,5,6 timeout 400 $TR --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu > $O/bench_n4.json 2> $O/bench_n4.err ) &
wait
echo "== bench.py N=8"
timeout 400 $TR --nproc-per-node 8 --master-port 29708 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > $O/bench_n8.json 2> $O/bench_n8.err; echo "exit $?"
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        line = [l for l in open(f"gpurun_out/multi/bench_n{n}.json") if l.startswith("{")][-1]
        r = json.loads(line); m = r["mip"]
        print(n, "LP pivots/s %.0f" % r["value"], "| MIP total_ms %.0f root %.0f node_phase %.0f node LPs/s %.0f committed/s %.0f node_lps %d rounds %d K %d slot frac %.3f coll %d" % (
            m["total_ms"], m["root_lp_ms"], m["node_phase_ms"], m["node_lps_per_s"], m["committed_per_s"], m["node_lps"], m["rounds"], m["spec_width"], m["roofline"]["frac"] or 0, m["collectives_per_rank"]))
    except Exception as e:
        print(n, "failed", e)
PY
if [ "${WITH_FARM:-1}" = "1" ]; then
  echo "== LargeFarm at N=1 and N=8 (auto shard policy: rounds of shared-memory nodes are not sharded)"
  NO_KNAP=1 SPEC_PER_RANK= SPEC=32 CUDA_VISIBLE_DEVICES=0 timeout 200 python scripts/mip_bench.py 2>/dev/null | grep '"impl": "b200"' | cut -c1-330
  NO_KNAP=1 SPEC_PER_RANK= SPEC=32 timeout 200 $TR --nproc-per-node 8 --master-port 29711 scripts/mip_bench.py 2>/dev/null | grep '"impl": "b200"' | cut -c1-330
fi
