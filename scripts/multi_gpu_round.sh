#!/bin/bash
# Multi-GPU round (run under `gpurun --gpus 8`): bit-exactness of the sharded frontier over NCCL, then
# MIP node throughput at 1/2/4/8 GPUs.  The 1-, 2- and 4-GPU runs use disjoint devices and run
# concurrently; every step has its own timeout.
set -u
mkdir -p gpurun_out/multi
O=gpurun_out/multi
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export NO_CPU=1 REPS=${REPS:-2} SPEC=${SPEC:-64} KNAP_NODES=${KNAP_NODES:-400} SPIN_S=0.2
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/gpus.txt 2>&1
echo "== dist_check N=8 (nccl)"; 
timeout 300 $TR --nproc-per-node 8 --master-port 29701 scripts/dist_check.py > $O/dist_check_n8.log 2>&1; echo "exit $?"; tail -n 3 $O/dist_check_n8.log
echo "== mip_bench N=1,2,4 concurrently on disjoint GPUs"
( CUDA_VISIBLE_DEVICES=0 timeout 300 python scripts/mip_bench.py > $O/mip_n1.log 2>&1 ) &
( CUDA_VISIBLE_DEVICES=1,2 timeout 300 $TR --nproc-per-node 2 --master-port 29702 scripts/mip_bench.py > $O/mip_n2.log 2>&1 ) &
( CUDA_VISIBLE_DEVICES=3,4,5,6 timeout 300 $TR --nproc-per-node 4 --master-port 29704 scripts/mip_bench.py > $O/mip_n4.log 2>&1 ) &
wait
echo "== mip_bench N=8"
timeout 300 $TR --nproc-per-node 8 --master-port 29708 scripts/mip_bench.py > $O/mip_n8.log 2>&1; echo "exit $?"
grep -h '"impl": "b200"' $O/mip_n*.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print(r['n_gpus'], r['workload'][:28], 'K', r['spec_width'], 'gpu_ms %.1f' % r['gpu_ms'], 'root %.1f' % r['host_root_ms'], 'node_phase %.1f' % r['node_phase_ms'], 'iters', r['iterations'], 'lps', r['node_lps_all_ranks'], 'node-phase LPs/s %.0f' % (r['node_phase_node_lps_per_s'] or 0), 'committed/s %.0f' % (r['node_phase_committed_per_s'] or 0))
"
if [ "${WITH_BENCH:-1}" = "1" ]; then
  echo "== bench.py N=8 replicas"
  timeout 300 $TR --nproc-per-node 8 --master-port 29709 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > $O/bench_n8.json 2> $O/bench_n8.err; echo "exit $?"; cat $O/bench_n8.json
fi
