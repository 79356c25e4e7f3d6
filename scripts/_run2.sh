set -u
mkdir -p gpurun_out/multi
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 $TR --nproc-per-node 2 --master-port 29701 scripts/dist_check.py > gpurun_out/multi/dist_check_n2.log 2>&1; echo "dist_check exit $?"; tail -n 18 gpurun_out/multi/dist_check_n2.log
NO_CPU=1 KNAP_NODES=1000 SPEC=64 REPS=2 timeout 300 $TR --nproc-per-node 2 --master-port 29702 scripts/mip_bench.py 2>&1 | grep impl | cut -c1-1100 | tee gpurun_out/multi/mip_n2.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
