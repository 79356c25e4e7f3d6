set -u
mkdir -p gpurun_out
TOLS=0.05,0.03,0.02,0.015,0.01 CAP=6000 SPEC=32 timeout 900 python scripts/knap_explore.py 2>&1 | tee gpurun_out/knap_explore.log
VARIANTS=1,10,11 SHAPES=dense3000 timeout 300 python scripts/variant_bench.py 2>&1 | tee gpurun_out/variant_bench3.log
