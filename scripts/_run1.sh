set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/pytest.log
NO_FARM=0 NO_CPU=1 KNAP_NODES=1000 SPEC=32 REPS=2 timeout 300 python scripts/mip_bench.py 2>&1 | cut -c1-900 | tee gpurun_out/mip_1000.log
