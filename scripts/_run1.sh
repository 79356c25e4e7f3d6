set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "dynamic or model_level" > gpurun_out/pytest_dm.log 2>&1; echo "pytest dm exit $?"; tail -n 30 gpurun_out/pytest_dm.log
