set -u
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "enhanced_service" > gpurun_out/pytest_inc.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_inc.log
