set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench exit $?"; cat gpurun_out/bench_r02b.json | cut -c1-3000; tail -3 gpurun_out/bench_r02b.err
