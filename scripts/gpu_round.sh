#!/bin/bash
# Runs under gpurun on the B200 box: GPU tests, sanitizer pass, bench, ncu launch list + one full capture.
# Every step has its own timeout so a hung kernel cannot eat the whole lease.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" | tee gpurun_out/pytest.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS:-} >> gpurun_out/pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -n 40 gpurun_out/pytest.log
if [ "${SKIP_SANITIZER:-0}" != "1" ]; then
  echo "== compute-sanitizer (memcheck, racecheck) on smoke"
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/racecheck.log 2>&1; echo "racecheck exit $?"
  tail -n 5 gpurun_out/memcheck.log gpurun_out/racecheck.log
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  for eng in ${ENGINES:-2 1}; do
    echo "== bench engine $eng"
    timeout 600 python bench.py --steps ${STEPS:-3} --warmup ${WARMUP:-3} --engine $eng ${BENCH_ARGS:-} > gpurun_out/bench_e$eng.json 2> gpurun_out/bench_e$eng.err
    echo "bench exit $?"; cat gpurun_out/bench_e$eng.json; tail -n 5 gpurun_out/bench_e$eng.err
  done
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  echo "== bench --impl reference (CPU arm: oracle port on the box's host cores)"
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  echo "reference exit $?"; cat gpurun_out/bench_reference.json
fi
if [ "${SKIP_MIP:-0}" != "1" ]; then
  echo "== mip_bench (1 GPU)"
  SPEC=${SPEC:-16,64} REPS=3 timeout 400 python scripts/mip_bench.py > gpurun_out/mip_bench.log 2>&1; echo "mip exit $?"; cut -c1-400 gpurun_out/mip_bench.log
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
  echo "ncu list exit $?"
  echo "== ncu full capture of k_pivot_step"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pivot_step -s 3000 -c 3 -f -o gpurun_out/prof_pivot_step \
     python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
ls -la gpurun_out
