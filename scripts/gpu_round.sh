#!/bin/bash
# Runs under gpurun on the B200 box: GPU tests, smoke, sanitizer pass, bench (both arms), config report, ncu launch list +
# warm-cache application-replay capture of k_pivot_step (summarised by scripts/ncu_summary.py / summarize_profiles.py).
# Every step has its own timeout so a hung kernel cannot eat the whole lease.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" | tee gpurun_out/pytest.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS:-} >> gpurun_out/pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest.log
tail -n 6 gpurun_out/pytest.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
if [ "${SKIP_SANITIZER:-0}" != "1" ]; then
  echo "== compute-sanitizer (memcheck, racecheck) on the slot batch / MIR / dynamic-modification paths"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_lp_parity.py -m gpu -q -x -p no:cacheprovider \
     -k "(hbm_slots3 and Knapsack) or (mir_primitives) or (dynamic_modification_sequence and 31)" > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_lp_parity.py -m gpu -q -x -p no:cacheprovider \
     -k "hbm_slots3 and Knapsack" > gpurun_out/racecheck.log 2>&1; echo "racecheck exit $?"
  tail -n 2 gpurun_out/memcheck.log gpurun_out/racecheck.log
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  echo "== bench"
  timeout 900 python bench.py --steps ${STEPS:-20} --warmup ${WARMUP:-5} ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"; cut -c1-600 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
  echo "== bench --impl reference (CPU arm: oracle port on the box's host cores)"
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  echo "reference exit $?"; cut -c1-300 gpurun_out/bench_reference.json
  echo "== configs 1 and 2"
  timeout 300 python scripts/config_report.py > gpurun_out/config_report.log 2>&1; echo "config_report exit $?"; cut -c1-400 gpurun_out/config_report.log
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "== ncu launch list (caches left as they are)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 4000 -c 300 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu --mip-nodes 0 > gpurun_out/ncu_bench.log 2>&1
  echo "ncu list exit $?"
  echo "== ncu application replay of k_pivot_step mid-solve (natural cache state)"
  timeout 900 ncu --replay-mode application --cache-control none --clock-control none \
     --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread \
     -k regex:k_pivot_step -s 4000 -c 6 -f -o gpurun_out/prof_pivot_step_apprep python bench.py --steps 1 --warmup 1 --no-cpu --mip-nodes 0 > gpurun_out/ncu_apprep.log 2>&1
  echo "ncu apprep exit $?"
  ncu -i gpurun_out/prof_pivot_step_apprep.ncu-rep --page raw --csv > gpurun_out/pp_apprep_raw.csv 2> /dev/null
  python scripts/ncu_summary.py gpurun_out/pp_apprep_raw.csv gpurun_out/r02_k_pivot_step_ncu "ncu --replay-mode application --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:k_pivot_step -s 4000 -c 6 python bench.py --steps 1 --warmup 1 --no-cpu --mip-nodes 0"
fi
ls gpurun_out | head -40
