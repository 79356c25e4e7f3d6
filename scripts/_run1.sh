set -u
mkdir -p gpurun_out
for i in 1 2; do
JSLP_LIB=scripts/ab/libjslp_r01.so VARIANTS=1 SHAPES=dense2000 python scripts/variant_bench.py 2>&1 | grep -v Warn | sed 's/^/r01 /'
VARIANTS=1 SHAPES=dense2000 python scripts/variant_bench.py 2>&1 | sed 's/^/r02 /'
done
python - <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
from jslpsolver_b200 import _lib
from jslpsolver_b200.tableau import default_context
ctx = default_context()
for mb in (8, 16, 32, 48, 64, 128, 512):
    out = C.c_double()
    _lib.check(ctx.lib.jslp_debug_copy_gbs(ctx.handle, mb << 20, 50, C.byref(out)))
    print("copy ping-pong 2 x %d MB: %.0f GB/s" % (mb, out.value))
PY
echo "== ncu application replay, warm cache"
timeout 900 ncu --replay-mode application --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread \
   -k regex:k_pivot_step -s 4000 -c 6 -f -o gpurun_out/prof_pivot_step_apprep python bench.py --steps 1 --warmup 1 --no-cpu --mip-nodes 0 > gpurun_out/ncu_apprep.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/prof_pivot_step_apprep.ncu-rep --page raw --csv > gpurun_out/pp_apprep_raw.csv 2> /dev/null
python scripts/ncu_summary.py gpurun_out/pp_apprep_raw.csv gpurun_out/r02_k_pivot_step_ncu "ncu --replay-mode application --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:k_pivot_step -s 4000 -c 6 python bench.py --steps 1 --warmup 1 --no-cpu --mip-nodes 0"
tail -3 gpurun_out/ncu_apprep.log
