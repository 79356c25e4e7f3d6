set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/pytest.log
python - <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
from jslpsolver_b200 import _lib
from jslpsolver_b200.tableau import default_context
ctx = default_context()
for mb in (8, 16, 24, 32, 40, 48, 64, 256):
    out = C.c_double()
    _lib.check(ctx.lib.jslp_debug_copy_gbs(ctx.handle, mb << 20, 50, C.byref(out)))
    print("copy ping-pong 2 x %d MB: %.0f GB/s" % (mb, out.value))
PY
