"""The N-API addon (binding/jslp_addon.cc) EXECUTED on the GPU without Node.js: built against tests/stubs/node_api.h and
linked with an in-process emulation of the N-API calls it makes (tests/stubs/napi_emul.cc) and with libjslp_b200.so;
tests/cpp/addon_emul_test.cc then drives every addon method the way gpu-tableau.ts does."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_addon_runs_through_napi_emulation(tmp_path):
    from jslpsolver_b200 import _lib
    _lib.load()
    libdir = os.path.join(ROOT, "jslpsolver_b200")
    exe = tmp_path / "addon_emul_test"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
                    "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "addon_emul_test.cc"), os.path.join(ROOT, "binding", "jslp_addon.cc"),
                    os.path.join(ROOT, "tests", "stubs", "napi_emul.cc"), "-L", libdir, "-ljslp_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ADDON EMUL OK" in out.stdout, out.stdout + out.stderr
