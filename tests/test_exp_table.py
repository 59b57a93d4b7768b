"""The device exp (abx_core.cuh: exp_ni) restated on the host instruction for instruction (tools/check_exp.c, every fp64 step an explicit fma / add / mul, the same
table header): maximum error against long-double expl and agreement with the C library's exp, which is what the oracle calls.  The GPU suite checks the device
function itself against libm (tests/test_gpu_philox.py::test_device_exp_against_libm)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_table_exp_is_within_an_ulp_and_agrees_with_libm(tmp_path):
    exe = str(tmp_path / "check_exp")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-I", os.path.join(ROOT, "marl_optimal_execution_b200", "csrc"), "-o", exe, os.path.join(ROOT, "tools", "check_exp.c"), "-lm"])
    r = subprocess.run([exe, "4000000"], capture_output=True, text=True)
    if r.returncode < 0:                                   # a host without FMA instructions cannot run the restatement
        pytest.skip("host cannot execute -mfma code")
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"max error ([0-9.]+) ulp; differs from glibc exp on (\d+) of (\d+)", r.stdout)
    assert m, r.stdout
    assert float(m.group(1)) < 0.52
    assert int(m.group(2)) / int(m.group(3)) < 2e-3


def test_table_header_is_what_the_generator_writes(tmp_path):
    """abx_exp_table.h is generated (tools/gen_exp_table.py, 60-digit decimal arithmetic); the committed header must be its output."""
    committed = open(os.path.join(ROOT, "marl_optimal_execution_b200", "csrc", "abx_exp_table.h")).read()
    import shutil
    work = tmp_path / "w"
    (work / "tools").mkdir(parents=True)
    (work / "marl_optimal_execution_b200" / "csrc").mkdir(parents=True)
    shutil.copy(os.path.join(ROOT, "tools", "gen_exp_table.py"), work / "tools" / "gen_exp_table.py")
    subprocess.check_call(["python", str(work / "tools" / "gen_exp_table.py")], stdout=subprocess.DEVNULL)
    assert open(work / "marl_optimal_execution_b200" / "csrc" / "abx_exp_table.h").read() == committed
