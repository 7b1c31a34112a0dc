"""Memory safety of the solver text: the single-lane host build (tests/hostsim) compiled with AddressSanitizer runs every
path - ADMM + polish, warm polish, the interior-point fallback, the robust chain, the quadruped, the largest and the
smallest tree - and must finish without a report.  The same indexing runs on the device (compute-sanitizer is not
available on the GPU pool)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_solver_paths_are_clean_under_address_sanitizer(tmp_path):
    gxx = shutil.which("g++")
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not gxx or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("g++ with libasan is not available")
    lib = str(tmp_path / "libhostsim_asan.so")
    subprocess.check_call([gxx, "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-std=c++17", "-shared", "-fPIC",
                           "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "belief-planning_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp"), "-o", lib])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "hostsim", "asan_run.py"), lib], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ASAN_RUN_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    assert "AddressSanitizer" not in r.stderr
