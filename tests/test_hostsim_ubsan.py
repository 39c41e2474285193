"""UBSan over the kernel functors: the same ``__host__ __device__`` code the CUDA kernels inline
(pomdp_core.h / pomdp_envs.h: wrapping shifts, bit-field packing, the 128-bit bitboards, Philox) is compiled with
``g++ -fsanitize=undefined`` and driven through the parity, edge-case, rollout and query tests in a subprocess; any
"runtime error:" line (shift past the width, signed overflow, misaligned access ...) fails the test."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_functors_are_clean_under_ubsan(tmp_path):
    so = str(tmp_path / "libpomdp_hostsim_ubsan.so")
    src = os.path.join(ROOT, "tests", "hostsim", "pomdp_hostsim.cpp")
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fsanitize=undefined", "-shared", "-fPIC",
                            "-Wno-unknown-pragmas", "-o", so, src], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-2000:]
    assert "ubsan" in subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    env = dict(os.environ, POMDP_HOSTSIM_SO=so, UBSAN_OPTIONS="print_stacktrace=0")
    tests = ["tests/test_parity_golden.py", "tests/test_edge_cases.py", "tests/test_rollout.py", "tests/test_queries.py",
             "tests/test_rock_belief_stats.py", "tests/test_gym_surface.py"]
    run = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider"] + tests,
                         capture_output=True, text=True, cwd=ROOT, env=env, timeout=900)
    out = run.stdout + run.stderr
    assert run.returncode == 0, out[-3000:]
    assert "runtime error" not in out, [l for l in out.splitlines() if "runtime error" in l][:10]
