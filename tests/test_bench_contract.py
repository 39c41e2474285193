"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm really runs here
(bounded sample), and the last committed B200 line carries every key the contract names."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_runs_on_host_cores():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"].startswith("env-steps/sec RockSample(11,11) batch=2^22") and line["unit"] == "env-steps/s"
    assert line["value"] > 1e3 and line["higher_is_better"] is True and line["scaling"] == "weak"
    cb = line["cpu_baseline"]
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    # the unmodified reference's own step() loop wherever its package is reachable (/root/reference here, oracle/_ref on
    # the GPU box), the oracle's port of it otherwise
    if ref_shim.reference_available():
        assert cb["kind"] == "reference" and "RockEnv._set_state" in cb["sample"]
        assert cb["single_core"]["cores"] == 1 and cb["single_core"]["kind"] == "reference" and cb["single_core"]["value"] > 1e3
        assert cb["python_port"]["kind"] == "port"
    else:
        assert cb["kind"] == "port"
    assert cb["cores"] == os.cpu_count() and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("RockSample(11,11) step()")


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU path" in (out.stderr + out.stdout)


def test_committed_b200_line_has_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench.json")))
    assert files, "no committed bench line under profiles/"
    line = json.loads(open(files[-1]).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches"} <= set(line)
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - line["value"] / line["n_gpus"] * 24 / 1e9) / r["achieved"] < 1e-6      # 24 B per env-step
    assert line["gpu_launches"] == line["steps"] and line["vs_baseline"] is None
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < line["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0
