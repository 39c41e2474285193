"""The belief histogram fused with its all-reduce (pomdp_belief_hist_allreduce) through the public call,
``belief_histogram(all_reduce="fused")``: a one-rank torch.distributed world on the test GPU (torch symmetric memory for
the peer table), eager and replayed from a CUDA graph, against the NCCL form and the plain histogram.  The kernel's
multi-peer behaviour is covered on one device in test_edge_cases.py; real multi-GPU equality is part of the bench line
(``collective.fused.equals_nccl``) under torchrun."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_fused_histogram_through_the_public_call_in_a_one_rank_world():
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "bench_fused_hist.py"), "--log2-global", "18"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["all_variants_equal"] is True and out["bins"] == 271
    t = out["us_per_call_max_over_ranks"]
    assert set(t) == {"hist", "nccl", "fused"} and all("graph_us" in v for v in t.values()), t
    # the step with the histogram (and the all-reduce) in its epilogue: pomdp_rock_step_hist, eager and graph-replayed
    p = out["step_pipeline_us_max_over_ranks"]
    assert {"step", "step_then_hist_nccl", "step_then_hist_fused", "step_hist_local", "step_hist_fused"} == set(p)
    assert all("graph_us" in v for v in p.values()), p
