#!/bin/bash
# r03f: full GPU parity suite (heuristic kernels, three-stream host pipe), chi-square report with per-sub-table p-values, bench
TAG=${1:-r03f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; POMDP_DIST_REPORT=$OUT/chisq.jsonl timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-300
tail -3 $OUT/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r03f/bench.json").read().strip().splitlines()[-1])
e = d["e2e"]; print("e2e", e["value"], e["ms_per_step"], e["ceiling_ms"], e["frac_of_ceiling"], {k: e[k]["ms_per_step"] for k in ("python_pipeline", "unpacked", "zero_copy") if e.get(k)})
PY
echo "== heuristic rollout timing"
python - <<'PY'
import torch, time, json
import gym_pomdp_b200 as gp
dev = torch.device("cuda", 0)
res = {}
for name, mk in (("Rock(11,11) heuristic", lambda B: gp.make("Rock-v0", board_size=11, num_rocks=11, use_heuristic=True, batch_size=B, device=dev, seed=1)),
                 ("Tag-v0 heuristic", lambda B: gp.make("Tag-v0", batch_size=B, device=dev, seed=1))):
    B = 1 << 20
    env = mk(B)
    s, _ = env.init_states(B, step_ctr=1)
    for pol in ("legal", "preferred"):
        for _ in range(2):
            out = env.rollout(s, max_steps=32, step_ctr=5, policy=pol)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = env.rollout(s, max_steps=32, step_ctr=5, policy=pol)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        n = int(out[2].sum())
        res["%s policy=%s" % (name, pol)] = {"ms_per_launch": ms, "env_steps": n, "env_steps_per_s": n / ms * 1e3}
print(json.dumps(res, indent=1))
open("gpurun_out/r03f/heuristic_rollouts.json", "w").write(json.dumps(res, indent=1))
PY
ls $OUT
