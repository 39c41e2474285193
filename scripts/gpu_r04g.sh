#!/bin/bash
# r04g: per-action-class step timings + the belief-histogram rows after skipping unused counter words (with parity)
OUT=gpurun_out/r04g; mkdir -p $OUT
timeout 600 python -m pytest tests/test_edge_cases.py tests/test_fullsize_parity.py -m gpu -q -x -k "hist" 2>&1 | tail -2 | tee $OUT/pytest.log
python scripts/bench_action_classes.py --out $OUT/action_classes.json 2>&1 | cut -c1-200 | tee $OUT/action_classes.log
python scripts/bench_configs.py --kernels belief_hist --out $OUT/configs_hist.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('%-50s %-12s %8.2f us  %.3f of peak' % (d['config'], d['kernel'], d['us_per_launch'], d['frac_of_peak']))
" | tee $OUT/configs_hist.log
