#!/bin/bash
# r04s: histogram bit bins by carry-save addition -- parity + timings
OUT=gpurun_out/r04s; mkdir -p $OUT
timeout 900 python -m pytest tests/test_edge_cases.py tests/test_fullsize_parity.py tests/test_fused_collective.py tests/test_example_particle_filter.py -m gpu -q -x -k "hist or fused or particle or carry" 2>&1 | tail -3 | tee $OUT/pytest.log
python scripts/bench_configs.py --kernels belief_hist --out $OUT/configs.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('%-50s %-12s %8.2f us  %.3f of peak' % (d['config'], d['kernel'], d['us_per_launch'], d['frac_of_peak']))
" | tee $OUT/configs.log
