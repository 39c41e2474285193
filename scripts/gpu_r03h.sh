#!/bin/bash
TAG=${1:-r03h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest battleship"; timeout 900 python -m pytest tests -m gpu -q -x -k "battleship or ship or odd_boards or fullsize" 2>&1 | tail -4
echo "== configs BattleShip"; python scripts/bench_configs.py --no-rollout --only "BattleShip" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   %-34s %-12s %8.2f us  %.3f of peak' % (d['config'], d['kernel'], d['us_per_launch'], d['frac_of_peak']))
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reset" -c 3 -f -o $OUT/battleship_reset \
      python scripts/bench_configs.py --quick --no-rollout --only "BattleShip 10x10 B=2^18" > $OUT/ncu.log 2>&1
ls $OUT
