#!/bin/bash
# r03b: Network FMA-pipe decisions + reset fast paths: parity tests, variants, configs, ncu of the reworked kernels
TAG=${1:-r03b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== network variants"
bash scripts/exp_network_variants.sh $TAG "" "-DPOMDP_NET_FMA=0" "-DPOMDP_NET_UNROLL=2" "-DPOMDP_STEP_THREADS=256 -DPOMDP_STEP_MINB=4" \
   "-DPOMDP_NET_UNROLL=2 -DPOMDP_STEP_THREADS=256 -DPOMDP_STEP_MINB=3" "-DPOMDP_STEP_THREADS=1024 -DPOMDP_STEP_MINB=1" "-DPOMDP_NET_UNROLL=5"
echo "== configs"; timeout 900 python scripts/bench_configs.py --no-rollout --out $OUT/configs.json 2> $OUT/configs.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   %-48s %-12s %8.2f us  %.3f of peak' % (d['config'], d['kernel'], d['us_per_launch'], d['frac_of_peak']))
"
tail -3 $OUT/configs.err
for pair in "BattleShip 10x10 B=2^18:battleship" "Network-v0:network" "RockSample(11,11):rock11" "Tag-v0 B=2^20:tag"; do
  label="${pair%%:*}"; short="${pair##*:}"
  echo "== ncu full $short"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reset|NetworkEnv" -c 6 -f -o $OUT/$short \
      python scripts/bench_configs.py --quick --no-rollout --only "$label" > $OUT/ncu_$short.log 2>&1
done
ls -la $OUT
