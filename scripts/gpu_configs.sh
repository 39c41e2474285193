#!/bin/bash
# One gpurun call: per-config kernel timings + ncu full captures of the non-Rock kernels.
# Usage:  gpurun --timeout 1500 -- bash scripts/gpu_configs.sh [tag]
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== configs"; timeout 900 python scripts/bench_configs.py --out $OUT/configs.json 2> $OUT/configs.err | tail -40
tail -5 $OUT/configs.err
for pair in "Tag-v0 B=2^20:tag" "BattleShip 10x10 B=2^18:battleship" "Network-v0:network" "Tiger-v0:tiger" "RockSample(15,15) B=2^22:rock15"; do
  label="${pair%%:*}"; short="${pair##*:}"
  echo "== ncu full $short"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pomdp_ -c 12 -f -o $OUT/$short \
      python scripts/bench_configs.py --steps 20 --only "$label" > $OUT/ncu_$short.log 2>&1
done
ls -la $OUT
