#!/bin/bash
TAG=${1:-r03i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest battleship"; timeout 900 python -m pytest tests -m gpu -q -x -k "battleship or ship or odd_boards or fullsize" 2>&1 | tail -3
ONLY="BattleShip" bash scripts/exp_network_variants.sh $TAG "" "-DPOMDP_SHIP_RESET_THREADS=256" "-DPOMDP_SHIP_RESET_THREADS=512" "-DPOMDP_SHIP_RESET_THREADS=64"
