#!/bin/bash
# r04i: batches that share the agent's cell (the agent's position is observed in RockSample and Tag): the Tag pair table at
# a pitch of 33 vs the conflicting pitch of 32, plus GPU parity of the Tag paths
OUT=gpurun_out/r04i; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "tag or Tag or rollout or heuristic or fullsize" 2>&1 | tail -3 | tee $OUT/pytest.log
python scripts/bench_action_classes.py --only Tag --out $OUT/action_classes_tag.json 2>&1 | cut -c1-190 | tee $OUT/action_classes_tag.log
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DPOMDP_TAG_PAIR_PITCH=32 -o /tmp/lib_pitch32.so gym_pomdp_b200/csrc/pomdp_kernels.cu 2>&1 | grep -E "error"
echo "== pitch 32 (the old layout)" | tee -a $OUT/action_classes_tag.log
POMDP_B200_LIB=/tmp/lib_pitch32.so python scripts/bench_action_classes.py --only Tag 2>&1 | cut -c1-190 | tee -a $OUT/action_classes_tag.log
python scripts/bench_action_classes.py --only "RockSample(11,11)" 2>&1 | cut -c1-190 | tee $OUT/action_classes_rock.log
