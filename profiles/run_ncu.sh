#!/bin/bash
# Profiling pass of round N (run under gpurun, one GPU): profiles/run_ncu.sh <round-tag>
# 1. plain run of the exact command (must exit 0), 2. launch list, 3. one --set full capture of the
# dominant kernel (the fused rollout step with the dense fp32 encoder), 4. the same for the u8 contract.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-contracts --no-selfplay --no-host-tensor"
$CMD > $OUT/ncu_plain_$TAG.json 2> $OUT/ncu_plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/ncu_plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
# launches: k_reset, 2 x k_rollout_env_multi (100 desync steps), 3 warm-up + 20 timed k_rollout_ws<T>
ncu --set full --clock-control none --import-source on -k regex:k_rollout_ws -s 5 -c 1 -f -o $OUT/prof_${TAG}_rollout_d32 $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full d32 rc=$?"
CMD8="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-contracts --no-selfplay --no-host-tensor --contract d8"
$CMD8 > $OUT/ncu_plain_d8_$TAG.json 2> $OUT/ncu_plain_d8_$TAG.err && \
ncu --set full --clock-control none --import-source on -k regex:k_rollout_ws -s 5 -c 1 -f -o $OUT/prof_${TAG}_rollout_d8 $CMD8 > $OUT/ncu_full_d8_$TAG.log 2>&1
echo "full d8 rc=$?"
CMDE="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-contracts --no-selfplay --no-host-tensor --contract env"
$CMDE > $OUT/ncu_plain_env_$TAG.json 2> $OUT/ncu_plain_env_$TAG.err && \
ncu --set full --clock-control none --import-source on -k regex:k_rollout_env_multi -s 3 -c 1 -f -o $OUT/prof_${TAG}_rollout_env $CMDE > $OUT/ncu_full_env_$TAG.log 2>&1
echo "full env rc=$?"
CMDB="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-contracts --no-selfplay --no-host-tensor --contract bf16"
$CMDB > $OUT/ncu_plain_bf16_$TAG.json 2> $OUT/ncu_plain_bf16_$TAG.err && \
ncu --set full --clock-control none --import-source on -k regex:k_rollout_ws -s 5 -c 1 -f -o $OUT/prof_${TAG}_rollout_bf16 $CMDB > $OUT/ncu_full_bf16_$TAG.log 2>&1
echo "full bf16 rc=$?"
ls -la $OUT
cp open_spiel_coup_b200/libcoup_b200.so $OUT/libcoup_b200_$TAG.so   # for scripts/ncu_hotspots.py (SASS <-> source lines)
