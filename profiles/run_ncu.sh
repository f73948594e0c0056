#!/bin/bash
# Profiling pass (run under gpurun, one GPU): profiles/run_ncu.sh <tag>
# 1. plain run of the exact command (must exit 0), 2. launch list of the same command, 3. one --set full capture of each
# kernel the bench line talks about, exported on the box to the raw / details CSVs that are committed under profiles/
# (gpurun brings back at most 64 MiB, so only the reports listed in KEEP_REPS travel as .ncu-rep for source-level analysis).
set -u
TAG=${1:-r10}
OUT=gpurun_out
mkdir -p $OUT
BASE="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-contracts --no-selfplay --no-host-tensor"
$BASE > $OUT/ncu_plain_$TAG.json 2> $OUT/ncu_plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/ncu_plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $BASE > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
full() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > $OUT/ncu_plain_${name}_$TAG.log 2>&1 || { echo "plain $name failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o $OUT/prof_${TAG}_$name "$@" > $OUT/ncu_full_${name}_$TAG.log 2>&1
  echo "full $name rc=$?"
  ncu -i $OUT/prof_${TAG}_$name.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_${TAG}_$name.ncu-rep --page details --csv > $OUT/${TAG}_${name}_details.csv 2>/dev/null
  case " $KEEP_REPS " in *" $name "*) ;; *) rm -f $OUT/prof_${TAG}_$name.ncu-rep ;; esac
}
KEEP_REPS=${KEEP_REPS:-"rollout_env kernels_record incremental"}
full rollout_d32 k_rollout_ws 5 $BASE
full rollout_d8 k_rollout_ws 5 $BASE --contract d8
full rollout_bf16 k_rollout_ws 5 $BASE --contract bf16
full rollout_env k_rollout_env_multi 2 python scripts/env_only_probe.py --steps 128
full incremental k_rollout_incremental 3 python scripts/incremental_probe.py 20
full kernels_obs k_encode_obs 2 python scripts/kernel_probe.py obs
full kernels_policy k_sample_policy 2 python scripts/kernel_probe.py policy
full kernels_mask k_legal_actions_mask 2 python scripts/kernel_probe.py mask
full kernels_step "^k_step$" 2 python scripts/kernel_probe.py step
full kernels_record k_step_record 2 python scripts/kernel_probe.py record
full kernels_fork k_fork 2 python scripts/kernel_probe.py fork
cp open_spiel_coup_b200/libcoup_b200.so $OUT/libcoup_b200_$TAG.so   # for scripts/ncu_hotspots.py / ncu_functions.py (SASS <-> source lines)
du -sh $OUT; ls -la $OUT | tail -40
