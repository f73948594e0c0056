#!/bin/bash
# one configuration of the sweep in e2e_8gpu_sweep.sh: bash scripts/e2e_8gpu_one.sh <label> <port> [bench args]
label=$1; port=$2; shift 2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 \
    --steps 200 --warmup 5 --no-cpu-baseline --no-extra-contracts --no-selfplay --no-host-tensor "$@" 2> gpurun_out/e2e8_$label.err \
  | python -c 'import sys,json
lines=[l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")]
if not lines: print(sys.argv[1], "no JSON line"); sys.exit(0)
d=json.loads(lines[-1]); print(sys.argv[1], "value %.3e e2e %.3e threads %d" % (d["value"], d["e2e"]["value"], d["e2e"]["host_policy_threads"]))' $label
