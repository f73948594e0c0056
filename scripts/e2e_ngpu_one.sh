#!/bin/bash
# bash scripts/e2e_ngpu_one.sh <ngpus> <label> <port> [bench args]: end-to-end leg only, one configuration
n=$1; label=$2; port=$3; shift 3
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n \
    --steps 200 --warmup 5 --no-cpu-baseline --no-extra-contracts --no-selfplay --no-host-tensor "$@" 2> gpurun_out/e2e_$label.err \
  | python -c 'import sys,json
lines=[l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")]
if not lines: print(sys.argv[1], "no JSON line"); sys.exit(0)
d=json.loads(lines[-1]); print(sys.argv[1], "value %.3e e2e %.3e threads %d slabs %d" % (d["value"], d["e2e"]["value"], d["e2e"]["host_policy_threads"], d["e2e"]["sub_slabs"]))' $label
