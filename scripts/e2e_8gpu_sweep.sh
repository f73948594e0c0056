#!/bin/bash
# End-to-end leg of bench.py on 8 GPUs under different host-side settings (run under gpurun --gpus 8).
run() {  # label port extra-bench-args...
  local label=$1 port=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 \
      --steps 200 --warmup 5 --no-cpu-baseline --no-extra-contracts --no-selfplay --no-host-tensor "$@" 2> gpurun_out/e2e8_$label.err \
    | python -c 'import sys,json
lines=[l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")]
if not lines: print(sys.argv[1], "no JSON line (see gpurun_out/e2e8_%s.err)" % sys.argv[1]); sys.exit(0)
d=json.loads(lines[-1]); print(sys.argv[1], "value %.3e e2e %.3e threads %d" % (d["value"], d["e2e"]["value"], d["e2e"]["host_policy_threads"]))' $label
}
mkdir -p gpurun_out
nproc
run spin-4 29531
run block-4 29532 --blocking-sync
run spin-3 29533 --host-threads 3
run block-6 29534 --host-threads 6 --blocking-sync
run spin-4-8slabs 29535 --e2e-slabs 8
