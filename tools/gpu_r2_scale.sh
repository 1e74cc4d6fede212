#!/usr/bin/env bash
# round 2 (8 GPUs): data-parallel bench at 4 and 8 ranks, early gather on / off
set -u
O=gpurun_out/r2s; mkdir -p $O
for n in 8 4; do
  for early in 1 0; do
    if [ $early = 0 ]; then export GSR_DP_NO_EARLY=1; extra="--no-dp-check"; else unset GSR_DP_NO_EARLY; extra=""; fi
    GSR_DP_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n \
       bench.py --gpus $n --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity $extra > $O/bench_n${n}_early$early.json 2> $O/bench_n${n}_early$early.err; echo "bench n=$n early=$early exit $?"
    python - <<PY
import json
a=json.load(open("$O/bench_n${n}_early$early.json"))
print("n=$n early=$early: %.1f fps %.3f ms  e2e %.1f (%.3f ms)  exch %s  dp_check %s" % (a["value"], a["ms_per_step"], a["e2e"]["value"], a["e2e"]["ms_per_step"], a["stats"].get("exchange"), (a.get("dp_check") or {}).get("max_rel")))
PY
    grep -h "phases" $O/bench_n${n}_early$early.err
  done
done
