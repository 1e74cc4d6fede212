#!/usr/bin/env bash
# round 2 (8 GPUs): exchange microbenchmark at 4 and 8 ranks, then one bench run at 8 ranks with phase timing
set -u
O=gpurun_out/r2x8; mkdir -p $O
for n in 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
     tools/exchange_bench.py --reps 12 --out $O/exchange_bench.jsonl > $O/xb_n$n.log 2>&1; echo "n=$n exit $?"
  tail -n 1 $O/xb_n$n.log | cut -c1-1600
done
for slice in nvls p2p; do
GSR_DP_SLICE=$slice GSR_DP_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
   bench.py --gpus 8 --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity > $O/bench_n8_$slice.json 2> $O/bench_n8_$slice.err; echo "bench n8 $slice exit $?"
python - <<PY
import json
a=json.load(open("$O/bench_n8_$slice.json"))
print("%.1f fps %.3f ms  e2e %.1f  exch %s  dp_check %s" % (a["value"], a["ms_per_step"], a["e2e"]["value"], a["stats"].get("exchange"), (a.get("dp_check") or {}).get("max_rel")))
PY
grep -h "phases" $O/bench_n8_$slice.err
done
