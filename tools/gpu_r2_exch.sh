#!/usr/bin/env bash
# round 2: microbenchmark of the exchange pieces (tools/exchange_bench.py) at every rank count the box offers
set -u
O=gpurun_out/r2x; mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
     tools/exchange_bench.py --out $O/exchange_bench.jsonl > $O/xb_n$n.log 2>&1; echo "n=$n exit $?"
  tail -n 3 $O/xb_n$n.log | cut -c1-1500
done
