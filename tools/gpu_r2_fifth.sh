#!/usr/bin/env bash
# round 2, call 5 (2 GPUs): data-parallel exchange modes + dp_check + e2e scaling; 1-GPU checks of the new scan kernel / loss helper
set -u
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "rgbd or speculative or interleaved or tracking or golden or arena or exchange" > $O/pytest_subset.txt 2>&1; echo "pytest exit $?" >> $O/pytest_subset.txt
tail -n 4 $O/pytest_subset.txt
python bench.py --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity > $O/bench_n1.json 2> $O/bench_n1.err
for mode in auto factorized_sh allreduce; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --dp-mode $mode > $O/bench_n2_$mode.json 2> $O/bench_n2_$mode.err
  echo "n2 $mode exit $?"
done
GSR_DP_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-dp-check > $O/bench_n2_timing.json 2> $O/bench_n2_timing.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/bench_n2_ref.json 2> $O/bench_n2_ref.err
python - <<PY
import json
for f in ("bench_n1","bench_n2_auto","bench_n2_factorized_sh","bench_n2_allreduce","bench_n2_timing","bench_n2_ref"):
    try:
        a=json.load(open("$O/%s.json"%f))
        print(f, "%.1f fps %.3f ms  e2e %.1f  torch-loss e2e %s  exch %s  dp_check %s stages %s" % (a["value"], a["ms_per_step"], a["e2e"]["value"], (a.get("e2e_torch_loss") or {}).get("value"), a["stats"].get("exchange"), a.get("dp_check"), {k: round(v,3) for k,v in a.get("stages_ms_per_step",{}).items()}))
    except Exception as e:
        print(f, "ERR", e)
PY
grep -h "nvls\|exchange mode" $O/*.err | head -8
