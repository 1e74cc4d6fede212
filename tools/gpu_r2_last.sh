#!/usr/bin/env bash
# round 2: last sanity run of the final binaries: smoke, GPU suite, bench line (2-GPU lines when the box has two GPUs)
set -u
O=gpurun_out/r2z; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --cpu-frames 0 --no-extra > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
FILES="bench_n1"
if [ $(nvidia-smi -L | wc -l) -ge 2 ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
   bench.py --gpus 2 --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 \
   bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > $O/bench_n2_ref.json 2> $O/bench_n2_ref.err; echo "ref n2 rc=$?"
FILES="bench_n1 bench_n2 bench_n2_ref"
fi
python - <<PY
import json
for f in "$FILES".split():
    a=json.load(open("$O/%s.json"%f))
    print(f, round(a["value"],1), round(a["ms_per_step"],3), "e2e", round(a["e2e"]["value"],1), (a.get("parity") or {}).get("vs_reference",{}).get("ok"), (a.get("dp_check") or {}).get("max_rel"), a.get("roofline",{}).get("issue",{}).get("frac"))
PY
