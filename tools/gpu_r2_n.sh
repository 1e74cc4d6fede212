#!/usr/bin/env bash
# (historical record of a round-2 GPU call: the A/B option it toggles was removed after the measurement, see profiles/r02_ab_*.txt)
# round 2: full GPU suite after the clean-up, the tracker convergence test repeated (flaky?), bench
set -u
O=gpurun_out/r2n; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt; tail -n 4 $O/pytest_gpu.txt
for i in 1 2 3 4 5 6; do python -m pytest tests/test_tracking_gpu.py -m gpu -q -k converges 2>&1 | grep -E "assert [0-9]|passed|failed" | head -3; done
python - <<'PY'
import sys, os
sys.path.insert(0, "tests")
import test_tracking_gpu as tt
import __graft_entry__ as ge
for rep in range(4):
    s = tt._setup(); trk = tt._tracker(s)
    got = trk.run(100, alpha_thresh=0.5); more = trk.run(20, alpha_thresh=0.5)
    print("rep", rep, "last5", [round(x, 1) for x in got["loss"][-5:]], "next5", [round(x, 1) for x in more["loss"][:5]])
    trk.close()
PY
python bench.py --steps 30 --warmup 5 --cpu-frames 0 --no-parity --no-extra > $O/bench.json 2> $O/bench.err
python -c "
import json; a=json.load(open('$O/bench.json')); print(a['value'], a['ms_per_step'], 'e2e', a['e2e']['value'], a['stages_ms_per_step'], a['roofline']['issue'])"
