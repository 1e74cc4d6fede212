#!/usr/bin/env python
"""Turn the raw outputs of tools/gpu_r2_final.sh (gpurun_out/final2/) into the committed evidence under
profiles/: bench lines, summarised ncu launch lists (per kernel: launches, mean time, share of the
frame), ncu --set full summaries (tools/ncu_summary.py) and the DRAM-traffic table bench.py reads.
Usage: python tools/collect_profiles.py <tag>      e.g. v2   (after tools/gpu_r2_final.sh)"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "final2")
DST = os.path.join(ROOT, "profiles")


def launch_summary(path, out, title):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[ik], float(r[iv].replace(",", ""))) for r in rows[1:]]
    agg = collections.OrderedDict()
    for n, v in seq:
        n = n.split("(")[0].replace("void ", "").replace("gsr::<unnamed>::", "").replace("unnamed>::", "")
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# %s\n# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold caches:\n"
                "# the SHARE of the step is what must agree with bench.py's stage timers, not the absolute)\n" % title)
        f.write("%-64s %8s %12s %8s\n" % ("kernel", "launches", "mean us", "share"))
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-64s %8d %12.2f %7.1f%%\n" % (n[:64], c, t / c / 1000.0, 100.0 * t / total))
    return agg


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "vX"
    cp = lambda a, b: shutil.copy(os.path.join(SRC, a), os.path.join(DST, b))
    cp("bench.json", "r02_bench_b200_C3_full_%s.json" % tag)
    cp("bench_light.json", "r02_bench_b200_C3_light_%s.json" % tag)
    cp("bench_ref.json", "r02_bench_reference_C3_full_%s.json" % tag)
    cp("tracking_C2.jsonl", "r02_tracking_C2_%s.jsonl" % tag)
    cp("configs.txt", "r02_configs_%s.txt" % tag)
    launch_summary(os.path.join(SRC, "launches_C3.csv"), os.path.join(DST, "r02_ncu_launches_C3_full_%s.txt" % tag),
                   "python bench.py --steps 2 --warmup 3 --cpu-frames 0   (C3 full; includes torch's own kernels)")
    launch_summary(os.path.join(SRC, "launches_C4.csv"), os.path.join(DST, "r02_ncu_launches_C4_full_%s.txt" % tag),
                   "python bench.py --config C4 --steps 2 --warmup 3 --cpu-frames 0   (C4 full)")
    launch_summary(os.path.join(SRC, "tracking_launches.csv"), os.path.join(DST, "r02_ncu_launches_tracking_C2_%s.txt" % tag),
                   "python tools/bench_tracking.py --arms tracker --iters 6 --reps 1 under ncu --graph-profiling node")
    for cfg in ("C3", "C4"):
        raw = os.path.join(SRC, "prof_%s_raw.csv" % cfg)
        src = os.path.join(SRC, "prof_%s_src.csv" % cfg)
        args = [sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw] + ([src] if os.path.exists(src) else [])
        txt = subprocess.run(args, capture_output=True, text=True).stdout
        with open(os.path.join(DST, "r02_ncu_full_%s_%s.txt" % (cfg, tag)), "w") as f:
            f.write("# ncu --set full --clock-control none, one frame of bench.py --config %s (full variant)\n" % cfg)
            f.write(txt)
    # DRAM traffic per launch of the blend kernels (bench.py: roofline.traffic)
    traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full "
               "--clock-control none` captures under gpurun (bytes); read by bench.py for roofline.traffic. "
               "Source: profiles/r02_ncu_full_C3_%s.txt / r02_ncu_full_C4_%s.txt" % (tag, tag)}
    for cfg in ("C3", "C4"):
        rows = list(csv.reader(open(os.path.join(SRC, "prof_%s_raw.csv" % cfg))))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = d["Kernel Name"]
            key = ("render_bwd" if "render_bwd" in name else "render_fwd" if "render_fwd" in name else
                   "preprocess_bwd" if "preprocess_bwd" in name else "preprocess_fwd" if "preprocess_fwd" in name else None)
            if key is None:
                continue
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            tot = sum(float(d[m]) * scale.get(u[m], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            traffic.setdefault(key, {})[cfg] = int(tot)
            traffic.setdefault("inst_executed", {}).setdefault(key, {})[cfg] = int(float(d["smsp__inst_executed.sum"]))
    json.dump(traffic, open(os.path.join(DST, "ncu_traffic.json"), "w"), indent=1)
    # GPU test summary and sanitizer summaries
    with open(os.path.join(DST, "r02_gputest_%s.txt" % tag), "w") as f:
        f.write("# python -m pytest tests -m gpu -q   (B200, tools/gpu_r2_final.sh)\n")
        f.write("".join(open(os.path.join(SRC, "pytest_gpu.log"), errors="replace").readlines()[-6:]))
        f.write("\n# python -c 'import __graft_entry__ as g; g.smoke()'\n")
        f.write("".join(open(os.path.join(SRC, "smoke.log"), errors="replace").readlines()[-6:]))
    for seed in (2, 3):
        fn = os.path.join(SRC, "parity_fuzz_1500_seed%d.txt" % seed)
        if os.path.exists(fn):
            shutil.copy(fn, os.path.join(DST, "r02_parity_fuzz_1500_seed%d_%s.txt" % (seed, tag)))
    with open(os.path.join(DST, "r02_sanitizers_%s.txt" % tag), "w") as f:
        f.write("# compute-sanitizer --tool <tool> python tools/sanitize_small.py  (both variants, every blend-kernel variant,\n"
                "# radix + tile-local binning, a speculative frame that overflows and is redone, the exchange helpers, the loss\n"
                "# helper, the tracker's CUDA graph); initcheck with bulk_sh=0 (it does not track cp.async.bulk global writes)\n")
        for tool in ("memcheck", "racecheck", "initcheck"):
            fn = os.path.join(SRC, "sanitize_%s.txt" % tool)
            if os.path.exists(fn):
                lines = open(fn, errors="replace").readlines()
                keep = [l for l in lines if "ERROR SUMMARY" in l or "RACECHECK SUMMARY" in l or l.startswith("exit ") or "done" in l]
                f.write("== %s ==\n%s" % (tool, "".join(keep[-8:])))
    print(open(os.path.join(DST, "r02_ncu_launches_C3_full_%s.txt" % tag)).read())
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
