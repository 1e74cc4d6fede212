#!/usr/bin/env python
"""Interleaved A/B of one library option on small (host-bound, noisy) frames: alternates the option's values
inside ONE process, several rounds each, and prints the per-value median of the per-round device times.
Usage: python tools/ab_small.py --opt spec_render --values 0,1 [--configs C1,C2] [--variant light]"""
import argparse, ctypes, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--opt", required=True); ap.add_argument("--values", default="0,1")
    ap.add_argument("--configs", default="C1,C2"); ap.add_argument("--variant", default="light")
    ap.add_argument("--rounds", type=int, default=6); ap.add_argument("--iters", type=int, default=150)
    a = ap.parse_args()
    vals = [int(v) for v in a.values.split(",")]
    torch.cuda.set_device(0)
    lib = ctypes.CDLL(ge.core_library_path())
    mod = ge.load_variant(a.variant)
    for cfg in a.configs.split(","):
        sc, cam, scene, cot = bench.build_inputs(ge, cfg, a.variant, "cuda:0", 0)
        f = bench.Frame(mod, a.variant, cam, scene, cot, "cuda:0")
        res = {(v, leg): [] for v in vals for leg in ("device", "e2e")}
        for r in range(a.rounds + 1):
            for v in vals:
                assert lib.gsr_set_option(a.opt.encode(), v) >= 0
                for leg, fn in (("device", f.step), ("e2e", lambda: f.step_e2e(True))):
                    for _ in range(10):
                        f.zero_grad(); fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.iters):
                        f.zero_grad(); fn()
                    e1.record(); torch.cuda.synchronize()
                    if r > 0:
                        res[(v, leg)].append(e0.elapsed_time(e1) / a.iters)
        for leg in ("device", "e2e"):
            print("%s %s %-6s " % (cfg, a.variant, leg) + "   ".join(
                "%s=%d: median %.3f ms (min %.3f max %.3f)" % (a.opt, v, statistics.median(res[(v, leg)]), min(res[(v, leg)]), max(res[(v, leg)]))
                for v in vals), flush=True)


if __name__ == "__main__":
    main()
