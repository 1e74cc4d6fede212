#!/usr/bin/env python
"""Tracking-loop benchmark: iterations per second of K pose-tracking
iterations at config C2 (100 k Gaussians, 640x480, -light, map_off) for
  tracker        : the device-side tracker (CUDA graph, fused loss, on-chip pose update)
  torch_loop     : the same loop through our -light package + torch loss / autograd / Adam
  reference_loop : the same loop through the reference's own CUDA build (baseline/_ref), if present
Prints one JSON line per arm.  Usage: python tools/bench_tracking.py [--config C2] [--iters 50]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--arms", default="tracker,torch_loop,reference_loop")
    a = ap.parse_args()
    import test_tracking_gpu as tt
    sc = ge.load_scene_module()
    P, W, H, sig = sc.CONFIGS[a.config]
    s = tt._setup(P=P, W=W, H=H, sig=sig, seed=0)
    dev = torch.device("cuda:0")

    def timed(fn):
        fn()  # warm-up (graph capture, allocator, autotune)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(a.reps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    base = dict(metric="pose-tracking iterations/sec", config=a.config, gaussians=P, width=W, height=H,
                iterations=a.iters, unit="iterations/s", timing="host wall clock around the whole "
                "K-iteration call incl. final pose read-back, best of %d" % a.reps)
    trk = tt._tracker(s, max_iterations=max(a.iters, 4))
    res = {}

    def run_tracker():
        trk.set_pose(s["q0"], s["t0"])
        res["t"] = trk.run(a.iters)
    dt = timed(run_tracker)
    print(json.dumps(dict(base, arm="tracker", value=a.iters / dt, ms_per_iteration=1e3 * dt / a.iters,
                          num_rendered=res["t"]["num_rendered"], loss_first=res["t"]["loss"][0],
                          loss_last=res["t"]["loss"][-1],
                          graph_nodes_per_iteration=res["t"]["kernels_per_iteration"])))
    for arm, mod in (("torch_loop", s["mod"]), ("reference_loop", ge.load_reference("light"))):
        if arm not in a.arms.split(","):
            continue
        if mod is None:
            print(json.dumps(dict(base, arm=arm, unavailable="baseline/_ref not on this box")))
            continue
        def run_loop():
            res[arm] = tt._loop(s, mod, a.iters)
        dt = timed(run_loop)
        print(json.dumps(dict(base, arm=arm, value=a.iters / dt, ms_per_iteration=1e3 * dt / a.iters,
                              loss_first=res[arm]["loss"][0], loss_last=res[arm]["loss"][-1])))


if __name__ == "__main__":
    main()
