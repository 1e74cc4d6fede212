#!/usr/bin/env python
"""Where does the HOST time of a small frame go?  Runs bench.py's Frame.step / step_e2e at a small config and
prints (a) wall time per step against the device time per step, (b) host time of the forward call, the backward
call and the rest, measured with perf_counter around them (no extra synchronisation: what the host spends
enqueueing), (c) the top of a cProfile of the same loop.
Usage: python tools/host_profile.py [--config C2] [--variant light] [--iters 300] [--impl b200|reference]"""
import argparse, cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2"); ap.add_argument("--variant", default="light")
    ap.add_argument("--iters", type=int, default=300); ap.add_argument("--impl", default="b200")
    ap.add_argument("--e2e", action="store_true")
    a = ap.parse_args()
    dev = "cuda:0"
    torch.cuda.set_device(0)
    mod = ge.load_reference(a.variant) if a.impl == "reference" else ge.load_variant(a.variant)
    sc, cam, scene, cot = bench.build_inputs(ge, a.config, a.variant, dev, 0)
    f = bench.Frame(mod, a.variant, cam, scene, cot, dev)
    fn = (lambda: f.step_e2e(True)) if a.e2e else f.step

    def run(n):
        for _ in range(n):
            f.zero_grad()
            fn()
    run(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); run(a.iters); e1.record(); t_enq = time.perf_counter() - t0
    torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    print("%s %s %s%s: device %.3f ms/step, host enqueue %.3f ms/step, wall %.3f ms/step" % (
        a.impl, a.config, a.variant, " e2e" if a.e2e else "", e0.elapsed_time(e1) / a.iters, 1e3 * t_enq / a.iters, 1e3 * t_all / a.iters))
    if not a.e2e:
        # split: forward call / backward call (host side only)
        p = f.params
        tf = tb = 0.0
        for _ in range(a.iters):
            f.zero_grad()
            t0 = time.perf_counter()
            res = f.rast(means3D=p["means3D"], means2D=f.means2D, opacities=p["opacities"], shs=p["shs"],
                         scales=p["scales"], rotations=p["rotations"], viewmatrix=f.view, gt_depth=f.gt)
            t1 = time.perf_counter()
            torch.autograd.backward(f._outs(res), f.cots)
            t2 = time.perf_counter()
            tf += t1 - t0; tb += t2 - t1
        torch.cuda.synchronize()
        print("  host: forward call %.3f ms, backward call %.3f ms per step" % (1e3 * tf / a.iters, 1e3 * tb / a.iters))
    pr = cProfile.Profile(); pr.enable(); run(a.iters); torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print("\n".join(l[:170] for l in s.getvalue().splitlines()[:48]))


if __name__ == "__main__":
    main()
