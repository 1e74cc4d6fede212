#!/usr/bin/env python
"""Multi-GPU check of the gradient exchanges (run under torchrun on >= 2 GPUs):
every rank renders its own view of one scene (fwd+bwd, -full), the scene gradients are exchanged with
each mode of dp.SceneGradReducer, and the reduced gradients of "factorized_sh" and "nvls" are compared
with the plain all-reduce (same sums up to summation order).  Prints one line per mode; exit code 1 on
mismatch.  Usage: python -m torch.distributed.run --nproc-per-node 2 tools/dp_check.py [--config C2]"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist.init_process_group("nccl", device_id=torch.device(dev))
    import bench
    mod = ge.load_variant("full")
    dp = ge.load_dp_module()
    sc, cam, scene, cot = bench.build_inputs(ge, a.config, "full", dev, rank)
    ok = True
    results = {}
    for mode in ("allreduce", "factorized_sh", "nvls"):
        frame = bench.Frame(mod, "full", cam, scene, cot, dev, False, False)
        shapes = {k: tuple(v.shape) for k, v in frame.params.items()}
        red = dp.SceneGradReducer(shapes, dev, mode=mode, means3D=frame.params["means3D"], sh_degree=3)
        red.attach(mod)
        out = None
        for _ in range(a.steps):       # several steps: exercises the double buffering of the nvls arena
            frame.zero_grad()
            frame.step()
            red.reduce_async(frame.grads())
            out = {k: v.clone() for k, v in red.wait().items()}
        torch.cuda.synchronize()
        results[mode] = (red.mode, out, getattr(red, "nvls_note", None))
        red.detach()
        del frame, red
    base = results["allreduce"][1]
    for mode in ("factorized_sh", "nvls"):
        used, out, note = results[mode]
        worst = 0.0
        for k, v in base.items():
            d = (out[k].double() - v.double()).abs().max().item()
            s = v.double().abs().max().item()
            worst = max(worst, d / max(s, 1e-30))
        good = worst < 1e-4
        ok &= good
        if rank == 0:
            print("%s (ran as %s%s): max rel diff vs allreduce %.3e %s" % (
                mode, used, ", " + note if note else "", worst, "ok" if good else "MISMATCH"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
