#!/usr/bin/env python
"""Device-side phase times of bench.py's end-to-end step (CUDA events on the main stream between the phases:
upload wait + input conversion | forward | loss | backward | result packing), to see where the end-to-end
step spends more than the device-resident one.  Usage: python tools/e2e_phases.py [--config C3] [--variant full]"""
import argparse, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3"); ap.add_argument("--variant", default="full")
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--no-upload", action="store_true", help="diagnostic: reuse the first uploaded input set")
    ap.add_argument("--no-convert", action="store_true", help="diagnostic: reuse the first converted depth frame")
    a = ap.parse_args()
    dev = "cuda:0"
    torch.cuda.set_device(0)
    mod = ge.load_variant(a.variant)
    sc, cam, scene, cot = bench.build_inputs(ge, a.config, a.variant, dev, 0)
    f = bench.Frame(mod, a.variant, cam, scene, cot, dev)
    rows = []

    def step(record):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        main_s = torch.cuda.current_stream()
        f.zero_grad()
        ev[0].record()
        if f.pending is None:
            f.pending = f._upload()
        inp, e_up, in_slot = f.pending
        if not a.no_upload:
            f.pending = f._upload()
        main_s.wait_event(e_up)
        d_view, d_proj, d_campos, d_rgb, d_mm = inp
        view = d_view.detach().requires_grad_(True)
        if a.no_convert and hasattr(f, "_gt_cached"):
            gt_d = f._gt_cached
        else:
            gt_d = f._gt_cached = (d_mm * 1e-3).unsqueeze(0)
        rast = f._rasterizer(view.detach(), d_proj, d_campos)
        ev[1].record()
        p = f.params
        res = rast(means3D=p["means3D"], means2D=f.means2D, opacities=p["opacities"], shs=p["shs"],
                   scales=p["scales"], rotations=p["rotations"], viewmatrix=view, gt_depth=gt_d)
        ev[2].record()
        loss, tensors, cots = mod.rgbd_l1_loss(res, d_rgb, d_mm)
        ev[3].record()
        torch.autograd.backward(tensors, cots)
        ev[4].record()
        with torch.no_grad():
            packed = torch.cat([loss.detach().reshape(1), view.grad.reshape(16)])
        slot = f.e2e_steps & 1
        f.e2e_steps += 1
        if f.res_ev[slot] is not None:
            f.res_ev[slot].synchronize()
        f.h_results[slot].copy_(packed, non_blocking=True)
        f.res_ev[slot] = torch.cuda.Event()
        f.res_ev[slot].record(main_s)
        f.slot_free[in_slot] = f.res_ev[slot]
        ev[5].record()
        if record:
            rows.append(ev)

    for _ in range(8):
        step(False)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.iters):
        step(True)
    t1.record()
    torch.cuda.synchronize()
    names = ["inputs", "forward", "loss", "backward", "pack+readback"]
    print("%s %s e2e: %.3f ms/step" % (a.config, a.variant, t0.elapsed_time(t1) / a.iters))
    for i, nme in enumerate(names):
        print("  %-14s median %.3f ms" % (nme, statistics.median(r[i].elapsed_time(r[i + 1]) for r in rows)))
    gaps = [rows[k][5].elapsed_time(rows[k + 1][0]) for k in range(len(rows) - 1)]
    print("  %-14s median %.3f ms" % ("between steps", statistics.median(gaps)))
    # the device-resident step for comparison
    for _ in range(5):
        f.zero_grad(); f.step()
    torch.cuda.synchronize(); t0.record()
    for _ in range(a.iters):
        f.zero_grad(); f.step()
    t1.record(); torch.cuda.synchronize()
    print("  device-resident step: %.3f ms" % (t0.elapsed_time(t1) / a.iters))


if __name__ == "__main__":
    main()
