#!/usr/bin/env python
"""Microbenchmark of the pieces of the view-parallel gradient exchange (dp.py), one process per GPU:

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/exchange_bench.py [--P 1000000]

Every piece is timed alone with CUDA events on the launching stream, after a device-side barrier so that
all ranks start together, `--reps` times; printed: median over repetitions of the MAX over ranks.
  nvls b=K      gsr_nvls_allreduce_slice (in-switch multimem.ld_reduce + multimem.st), at most K CTAs
  p2p  b=K      gsr_p2p_allreduce_slice (two-shot over plain P2P loads / stores), at most K CTAs
  sh_ptrs       gsr_sh_grad_from_view_ptrs (SH rebuild reading the peers' masked colour gradients in place)
  sh_local      gsr_sh_grad_from_views on a local gathered buffer (what the NCCL path runs after its all-gather)
  nccl_ar / nccl_ag   NCCL all_reduce(11 P floats) / all_gather(3 P + 4 floats)
  barrier       the symmetric-memory barrier alone
  pair(...)     slice all-reduce on a second stream next to the SH rebuild (what dp.py's exchange runs)
"""
import argparse
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diff-gaussian-rasterization_b200", "full"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem
    C = importlib.import_module("diff_gaussian_rasterization")._C
    P, M = a.P, 16
    head = 3 * P + 4
    numel = 14 * P + 4
    flat = symm_mem.empty(numel, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(flat, dist.group.WORLD)
    mc = int(hdl.multicast_ptr)
    peers = [int(p) for p in hdl.buffer_ptrs]
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    flat.copy_(torch.randn(numel, device=dev, generator=g) * 1e-3)
    means = torch.randn(P, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    gathered = torch.randn(world, head, device=dev) * 1e-3
    plain = torch.randn(11 * P, device=dev)
    head_t = torch.randn(head, device=dev)
    gat_out = torch.empty(world * head, device=dev)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    hdl.barrier(channel=0, timeout_ms=30000)

    def timed(fn, reps=a.reps):
        ts = []
        for i in range(reps + 3):
            hdl.barrier(channel=0, timeout_ms=30000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        t = torch.tensor(ts, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.median())

    res = {"world": world, "P": P}

    def pair(kind, blocks):
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if kind == "nvls":
                C.nvls_allreduce_slice(mc, head, 11 * P, rank, world, blocks)
            else:
                C.p2p_allreduce_slice(peers, head, 11 * P, rank, blocks)
        C.sh_grad_from_view_ptrs(means, peers, [p + 4 * 3 * P for p in peers], 3, M)
        cur.wait_stream(side)

    res["barrier"] = timed(lambda: hdl.barrier(channel=1, timeout_ms=30000))
    for b in (32, 128, 592, 0):
        if mc:
            res["nvls b=%d" % b] = timed(lambda: C.nvls_allreduce_slice(mc, head, 11 * P, rank, world, b))
    for b in (37, 74, 148, 296, 0):
        res["p2p b=%d" % b] = timed(lambda: C.p2p_allreduce_slice(peers, head, 11 * P, rank, b))
    res["sh_ptrs"] = timed(lambda: C.sh_grad_from_view_ptrs(means, peers, [p + 4 * 3 * P for p in peers], 3, M))
    for b in (37, 74, 148, 0):
        res["gather b=%d" % b] = timed(lambda: C.p2p_gather(peers, head, gathered, b))
    res["sh_local"] = timed(lambda: C.sh_grad_from_views(means, gathered, 3, M))
    res["nccl_ar"] = timed(lambda: dist.all_reduce(plain))
    res["nccl_ag"] = timed(lambda: dist.all_gather_into_tensor(gat_out, head_t))
    for kind, b in (("nvls", 32), ("nvls", 128), ("p2p", 37), ("p2p", 74), ("p2p", 148), ("p2p", 0)):
        if kind == "nvls" and not mc:
            continue
        res["pair(%s b=%d)" % (kind, b)] = timed(lambda: pair(kind, b))

    # correctness of the p2p all-reduce against NCCL on the same data (every rank, bit-identical across ranks)
    flat.copy_(torch.randn(numel, device=dev, generator=g))
    want = flat[head:].clone()
    dist.all_reduce(want)
    torch.cuda.synchronize()
    hdl.barrier(channel=0, timeout_ms=30000)
    C.p2p_allreduce_slice(peers, head, 11 * P, rank, 0)
    hdl.barrier(channel=1, timeout_ms=30000)
    torch.cuda.synchronize()
    got = flat[head:]
    err = float((got - want).abs().max() / want.abs().max())
    chk = got.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res["p2p_vs_nccl_max_rel"] = err
    res["p2p_identical_across_ranks"] = bool(float(lo) == float(hi))
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if a.out:
            with open(a.out, "a") as f:
                f.write(line + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
