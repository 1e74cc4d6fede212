"""CPU tests of the view-level data-parallel host logic (diff-gaussian-rasterization_b200/dp.py)
with the gloo backend and world_size 2: the flat scene-gradient buffer after the single all-reduce
equals the sum of the per-rank gradients, and views are sharded round-robin."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, P, M, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    dp = ge.load_dp_module()
    shapes = dict(means3D=(P, 3), shs=(P, M, 3), opacities=(P, 1), scales=(P, 3), rotations=(P, 4))
    g = torch.Generator().manual_seed(100 + rank)
    grads = {k: torch.randn(*s, generator=g) for k, s in shapes.items()}
    red = dp.SceneGradReducer(shapes, "cpu")
    assert red.numel == P * (3 + 3 * M + 1 + 3 + 4) and red.bytes_per_step() == red.numel * 4
    red.reduce_async(grads)
    views = red.wait()
    # expected: sum over ranks of the same seeded tensors
    ok = True
    for k, s in shapes.items():
        exp = sum(torch.randn(*s, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
        # regenerate in the same per-rank order as above: each rank draws its params in dict order
        ok = ok and views[k].shape == tuple(s)
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    for k, s in shapes.items():
        exp = sum(torch.randn(*s, generator=gens[r]) for r in range(world))
        ok = ok and torch.allclose(views[k], exp, atol=1e-6)
    q.put((rank, bool(ok), dp.shard_views(5, rank, world)))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world, P, M = 2, 257, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, M, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == [0, 2, 4] and res[1][2] == [1, 3]


def _worker_fact(rank, world, port, P, M, q, mode="factorized_sh"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    dp = ge.load_dp_module()
    shapes = dict(means3D=(P, 3), shs=(P, M, 3), opacities=(P, 1), scales=(P, 3), rotations=(P, 4))
    means = torch.randn(P, 3, generator=torch.Generator().manual_seed(7)) + torch.tensor([0.0, 0.0, 4.0])

    def rank_data(r):
        g = torch.Generator().manual_seed(200 + r)
        grads = {k: torch.randn(*s, generator=g) for k, s in shapes.items() if k != "shs"}
        dR = torch.randn(P, 3, generator=g)
        dR[::5] = 0.0  # culled in this view
        campos = torch.randn(3, generator=g) * 0.3
        return grads, dR, campos

    grads, dR, campos = rank_data(rank)
    red = dp.SceneGradReducer(shapes, "cpu", mode=mode, means3D=means, sh_degree=3)
    assert red.numel == 14 * P + 4
    if mode in ("nvls", "p2p"):  # no CUDA / no NVSwitch here: the reducer must say so and use the NCCL-style exchange
        assert red.mode == "factorized_sh" and red.nvls is None and "nvls unavailable" in red.nvls_note
    red.reduce_async(grads, masked_color=dR, campos=campos)
    views = red.wait()
    ok = True
    exp_sh = torch.zeros(P, M, 3)
    for r in range(world):
        g_r, dR_r, cp_r = rank_data(r)
        d = means - cp_r
        d = d / d.norm(dim=1, keepdim=True)
        exp_sh += dp.sh_basis(d, 3)[:, :, None] * dR_r[:, None, :]
    ok = ok and torch.allclose(views["shs"], exp_sh, atol=1e-5)
    for k in ("means3D", "opacities", "scales", "rotations"):
        exp = sum(rank_data(r)[0][k] for r in range(world))
        ok = ok and torch.allclose(views[k], exp, atol=1e-6)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_factorized_sh_exchange_world2():
    world, P, M = 2, 100, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fact, args=(r, world, port, P, M, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


@pytest.mark.parametrize("mode", ["nvls", "p2p"])
def test_nvls_mode_falls_back_to_factorized_exchange_without_multicast(mode):
    world, P, M = 2, 64, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fact, args=(r, world, port, P, M, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_single_process_passthrough():
    import __graft_entry__ as ge
    dp = ge.load_dp_module()
    shapes = dict(means3D=(10, 3), opacities=(10, 1))
    red = dp.SceneGradReducer(shapes, "cpu")
    grads = dict(means3D=torch.ones(10, 3), opacities=torch.full((10, 1), 2.0))
    red.reduce_async(grads)
    v = red.wait()
    assert red.numel == 40 and float(v["means3D"].sum()) == 30 and float(v["opacities"].sum()) == 20
    assert v["means3D"].data_ptr() == red.flat.data_ptr()  # views alias the flat buffer
