#!/usr/bin/env python
"""bench.py — fwd+bwd frames/s of the differentiable Gaussian-splat rasterizer hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json metric "fwd+bwd frames/sec @1080p, 1M Gaussians", configs[2] = C3):
1 000 000 synthetic Gaussians, 1920x1080, -full variant, SH degree 3, forward + backward with
cotangents on every differentiable output (colour, depth, silhouette) incl. dL/dviewmatrix.
A "step" is one frame = GaussianRasterizer.forward + autograd backward of one camera view per GPU;
at N > 1 every rank renders its own view of the same replicated scene and the step ends with ONE
all-reduce of the flat scene-parameter gradient buffer (59 floats per Gaussian).

  value      frames/s with every input resident in HBM (device-timed with CUDA events, max over ranks;
             seeded N(0,1) cotangents on every differentiable output)
  e2e        one training step through the same public API the way a CG-SLAM-style caller runs it:
             that step's inputs — the camera (viewmatrix, projmatrix, campos) and the ground-truth
             RGB-D frame (colour uint8 [3,H,W], depth int16 millimetres [H,W], the formats RGB-D
             datasets ship) — are copied from PINNED HOST memory inside the timed region, the L1
             colour + depth (+ silhouette) loss is evaluated with torch on the device, autograd
             produces the per-pixel cotangents, and the step's result (loss + dL/dviewmatrix, 17
             floats) is read back to the host.  The Gaussians are the caller's model state and stay
             resident, as in the reference's API contract (all tensor arguments are CUDA tensors).
             (Round 1 uploaded 41.5 MB of synthetic fp32 cotangents per step instead, which no caller
             does and which made N ranks on one host memory controller PCIe-bound.)
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event time, measured
             live through the library's stage timer on the launching stream, against
             MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the CPU oracle (a port: the reference has NO CPU path) on the host cores, bounded sample

--impl reference times the UNMODIFIED reference CUDA build (baseline/_ref, compiled for sm_100 by
baseline/build_ref.sh) through its own public API on the same tensors; if that build is not on the
box it falls back to the CPU oracle port on a bounded sample (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fwd+bwd frames/sec @1080p, 1M Gaussians"
UNIT = "frames/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---- clocks ---------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- algorithmic bytes per stage ----------------------------------------------------------------
# Per-Gaussian and blend stages: SURVEY.md 8d (reference state set, M SH coefficients).
# Binning: what the stage that is actually timed moves.  With tile-local binning (default) that is
#   tile_scan     16 B / tile        (counter in, range + scatter cursor out)
#   tile_scatter  28 B / Gaussian in (tiles_touched, rect, depth) + 12 B / duplicate (cursor atomic, 8-byte entry out)
#   tile_sort     12 B / duplicate   (8-byte entry in, 4-byte index out; the sort itself is on chip)
# against the reference model's scan 8 P + emit 20 P + 12 N + radix sort (24 b + 8) N + ranges 8 N + 8 tiles,
# which is quoted separately ("reference_model_bytes") and charged only when the radix path runs.

STAGE_LABEL_TILE = {"scan": "tile_scan", "emit_keys": "tile_scatter", "radix_sort": "tile_sort"}


def stage_bytes(stage, P, N, HW, tiles, M, variant, sort_bits, tile_local=True):
    b = (sort_bits + 7) // 8
    c_out = 7 if variant == "light" else 5
    c_grad = 6 if variant == "light" else 5
    table = {
        "preprocess_fwd": (44 + 12 * M) * P + 75 * P,
        "render_fwd": 44 * N + HW * (4 + 4 * c_out + 8),
        "render_bwd": 44 * N + HW * (4 * c_grad + 12) + 48 * P,
        # SURVEY 8d's 359 + 44 + 12 M bytes minus the 56 bytes per Gaussian of gradients nobody reads, which the
        # kernel no longer writes (dL/dconic 16, dL/ddepth 4, dL/dcolors_precomp 12, dL/dcov3D_precomp 24)
        "preprocess_bwd": (303 + 44 + 12 * M) * P,
    }
    if tile_local:
        table.update({"scan": 16 * tiles, "emit_keys": 28 * P + 12 * N, "radix_sort": 12 * N})
    else:
        table.update({"scan": 8 * P, "emit_keys": 20 * P + 12 * N, "radix_sort": (24 * b + 8) * N,
                      "tile_ranges": 8 * N + 8 * tiles})
    return table.get(stage)


def reference_binning_bytes(P, N, tiles, sort_bits):
    b = (sort_bits + 7) // 8
    return 8 * P + 20 * P + 12 * N + (24 * b + 8) * N + 8 * N + 8 * tiles


# ---- the workload ---------------------------------------------------------------------------

def build_inputs(ge, cfg_name, variant, device, view_seed, pin=False):
    import torch
    sc = ge.load_scene_module()
    P, W, H, sig = sc.CONFIGS[cfg_name]
    cam0 = sc.make_camera(W, H)
    scene = sc.make_scene(P, cam0, sig, seed=0)          # same replicated scene on every rank
    cam = sc.make_camera(W, H, seed=view_seed)           # view k of the batch
    n_aux = 3 if variant == "light" else 2
    cot = sc.make_cotangents(cam, n_aux, seed=1 + view_seed)
    return sc, cam, scene, cot


class Frame:
    """One camera view's fwd+bwd through the public Python API of `mod`."""

    def __init__(self, mod, variant, cam, scene, cot, device, track_off=False, map_off=False, view_seed=0,
                 host_inputs=True):
        import torch
        self.track_off, self.map_off = track_off, map_off
        self.torch = torch
        self.mod, self.variant, self.device = mod, variant, device
        d = lambda t, rg=False: t.to(device).clone().requires_grad_(rg)
        self.params = dict(means3D=d(scene.means3D, True), shs=d(scene.shs, True),
                           opacities=d(scene.opacities, True), scales=d(scene.scales, True),
                           rotations=d(scene.rotations, True))
        self.means2D = torch.zeros_like(self.params["means3D"], requires_grad=True)
        self.view = d(cam.viewmatrix, True)
        self.gt = scene.gt_depth.to(device)
        self.cots = [cot[0].to(device)] + [c.to(device) for c in cot[1]]
        self.cam, self.scene = cam, scene
        # constants of every frame live on the device once (a `.to(device)` of a pageable CPU tensor is a
        # blocking copy: inside the step it would make the host wait for the whole previous step)
        self.bg_dev, self.perspec_dev = scene.bg.to(device), cam.perspec_matrix.to(device)
        self.rast = self._rasterizer(cam.viewmatrix.to(device), cam.projmatrix.to(device), cam.campos.to(device))
        self.last = None
        self.copy_stream = None
        self.h2d_bytes = self.d2h_bytes = 0
        if host_inputs:
            # pinned host copies of ONE step's inputs (e2e leg): the camera and the ground-truth RGB-D
            # frame in the formats RGB-D datasets ship (8-bit colour, 16-bit depth in millimetres)
            pin = lambda t: t.clone().pin_memory()
            g = torch.Generator(device="cpu").manual_seed(77 + view_seed)
            self.h_view, self.h_proj, self.h_campos = pin(cam.viewmatrix), pin(cam.projmatrix), pin(cam.campos)
            self.h_rgb = pin(torch.randint(0, 256, (3, cam.H, cam.W), generator=g, dtype=torch.uint8))
            self.h_depth_mm = pin((scene.gt_depth[0] * 1000.0).round().to(torch.int16))  # 500 .. 10 000 mm
            self.h_inputs = [self.h_view, self.h_proj, self.h_campos, self.h_rgb, self.h_depth_mm]
            self.h_results = [torch.empty(17, dtype=torch.float32).pin_memory() for _ in range(2)]
            self.res_ev = [None, None]
            self.e2e_steps = 0
            self.pending = None
            self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.h_inputs)
            self.d2h_bytes = 17 * 4

    def _rasterizer(self, view, proj, campos):
        torch, cam, scene, dev = self.torch, self.cam, self.scene, self.device
        kw = dict(image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                  bg=self.bg_dev, scale_modifier=1.0, viewmatrix=view, projmatrix=proj, sh_degree=3,
                  campos=campos, prefiltered=False, perspec_matrix=self.perspec_dev)
        if self.variant == "light":
            kw.update(debug=False, track_off=self.track_off, map_off=self.map_off)
        return self.mod.GaussianRasterizer(self.mod.GaussianRasterizationSettings(**kw))

    def _outs(self, res):
        if self.variant == "light":
            return [res[0], res[2], res[3], res[4]]
        return [res[0], res[2], res[3]]

    def zero_grad(self):
        for t in list(self.params.values()) + [self.means2D, self.view]:
            t.grad = None

    def step(self):
        """Device-resident frame: forward + backward."""
        p = self.params
        res = self.rast(means3D=p["means3D"], means2D=self.means2D, opacities=p["opacities"], shs=p["shs"],
                        scales=p["scales"], rotations=p["rotations"], viewmatrix=self.view, gt_depth=self.gt)
        self.torch.autograd.backward(self._outs(res), self.cots)
        self.last = res

    def grads(self):
        return {k: v.grad for k, v in self.params.items()}

    def _upload(self):
        """Enqueue the H2D copies of ONE step's inputs (camera + RGB-D frame) from pinned host memory on
        the copy stream, into one of two preallocated device input sets (no allocator traffic inside
        the timed region: per-step allocations on a side stream made the caching allocator fall back
        to cudaMalloc now and then, which showed as 5x outliers).
        Returns the device tensors, a completion event and the slot index."""
        torch, dev = self.torch, self.device
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=dev)
            self.dev_in = [[torch.empty_like(h, device=dev) for h in self.h_inputs] for _ in range(2)]
            self.slot_free = [None, None]
            self.uploads = 0
        slot = self.uploads & 1
        self.uploads += 1
        d = self.dev_in[slot]
        with torch.cuda.stream(self.copy_stream):
            if self.slot_free[slot] is not None:      # the step that consumed this set has finished
                self.copy_stream.wait_event(self.slot_free[slot])
            for dt, ht in zip(d, self.h_inputs):
                dt.copy_(ht, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return d, ev, slot

    def step_e2e(self, fused_loss=True):
        """One training step with the step's inputs coming from pinned host memory and its result
        (loss, dL/dviewmatrix) going back to the host, organised like an input pipeline: every step
        uploads one step's inputs (the NEXT step's, on a copy stream, while this step computes) and
        reads one step's result (asynchronously into a pinned double buffer; the host consumes the
        value one step later).  Exactly one upload and one result read-back per step."""
        torch, dev = self.torch, self.device
        main = torch.cuda.current_stream()
        if self.pending is None:
            self.pending = self._upload()
        inp, ev, in_slot = self.pending
        self.pending = self._upload()
        main.wait_event(ev)
        d_view, d_proj, d_campos, d_rgb, d_mm = inp
        view = d_view.detach().requires_grad_(True)   # fresh leaf on the reused storage
        gt_d = (d_mm * 1e-3).unsqueeze(0)             # int16 millimetres -> fp32 metres [1,H,W]
        if not (fused_loss and hasattr(self.mod, "rgbd_l1_loss")):
            gt_rgb = d_rgb * (1.0 / 255.0)            # uint8 -> fp32 [3,H,W]
        rast = self._rasterizer(view.detach(), d_proj, d_campos)
        p = self.params
        res = rast(means3D=p["means3D"], means2D=self.means2D, opacities=p["opacities"], shs=p["shs"],
                   scales=p["scales"], rotations=p["rotations"], viewmatrix=view, gt_depth=gt_d)
        outs = self._outs(res)
        # L1 colour + depth loss (+ median depth and the depth-variance channel for -light, the
        # silhouette for -full): every differentiable output receives a cotangent, as in `step`.
        # The B200 package evaluates it and its cotangents in one pass (its public helper
        # rgbd_l1_loss, straight from the uint8 / int16 frame); the reference package has no such
        # helper, its arm (and the "e2e_torch_loss" leg of ours) use the same loss written in torch.
        helper = getattr(self.mod, "rgbd_l1_loss", None) if fused_loss else None
        if helper is not None:
            loss, tensors, cots = helper(res, d_rgb, d_mm)
            torch.autograd.backward(tensors, cots)
        else:
            loss = (outs[0] - gt_rgb).abs().sum() + (outs[1] - gt_d).abs().sum()
            if self.variant == "light":
                loss = loss + (outs[2] - gt_d).abs().sum() + outs[3].sum()
            else:
                loss = loss + (1.0 - outs[2]).sum()
            loss.backward()
        with torch.no_grad():
            packed = torch.cat([loss.detach().reshape(1), view.grad.reshape(16)])
        slot = self.e2e_steps & 1
        self.e2e_steps += 1
        value = None
        if self.res_ev[slot] is not None:      # result of two steps ago: complete long since
            self.res_ev[slot].synchronize()
            value = float(self.h_results[slot][0])
        self.h_results[slot].copy_(packed, non_blocking=True)
        self.res_ev[slot] = torch.cuda.Event()
        self.res_ev[slot].record(main)
        self.slot_free[in_slot] = self.res_ev[slot]   # this step's device inputs may be overwritten after it
        return value


def timed_region(torch, dist, fn, steps, world):
    """Barrier + sync, time `steps` calls of fn with CUDA events, sync + barrier.
    Returns (total ms, max over ranks; this rank's per-step ms list from one event per step)."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[steps])
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms, per_step


def spread(per_step):
    """median / p10 / p90 of the per-step device times (BASELINE.md protocol)."""
    v = sorted(per_step)
    q = lambda f: v[min(len(v) - 1, max(0, int(round(f * (len(v) - 1)))))]
    return {"median_ms": q(0.5), "p10_ms": q(0.1), "p90_ms": q(0.9), "min_ms": v[0], "max_ms": v[-1]}


def cpu_oracle_sample(ge, cfg_name, variant, frames):
    """Time the CPU oracle port (fwd+bwd) on `frames` frames of the workload.
    Returns (frames/s, seconds, (outputs, gradients) of the last frame)."""
    import parity_util as pu
    sc, cam, scene, cot = build_inputs(ge, cfg_name, variant, "cpu", 0)
    pu.run_oracle(variant, sc.make_camera(64, 48), sc.make_scene(200, sc.make_camera(64, 48)),
                  sc.make_cotangents(sc.make_camera(64, 48), 3 if variant == "light" else 2))  # warm the .so
    t0 = time.perf_counter()
    last = None
    for _ in range(frames):
        last = pu.run_oracle(variant, cam, scene, cot)
    dt = time.perf_counter() - t0
    return frames / dt, dt, last


def parity_report(ge, mod, variant, cam, scene, cot, device, oracle_result, aligned):
    """The B200 arm on the bench tensors against (a) the CPU oracle's outputs of the cpu_baseline leg
    and (b) one frame of the reference CUDA build when baseline/_ref is on the box.  Runs after every
    timed region.  -full's dL/dviewmatrix is left out when the image is not 16-aligned (the reference's
    ComputePG reads uninitialised shared memory in partly-outside tiles)."""
    import torch
    import parity_util as pu
    o_m, g_m = pu.run_variant(mod, variant, cam, scene, cot, device=device)
    drop_view = variant == "full" and not aligned
    rep = {"tolerances": {"fwd_abs": pu.FWD_ATOL, "grad_rel": pu.GRAD_RTOL},
           "dL_dview_compared": not drop_view}

    def one(o_r, g_r, strict):
        g_a, g_b = dict(g_m), dict(g_r)
        if drop_view:
            g_a.pop("viewmatrix", None), g_b.pop("viewmatrix", None)
        st = {}
        ok, lines = pu.compare_runs(o_m, g_a, o_r, g_b, flip_budget=1e-3, grad_budget=1e-2, strict=strict, stats=st,
                                    outliers_ok=not strict)
        st["ok"] = bool(ok)
        st["criterion"] = ("strict: 0 image elements over 1e-4, integers equal, every gradient within 1e-3 of its "
                           "tensor's maximum") if strict else (
            "fp32 CPU port, not bit-identical to either CUDA build: <= 1e-3 of the image elements over 1e-4 "
            "(flipped hard decisions), <= 1e-2 of the gradient elements outside 1e-3")
        if not ok:
            st["report"] = [l for l in lines if "FAIL" in l]
        return st
    if oracle_result is not None:
        rep["vs_oracle"] = one(oracle_result[0], oracle_result[1], strict=False)
    ref = ge.load_reference(variant)
    if ref is not None:
        torch.cuda.empty_cache()
        o_r, g_r = pu.run_variant(ref, variant, cam, scene, cot, device=device)
        rep["vs_reference"] = one(o_r, g_r, strict=True)
        rep["vs_reference"]["color_bit_identical"] = bool((o_m["color"] == o_r["color"]).all())
    best = rep.get("vs_reference") or rep.get("vs_oracle") or {}
    for k in ("fwd_max_abs", "pixels_over", "grad_max_rel"):
        rep[k] = best.get(k)
    rep["against"] = "reference CUDA build" if "vs_reference" in rep else ("CPU oracle" if "vs_oracle" in rep else None)
    return rep


def dp_sum_check(torch, dist, ge, mod, a, frame, reducer, scene, world, rank, device):
    """SURVEY.md 8e test, on hardware: the reduced scene gradients of one N-GPU step against the sum of
    the N single-GPU gradients (rank 0 renders all N views alone, gradient arena detached)."""
    frame.zero_grad()
    frame.step()
    reducer.reduce_async(frame.grads())
    views = reducer.wait()
    torch.cuda.synchronize()
    rep = None
    if rank == 0:
        reduced = {k: v.detach().clone() for k, v in views.items()}
        reducer.detach()
        total = {}
        for v in range(world):
            _sc, cam_v, _scene, cot_v = build_inputs(ge, a.config, a.variant, device, v)
            f = Frame(mod, a.variant, cam_v, scene, cot_v, device, a.track_off, a.map_off, host_inputs=False)
            f.zero_grad()
            f.step()
            torch.cuda.synchronize()
            for k, g in f.grads().items():
                total[k] = g.double() if k not in total else total[k] + g.double()
            del f
            torch.cuda.empty_cache()
        per = {}
        for k, t in total.items():
            r = reduced[k].reshape(t.shape).double()
            per[k] = float((r - t).abs().max() / t.abs().max().clamp_min(1e-30))
        rep = {"mode": reducer.mode, "views": world, "tolerance": 1e-5, "max_rel": max(per.values()),
               "per_param": per, "ok": max(per.values()) <= 1e-5}
        reducer.attach(mod)
    dist.barrier()
    return rep


def run_arm(cfg, variant, impl, steps, warmup, timeout=900):
    """One bench line of another configuration in its own process (keeps library / allocator state
    and the loaded .so set of this process untouched)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--config", cfg, "--variant", variant, "--impl", impl,
           "--steps", str(steps), "--warmup", str(warmup), "--cpu-frames", "0", "--no-extra", "--no-parity"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
        return json.loads(lines[-1]) if lines else {"error": (r.stderr or "no output")[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


EXTRA = (("C2", "light"), ("C3", "light"), ("C4", "full"))


def extra_configs(steps, warmup):
    """Both arms at the other BASELINE.json configurations, so that the driver's record
    carries them: frames/s device-resident and end to end, and the ratio."""
    out = {}
    for cfg, variant in EXTRA:
        # small frames are host-bound and noisy: many more (sub-millisecond) steps there
        k = 10 if cfg in ("C1", "C2") else 1
        ours = run_arm(cfg, variant, "b200", steps * k, warmup * k)
        ref = run_arm(cfg, variant, "reference", max(3, steps // 4) * k, 3 * k)
        row = {}
        if "value" in ours:
            row.update(b200_fps=ours["value"], b200_ms=ours["ms_per_step"], b200_e2e_fps=ours["e2e"]["value"],
                       stages_ms=ours.get("stages_ms_per_step"), stats=ours.get("stats"))
        else:
            row["b200_error"] = ours.get("error")
        if "value" in ref and not ref.get("unavailable"):
            row.update(reference_fps=ref["value"], reference_ms=ref["ms_per_step"],
                       reference_e2e_fps=ref["e2e"]["value"], reference_kind=ref.get("cpu_baseline", {}).get("kind"))
            if "b200_fps" in row:
                row["ratio"] = row["b200_fps"] / ref["value"]
                row["e2e_ratio"] = row["b200_e2e_fps"] / ref["e2e"]["value"]
        else:
            row["reference_error"] = ref.get("error") or ref.get("unavailable")
        out["%s_%s" % (cfg, variant)] = row
    return out


def oracle_threads():
    try:
        out = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-n"], capture_output=True, text=True).stdout
    except Exception:
        out = ""
    omp = "-fopenmp" in out or os.path.exists(os.path.join(ROOT, "oracle", "liboracle_f32.so"))
    try:
        flags = subprocess.run(["sh", "-c", "ldd %s | grep -c gomp" % os.path.join(ROOT, "oracle", "liboracle_f32.so")],
                               capture_output=True, text=True).stdout.strip()
        omp = flags not in ("", "0")
    except Exception:
        pass
    return (os.cpu_count() or 1) if omp else 1


def _private_stdout():
    """Keep the real stdout for the ONE JSON line and point fd 1 at stderr for everything else
    (NCCL and other libraries print banners to stdout)."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


def main():
    out_stream = _private_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--variant", default="full", choices=["light", "full"])
    ap.add_argument("--cpu-frames", type=int, default=2, help="frames of the CPU oracle sample (0 = skip)")
    ap.add_argument("--no-stage-timing", action="store_true", help="skip the second (stage-timed) pass")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity object")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs sub-runs")
    ap.add_argument("--no-dp-check", action="store_true", help="N > 1: skip the gradient-sum check")
    ap.add_argument("--track-off", action="store_true", help="-light only: mapping mode (no pose gradient)")
    ap.add_argument("--map-off", action="store_true", help="-light only: tracking mode (pose gradient only)")
    ap.add_argument("--dp-mode", default="auto", choices=["auto", "allreduce", "factorized_sh", "nvls", "p2p"],
                    help="gradient exchange at N > 1 (diff-gaussian-rasterization_b200/dp.py)")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (gsr_set_option)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        log("warning: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE" % (a.gpus, world))
    sc_mod = ge.load_scene_module()
    P, W, H, _sig = sc_mod.CONFIGS[a.config]
    workload = "%s: %d Gaussians, %dx%d, -%s variant, SH degree 3, fwd+bwd incl. dL/dviewmatrix" % (
        a.config, P, W, H, a.variant)
    if a.track_off or a.map_off:
        workload += " [%s]" % ("tracking: map_off" if a.map_off else "mapping: track_off")
    # `config` holds only what defines the workload (identical for both arms); everything measured or
    # implementation-dependent goes to `stats`
    config = {"workload": workload, "variant": a.variant, "gaussians": P, "width": W, "height": H,
              "parallelism": ("view-dp%d: one view per GPU, one exchange of the scene-parameter gradients per step"
                              % world) if world > 1 else "single GPU",
              "l2": "inputs larger than L2 (%d MB of scene parameters are streamed every step; 126 MB L2)"
                    % ((59 * 4 * P) >> 20)}

    have_ref = (a.impl == "reference" and torch.cuda.is_available()
                and ge.load_reference(a.variant) is not None)
    if a.impl == "reference" and not have_ref:
        # No reference CUDA build on this box: the CPU oracle port, rank 0 only.
        if rank != 0:
            return
        frames = max(1, a.cpu_frames)
        fps, dt, _ = cpu_oracle_sample(ge, a.config, a.variant, frames)
        cores = oracle_threads()
        sample = "%d full frames of the workload through the CPU oracle port (fwd+bwd), %.1f s" % (frames, dt)
        print(file=out_stream, flush=True, *[json.dumps({
            "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus,
            "steps": frames, "warmup": 0, "ms_per_step": 1000.0 / fps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "note": "baseline/_ref (reference CUDA build) not on this box; the reference has no CPU path, "
                    "this is the oracle port",
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})])
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback of the product path)"
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))

    mod = ge.load_reference(a.variant) if a.impl == "reference" else ge.load_variant(a.variant)
    sc, cam, scene, cot = build_inputs(ge, a.config, a.variant, device, rank)
    frame = Frame(mod, a.variant, cam, scene, cot, device, a.track_off, a.map_off, view_seed=rank)

    reducer = None
    if world > 1:
        dp = ge.load_dp_module()
        shapes = {k: tuple(v.shape) for k, v in frame.params.items()}
        want = a.dp_mode if a.dp_mode != "auto" else dp.DEFAULT_MODE
        dp_mode = want if a.impl == "b200" else "allreduce"   # the reference has no masked colour output
        reducer = dp.SceneGradReducer(shapes, device, mode=dp_mode, means3D=frame.params["means3D"], sh_degree=3)
        zero_copy = reducer.attach(mod)   # B200 arm: backward writes into the flat buffer directly
        log("rank %d: exchange mode %s (requested %s), gradient arena attached: %s %s" % (
            rank, reducer.mode, dp_mode, zero_copy, getattr(reducer, "nvls_note", None) or ""))

    lib = None
    if a.impl == "b200":
        lib = ctypes.CDLL(ge.core_library_path())
        lib.gsr_stage_name.restype = ctypes.c_char_p
        lib.gsr_launch_count.restype = ctypes.c_longlong
        for kv in a.opt:
            k, v = kv.split("=")
            assert lib.gsr_set_option(k.encode(), int(v)) >= 0, "unknown option " + k

    def step():
        frame.zero_grad()
        frame.step()
        if reducer is not None:
            reducer.reduce_async(frame.grads())
            reducer.wait()

    def step_e2e(fused_loss=True):
        frame.zero_grad()
        frame.step_e2e(fused_loss)
        if reducer is not None:
            reducer.reduce_async(frame.grads())
            reducer.wait()

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    num_rendered = num_related = None
    try:
        # the forward's ctx is gone; take the counts from one raw forward call
        E = torch.Tensor([])
        p = frame.params
        args = [scene.bg.to(device), p["means3D"].detach(), E, p["opacities"].detach(), p["scales"].detach(),
                p["rotations"].detach(), 1.0, E, frame.view.detach(), frame.gt, cam.projmatrix.to(device),
                cam.tanfovx, cam.tanfovy, cam.H, cam.W, p["shs"].detach(), 3, cam.campos.to(device), False]
        if a.variant == "light":
            r = mod._C.rasterize_gaussians(*args, False)
            num_rendered = int(r[0])
        else:
            # NG (valid (pixel, Gaussian) pairs) is only counted on request: one probe with "exact_ng"
            old_ng = lib.gsr_set_option(b"exact_ng", 1) if lib is not None else None
            try:
                r = mod._C.rasterize_gaussians(*args)
                num_rendered, num_related = int(r[0]), int(r[1])
            finally:
                if lib is not None and old_ng is not None and old_ng >= 0:
                    lib.gsr_set_option(b"exact_ng", old_ng)
        del r
    except Exception as e:  # informational only
        log("count probe failed: %r" % (e,))

    # ---- timed region 1: device-resident, no library timers -> `value` ---------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    if lib is not None:
        lib.gsr_launch_count(1)
    total_ms, per_step = timed_region(torch, dist, step, a.steps, world)
    launches = int(lib.gsr_launch_count(1)) if lib is not None else 0
    # ---- timed region 2: the same K steps with the library's per-stage CUDA events on the launching
    # stream -> stage times / roofline; its total against region 1 is the stage timers' overhead -------
    stage_ms, stage_scopes, stage_launch, staged_total_ms = {}, {}, {}, None
    if lib is not None and not a.no_stage_timing:
        lib.gsr_set_option(b"stage_timing", 1)
        lib.gsr_stage_times(None, None, None, 1)
        staged_total_ms, _ = timed_region(torch, dist, step, a.steps, world)
        n = 10
        ms = (ctypes.c_double * n)()
        scopes = (ctypes.c_longlong * n)()
        launches_a = (ctypes.c_longlong * n)()
        lib.gsr_stage_times(ms, scopes, launches_a, 1)
        lib.gsr_set_option(b"stage_timing", 0)
        for i in range(n):
            name = lib.gsr_stage_name(i).decode()
            if scopes[i]:
                stage_ms[name] = ms[i] / a.steps
                stage_scopes[name] = scopes[i]
                stage_launch[name] = launches_a[i]
    # ---- timed region 3: end to end (host buffers) ------------------------------------------------------
    for _ in range(3):
        step_e2e()
    e2e_ms, e2e_per_step = timed_region(torch, dist, step_e2e, a.steps, world)
    e2e_torch_ms = None
    if a.impl == "b200" and hasattr(mod, "rgbd_l1_loss"):
        # the same end-to-end step with the loss written in torch, exactly as the reference arm runs it
        for _ in range(2):
            step_e2e(False)
        e2e_torch_ms, _ = timed_region(torch, dist, lambda: step_e2e(False), a.steps, world)
    clk = clocks.stop() if rank == 0 else None

    if reducer is not None and getattr(reducer, "nvls", None) and rank == 0:
        tt = reducer.nvls_timing()
        if tt:
            log("nvls exchange phases (ms): %s" % (tt,))
    dp_check = None
    if world > 1 and a.impl == "b200" and not a.no_dp_check:
        dp_check = dp_sum_check(torch, dist, ge, mod, a, frame, reducer, scene, world, rank, device)
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / a.steps
    value = world * 1000.0 / ms_per_step
    e2e_value = world * 1000.0 / (e2e_ms / a.steps)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    sort_bits = 32 + max(1, (tiles).bit_length())
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "step_ms": spread(per_step),
        "stats": {"num_rendered": num_rendered, "num_related": num_related,
                  "mean_tile_list": (num_rendered / float(tiles)) if num_rendered else None,
                  "mean_valid_contributors_per_pixel": (num_related / float(W * H)) if num_related else None,
                  "exchange": ("%s%s: %d MB of scene gradients per rank and step" % (
                      reducer.mode, (" (slice all-reduce: %s)" % reducer.nvls["slice"]) if getattr(reducer, "nvls", None) else "",
                      reducer.bytes_per_step() >> 20)) if reducer is not None else None},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": frame.h2d_bytes,
                "d2h_bytes_per_step": frame.d2h_bytes, "ms_per_step": e2e_ms / a.steps,
                "step_ms": spread(e2e_per_step),
                "inputs": "camera (3 tensors) + ground-truth RGB-D frame (uint8 colour, int16 mm depth) from "
                          "pinned host memory; L1 loss + cotangents on the device (B200 arm: the package's "
                          "one-pass rgbd_l1_loss helper; reference arm: the same loss in torch); loss + "
                          "dL/dviewmatrix read back"},
        "clocks": clk,
    }
    if e2e_torch_ms is not None:
        out["e2e_torch_loss"] = {"value": world * 1000.0 / (e2e_torch_ms / a.steps), "unit": UNIT,
                                 "ms_per_step": e2e_torch_ms / a.steps,
                                 "note": "the same end-to-end step with the loss evaluated by torch ops, as in "
                                         "the reference arm"}
    if dp_check is not None:
        out["dp_check"] = dp_check
    if a.impl == "reference":
        out["impl"] = "reference"
        out["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                               "sample": "the reference has no CPU path: this is its own CUDA build for sm_100 "
                                         "(baseline/_ref), %d full frames on the GPU" % a.steps}
        out["gpu_launches"] = 0
        out["note"] = "unmodified reference CUDA sources (baseline/build_ref.sh), stock Python API"
    else:
        out["impl"] = "b200"
        out["gpu_launches"] = launches
        tile_local = lib.gsr_get_option(b"tile_sort") != 0
        label = lambda k: STAGE_LABEL_TILE.get(k, k) if tile_local else k
        nb = lambda k: stage_bytes(k, P, num_rendered or 0, W * H, tiles, 16, a.variant, sort_bits, tile_local)
        if staged_total_ms is not None:
            out["stage_timing"] = {"ms_per_step_with_stage_timers": staged_total_ms / a.steps,
                                   "overhead_ms_per_step": staged_total_ms / a.steps - ms_per_step,
                                   "note": "stage times come from a second pass of the same K steps with the "
                                           "library's CUDA-event stage timers on; `value` is the pass without them"}
        out["stages_ms_per_step"] = {label(k): v for k, v in stage_ms.items()}
        if stage_ms:
            top = max((k for k in stage_ms if nb(k)), key=lambda k: stage_ms[k])
            nbytes = nb(top)
            launches_per_step = max(1, stage_scopes[top] // a.steps)
            dur_ms = stage_ms[top] / launches_per_step
            peak, peak_src = 6650.0, "fallback"
            try:
                peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
                peak_src = "measured"
            except Exception:
                pass
            achieved = nbytes / (dur_ms * 1e-3) / 1e9
            traffic = inst = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                traffic = tj.get(top, {}).get(a.config)
                inst = tj.get("inst_executed", {}).get(top, {}).get(a.config)
            except Exception:
                pass
            out["roofline"] = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                               "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                               "algorithmic_bytes": nbytes, "kernel_ms": dur_ms,
                               "share_of_step": stage_ms[top] / (staged_total_ms / a.steps),
                               "note": "blend kernels are FP32-issue / MUFU / L2-atomic bound, not HBM bound "
                                       "(DESIGN.md); the HBM fraction is reported because the metric asks for it"}
            if inst and clk and clk.get("sm_mhz"):
                # the bound that actually limits the blend kernels: warp-instruction issue slots
                # (148 SMs x 4 schedulers x SM clock); instruction count from the committed ncu capture
                peak_ips = 148 * 4 * float(clk["sm_mhz"]) * 1e6
                out["roofline"]["issue"] = {"warp_instructions": inst, "peak_per_s": peak_ips,
                                            "achieved_per_s": inst / (dur_ms * 1e-3),
                                            "frac": inst / (dur_ms * 1e-3) / peak_ips}
            # every stage against the HBM roofline (bytes the stage moves / measured stage time)
            out["roofline_by_stage"] = {
                label(k): {"ms": stage_ms[k], "algorithmic_bytes": nb(k),
                           "achieved_gbs": nb(k) / (stage_ms[k] * 1e-3) / 1e9,
                           "frac": nb(k) / (stage_ms[k] * 1e-3) / 1e9 / peak}
                for k in stage_ms if nb(k)}
            frame_bytes = sum(nb(k) or 0 for k in ("preprocess_fwd", "scan", "emit_keys", "radix_sort", "tile_ranges",
                                                   "render_fwd", "render_bwd", "preprocess_bwd"))
            ref_bin = reference_binning_bytes(P, num_rendered or 0, tiles, sort_bits)
            out["frame_hbm"] = {"algorithmic_bytes": frame_bytes, "achieved_gbs": frame_bytes / (ms_per_step * 1e-3) / 1e9,
                                "frac_of_measured_peak": frame_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                                "frac_of_8TBs": frame_bytes / (ms_per_step * 1e-3) / 8e12,
                                "reference_model_binning_bytes": ref_bin,
                                "note": "binning is charged with the bytes the tile-local path moves; the "
                                        "reference's scan + emit + 6-pass radix sort + ranges would move "
                                        "reference_model_binning_bytes for the same N"}
        oracle_last = None
        if world == 1 and a.cpu_frames > 0:
            fps, dt, oracle_last = cpu_oracle_sample(ge, a.config, a.variant, a.cpu_frames)
            out["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": oracle_threads(), "kind": "port",
                                   "sample": "%d full frames of the workload (fwd+bwd) through the CPU oracle, "
                                             "%.1f s; the reference has no CPU implementation" % (a.cpu_frames, dt)}
        if world == 1 and not a.no_parity and not (a.track_off or a.map_off):
            try:
                out["parity"] = parity_report(ge, mod, a.variant, cam, scene, cot, device, oracle_last,
                                              aligned=(W % 16 == 0 and H % 16 == 0))
            except Exception as e:  # noqa: BLE001
                out["parity"] = {"error": repr(e)[:300]}
        if world == 1 and not a.no_extra and a.config == "C3" and a.variant == "full":
            del frame
            torch.cuda.empty_cache()
            out["extra_configs"] = extra_configs(20, 5)
    print(json.dumps(out), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
