"""View-level data parallelism for the rasterizer (new work: the reference has no multi-GPU code).

One process per GPU; every rank holds the same replicated scene parameters and renders its own
camera view(s) forward+backward.  The only exchange step is the sum over ranks of the
scene-parameter gradients; per-view results (images, radii, dL/dviewmatrix, dL/dmeans2D) are NOT
reduced.  Two exchange modes:

  "allreduce"  (default)  ONE all-reduce (sum) of a single flat fp32 buffer
        [ dL/dmeans3D (3P) | dL/dsh (3MP) | dL/dopacity (P) | dL/dscales (3P) | dL/drotations (4P) ]
      = 59 floats = 236 bytes per Gaussian at M = 16.
  "factorized_sh"  dL/dsh of a view is the rank-1 product basis_k(view direction) x masked dL/dcolor,
      so ranks all-gather 3 floats per Gaussian (+ the camera position) and all-reduce only the other
      11 floats; every rank rebuilds the summed dL/dsh locally (`_C.sh_grad_from_views`).  Same result
      up to summation order, 14 instead of 59 floats per Gaussian on the wire.
  "nvls"  the factorized exchange with our own kernels over NVSwitch instead of NCCL calls: the flat
      buffer lives in symmetric memory (torch.distributed._symmetric_memory).  After one device-side
      barrier every rank (a) all-reduces its 1/world slice of the 11 reduced floats IN THE SWITCH
      (multimem.ld_reduce pulls and adds the replicas, multimem.st broadcasts the sums:
      gsr_nvls_allreduce_slice) and (b) rebuilds the summed dL/dsh with a kernel that reads the peers'
      masked colour gradients in place through P2P loads (gsr_sh_grad_from_view_ptrs) — the all-gather
      is that kernel's input stream, no gathered copy is written or re-read.  A second barrier ends the
      step.  Falls back to "factorized_sh" when multicast / symmetric memory is unavailable.
      (A push variant — multimem.red issued by the backward's per-Gaussian kernel itself — was built
      and measured first: correct, 4 % faster than NCCL at 2 GPUs, but multimem.red multicasts every
      operand to every replica, so the ingress grows with the number of ranks: 9 % slower at 8.)

With `attach()` the backward of the B200-native packages writes its gradients straight into the
flat buffer (no packing copy).  The collectives run on a side stream; `wait()` makes the reduced
gradients visible to the current stream.  Works with any torch.distributed backend: NCCL over NVLink
on the B200 box, gloo on CPU (tests; the CPU path of the SH rebuild below is a plain-torch
restatement used by those tests only).
"""
from collections import OrderedDict

import torch
import torch.distributed as dist

# Exchange used by bench.py / callers that do not choose: our own kernels over NVSwitch when symmetric
# memory + multicast are available (fastest at 4 and 8 GPUs), else the NCCL factorized path.
DEFAULT_MODE = "nvls"

PARAM_ORDER = ("means3D", "shs", "opacities", "scales", "rotations")
FACT_ORDER = ("means3D", "opacities", "scales", "rotations")  # all-reduced part of the factorized layout

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435)


def shard_views(num_views, rank, world_size):
    """Views assigned to `rank`: view v -> rank v mod world_size."""
    return list(range(rank, num_views, world_size))


def sh_basis(dirs, degree):
    """Real SH basis values [P, (degree+1)^2] for unit directions [P,3] (forward.cu:20-71)."""
    x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    b = [torch.full_like(x, SH_C0)]
    if degree > 0:
        b += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if degree > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if degree > 2:
        b += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
              SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy),
              SH_C3[5] * z * (xx - yy), SH_C3[6] * x * (xx - 3 * yy)]
    return torch.stack(b, dim=1)


def sh_grad_from_views_torch(means3D, gathered, degree, M):
    """Plain-torch restatement of gsr_sh_grad_from_views (CPU tests only)."""
    P = means3D.shape[0]
    out = torch.zeros(P, M, 3, dtype=torch.float32, device=means3D.device)
    nc = (degree + 1) ** 2
    for v in range(gathered.shape[0]):
        dR = gathered[v, :3 * P].view(P, 3)
        campos = gathered[v, 3 * P:3 * P + 3]
        d = means3D - campos
        d = d / d.norm(dim=1, keepdim=True)
        out[:, :nc, :] += sh_basis(d, degree)[:, :, None] * dR[:, None, :]
    return out


class SceneGradReducer:
    def __init__(self, shapes, device, group=None, average=False, mode="allreduce", means3D=None,
                 sh_degree=3):
        """shapes: mapping name -> shape for the entries of PARAM_ORDER that exist.
        mode "factorized_sh" additionally needs the (replicated) means3D parameter tensor."""
        assert mode in ("allreduce", "factorized_sh", "nvls", "p2p")
        self.group, self.average, self.mode = group, average, mode
        self.nvls = None
        # "nvls": slice all-reduce in the switch (multimem) at 8 ranks and more, over plain P2P loads / stores
        # below that (measured, tools/exchange_bench.py); "p2p": always the P2P slice all-reduce (no multicast
        # needed).  GSR_DP_SLICE=nvls|p2p overrides the choice.
        self._slice_kind = None if mode == "nvls" else ("p2p" if mode == "p2p" else None)
        if mode in ("nvls", "p2p"):
            self.mode = mode = "factorized_sh"   # layout and SH rebuild are those of the factorized exchange
            self._nvls_requested = True
        else:
            self._nvls_requested = False
        self.is_cuda = torch.device(device).type == "cuda"
        # high priority: the exchange kernels are small and mostly wait for NVLink; when they are launched
        # underneath the per-Gaussian backward kernel they must not queue behind its CTAs
        self.stream = torch.cuda.Stream(device=device, priority=-1) if self.is_cuda else None
        self._work, self._done, self._attached = None, None, None
        import os
        # Early gather (the masked colour gradients leave the backward before its per-Gaussian kernel and are
        # gathered by a P2P copy kernel underneath it): built and measured — equal at 2 ranks, 12 % / 30 % SLOWER at
        # 4 / 8 (profiles/r02_scale_early_gather.txt: a copy kernel small enough not to disturb the per-Gaussian
        # kernel is latency-bound on NVLink, and one more barrier is paid).  Off unless GSR_DP_EARLY_MIN_WORLD is set.
        self._early, self._no_early = False, bool(os.environ.get("GSR_DP_NO_EARLY"))
        self._early_min_world = int(os.environ.get("GSR_DP_EARLY_MIN_WORLD", str(1 << 30)))
        self._gather_blocks = int(os.environ.get("GSR_DP_GATHER_BLOCKS", "74"))
        self.slices = OrderedDict()
        num = lambda shape: int(torch.Size(shape).numel())
        if mode == "allreduce":
            off = 0
            for name in PARAM_ORDER:
                if shapes.get(name) is not None:
                    self.slices[name] = (off, num(shapes[name]), tuple(int(s) for s in shapes[name]))
                    off += num(shapes[name])
            self.numel = off
        else:
            assert means3D is not None and shapes.get("shs") is not None
            self.means3D, self.sh_degree = means3D, int(sh_degree)
            self.P = P = int(shapes["means3D"][0])
            self.M = int(shapes["shs"][1])
            self.head = 3 * P + 4           # masked colour gradient + camera position (+ pad): all-gathered
            off = self.head
            for name in FACT_ORDER:
                self.slices[name] = (off, num(shapes[name]), tuple(int(s) for s in shapes[name]))
                off += num(shapes[name])
            self.numel = off                # 14 P + 4
            self.sh_shape = tuple(int(s) for s in shapes["shs"])
            self.sh_sum = None
            self.gathered = None
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        if self._nvls_requested:
            self._setup_nvls(device)

    # ---- NVLS exchange (our kernels over NVSwitch: in-switch reduce + P2P-reading SH rebuild) -----------
    def _setup_nvls(self, device):
        """Move the flat buffer into symmetric memory; on any failure stay on the NCCL path."""
        self.nvls_note = None
        try:
            if not (self.is_cuda and dist.is_available() and dist.is_initialized()) or self._world() < 2:
                raise RuntimeError("needs CUDA and an initialised process group with world_size > 1")
            if self.P % 4 != 0:
                raise RuntimeError("needs a Gaussian count that is a multiple of 4")
            import torch.distributed._symmetric_memory as symm_mem
            world, rank = self._world(), dist.get_rank(self.group)
            if world > 16:
                raise RuntimeError("at most 16 ranks")
            group = self.group if self.group is not None else dist.group.WORLD
            flat = symm_mem.empty(self.numel, dtype=torch.float32, device=torch.device(device))
            hdl = symm_mem.rendezvous(flat, group)
            import os
            kind = os.environ.get("GSR_DP_SLICE") or self._slice_kind or ("nvls" if world >= 8 else "p2p")
            if kind == "nvls" and not int(hdl.multicast_ptr):
                if self._slice_kind is None and not os.environ.get("GSR_DP_SLICE"):
                    kind = "p2p"
                else:
                    raise RuntimeError("no multicast support on this system")
            self._slice_kind = kind
            flat.zero_()
            torch.cuda.synchronize(device)
            hdl.barrier(channel=0, timeout_ms=30000)
            torch.cuda.synchronize(device)
            self.flat = flat          # same layout as "factorized_sh": [dR 3P | campos 3 | pad | reduced 11P]
            self.nvls = dict(handle=hdl, world=world, rank=rank, mc=int(hdl.multicast_ptr),
                             peers=[int(p) for p in hdl.buffer_ptrs], slice=kind)
            if os.environ.get("GSR_DP_TIMING"):   # per-phase CUDA events, see nvls_timing()
                self.nvls["timing"] = []
            self.mode = "nvls"
        except Exception as e:  # noqa: BLE001 - any failure means: use the NCCL exchange
            self.nvls = None
            self.nvls_note = "nvls unavailable (%s: %s); using factorized_sh" % (type(e).__name__, e)

    def nvls_timing(self):
        """Mean milliseconds of (barrier A, launch of the slice all-reduce on its own stream, SH rebuild
        overlapped with that all-reduce, barrier B) over the recorded steps (GSR_DP_TIMING=1), skipping
        the first five."""
        t = (self.nvls or {}).get("timing") or []
        torch.cuda.synchronize()
        rows = [[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in t[5:]]
        return [sum(r[i] for r in rows) / max(1, len(rows)) for i in range(4)] if rows else None

    def _nvls_exchange(self, cur, early):
        """On self.stream.  `cur`: the stream the backward ran on.  early: self.stream has been ordered after the
        masked colour gradient only (the per-Gaussian backward kernel may still be running on `cur`); else after
        the whole backward."""
        n, C = self.nvls, self._attached._C
        timing = n.get("timing")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timing is not None else None
        mark = (lambda i: ev[i].record()) if ev else (lambda i: None)
        side = n.get("stream2")
        if side is None:
            side = n["stream2"] = torch.cuda.Stream(device=self.flat.device, priority=-1)
        me = torch.cuda.current_stream()
        mark(0)
        n["handle"].barrier(channel=0, timeout_ms=30000)        # every rank's masked colour gradient is in place
        mark(1)
        if early:
            # all-gather of the masked colour gradients (+ camera positions) by P2P loads UNDERNEATH the
            # per-Gaussian backward kernel: a pure copy bound by the NVLink ports, a few CTAs are enough
            if self.gathered is None or self.gathered.shape[0] != n["world"]:
                self.gathered = torch.empty(n["world"], self.head, dtype=torch.float32, device=self.flat.device)
            C.p2p_gather(n["peers"], self.head, self.gathered, self._gather_blocks)
            me.wait_stream(cur)                                  # this rank's per-Gaussian backward kernel is done
            n["handle"].barrier(channel=2, timeout_ms=30000)     # ... and every rank's: the 11 reduced floats are in place
        # the slice all-reduce (bound by the NVLink ports) and the SH rebuild (early: from the gathered copy, bound
        # by HBM; else reading the peers in place) run side by side on two streams
        side.wait_stream(me)
        with torch.cuda.stream(side):
            # the time does not depend on the grid (tools/exchange_bench.py): a few CTAs are enough, and
            # they leave the SMs to the SH rebuild
            if n["slice"] == "p2p":
                C.p2p_allreduce_slice(n["peers"], self.head, 11 * self.P, n["rank"], 37)
            else:
                C.nvls_allreduce_slice(n["mc"], self.head, 11 * self.P, n["rank"], n["world"], 32)
        mark(2)
        if early:
            self.sh_sum = C.sh_grad_from_views(self.means3D.detach(), self.gathered, self.sh_degree, self.M)
        else:
            self.sh_sum = C.sh_grad_from_view_ptrs(self.means3D.detach(), n["peers"],
                                                   [p + 4 * 3 * self.P for p in n["peers"]], self.sh_degree, self.M)
        me.wait_stream(side)
        mark(3)
        n["handle"].barrier(channel=1, timeout_ms=30000)        # all slices broadcast, all peer reads done
        mark(4)
        if ev:
            timing.append(ev)
        if self.average:
            self.flat[self.head:].div_(n["world"])
            self.sh_sum.div_(n["world"])

    # ---- wiring -------------------------------------------------------------------------------
    def attach(self, rasterizer_module):
        """Register the flat buffer as the gradient arena of a B200-native rasterizer package
        (`_C.set_grad_arena`): its backward then writes the gradients straight into the buffer.
        Returns False (and changes nothing) for packages without that extension (the reference)."""
        fn = getattr(getattr(rasterizer_module, "_C", None), "set_grad_arena", None)
        if fn is None or not self.is_cuda:
            if self.mode == "nvls":      # the reference package cannot do it: NCCL exchange instead
                self.mode, self.nvls = "factorized_sh", None
            return False
        self._attached = rasterizer_module
        self._early = False
        if (self.mode == "nvls" and hasattr(rasterizer_module._C, "p2p_gather") and not self._no_early and
                self.nvls["world"] >= self._early_min_world):
            # the masked colour gradient leaves the backward before its per-Gaussian kernel runs: the peers'
            # reads of it (the SH rebuild) overlap that kernel
            fn(self.flat, True, True)
            self._early = True
        else:
            fn(self.flat, self.mode in ("factorized_sh", "nvls"))
        return True

    def detach(self):
        if self._attached is not None:
            self._attached._C.set_grad_arena(torch.Tensor(), False)
            self._attached = None

    def views(self):
        """Per-parameter reduced gradients (valid after wait())."""
        out = {k: self.flat[o:o + n].view(shape) for k, (o, n, shape) in self.slices.items()}
        if self.mode in ("factorized_sh", "nvls"):
            out["shs"] = self.sh_sum
        return out

    def pack(self, grads, masked_color=None, campos=None):
        for k, (o, n, _shape) in self.slices.items():
            g = grads.get(k)
            if g is not None and g.data_ptr() == self.flat.data_ptr() + 4 * o:
                continue                                   # already in place
            self.flat[o:o + n].copy_(g.reshape(-1) if g is not None else 0)
        if self.mode in ("factorized_sh", "nvls") and masked_color is not None:
            # (with attach() the armed backward has already written the head of the buffer)
            self.flat[:3 * self.P].copy_(masked_color.reshape(-1))
            self.flat[3 * self.P:3 * self.P + 3].copy_(campos.reshape(-1)[:3])

    def _aliases_flat(self, grads):
        lo = self.flat.data_ptr()
        for k, (o, _n, _shape) in self.slices.items():
            g = grads.get(k)
            if g is None or g.data_ptr() != lo + 4 * o:
                return False
        return True

    # ---- the exchange -------------------------------------------------------------------------
    def _world(self):
        if not dist.is_available() or not dist.is_initialized():
            return 1
        return dist.get_world_size(self.group)

    def _exchange(self):
        world = self._world()
        if self.mode == "allreduce":
            if world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                self.flat.div_(world)
            return
        head = self.flat[:self.head]
        if self.gathered is None or self.gathered.shape[0] != world:
            self.gathered = torch.empty(world, self.head, dtype=torch.float32, device=self.flat.device)
        if world > 1:
            dist.all_gather_into_tensor(self.gathered.view(-1), head, group=self.group)
            dist.all_reduce(self.flat[self.head:], op=dist.ReduceOp.SUM, group=self.group)
        else:
            self.gathered[0].copy_(head)
        if self.is_cuda and self._attached is not None:
            self.sh_sum = self._attached._C.sh_grad_from_views(self.means3D.detach(), self.gathered,
                                                               self.sh_degree, self.M)
        elif self.is_cuda:
            raise RuntimeError("factorized_sh on CUDA needs attach() to a B200-native rasterizer package")
        else:
            self.sh_sum = sh_grad_from_views_torch(self.means3D.detach(), self.gathered, self.sh_degree, self.M)
        if self.average:
            self.flat[self.head:].div_(world)
            self.sh_sum.div_(world)

    def reduce_async(self, grads=None, masked_color=None, campos=None):
        """Pack (unless the gradients already live in the flat buffer, see attach()) and launch the
        exchange.  Returns immediately on CUDA; call wait() before reading views() AND before the
        next backward (the re-armed arena is overwritten by it).

        Several backwards per exchange (several views per rank, a tracking pass on the same scene)
        are fine: the arena takes the first one, the others return fresh tensors that autograd adds
        to .grad — in place into the arena when .grad aliases it, else they are packed here.  In the
        factorized modes the first backward leaves no dL/dsh (it is rebuilt from the masked colour
        gradients); the dL/dsh of any further backward arrives in grads["shs"] and is all-reduced and
        added on top.  Parameters must enter the step with .grad = None (zero_grad's default): a
        zero-filled .grad that aliases the arena would be added to itself by autograd."""
        if self.mode == "nvls" and self._attached is None:
            raise RuntimeError("nvls exchange needs attach() to a B200-native rasterizer package")
        if grads is not None and not (self._attached is not None and self._aliases_flat(grads)):
            self.pack(grads, masked_color, campos)
        extra_sh = grads.get("shs") if (grads is not None and self.mode in ("factorized_sh", "nvls")) else None
        if self.is_cuda:
            cur = torch.cuda.current_stream()
            early = bool(self.mode == "nvls" and self._early and extra_sh is None and
                         self._attached._C.wait_masked_color(self.stream.cuda_stream))
            if not early:
                self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                if self.mode == "nvls":
                    self._nvls_exchange(cur, early)
                else:
                    self._exchange()
                if extra_sh is not None:
                    self._add_extra_sh(extra_sh)
                self._done = torch.cuda.Event()
                self._done.record(self.stream)
            if self._attached is not None:
                self._attached._C.arm_grad_arena()   # the next step's first backward takes the arena again
        else:
            self._exchange()
            if extra_sh is not None:
                self._add_extra_sh(extra_sh)

    def _add_extra_sh(self, extra_sh):
        t = extra_sh.detach().clone() if self.is_cuda else extra_sh.detach().clone()
        if self._world() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        if self.average:
            t.div_(self._world())
        self.sh_sum = self.sh_sum + t.view(self.sh_sum.shape)

    def wait(self):
        if self.is_cuda and self._done is not None:
            torch.cuda.current_stream().wait_event(self._done)
            self._done = None
        return self.views()

    def bytes_per_step(self):
        """fp32 bytes this rank contributes to the exchange per step."""
        return self.numel * 4
