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
  "nvls"  the factorized exchange done INSIDE the backward's per-Gaussian kernel over NVSwitch
      multicast: the gradient arena is symmetric memory (torch.distributed._symmetric_memory), the kernel
      adds the 11 reduced floats of every visible Gaussian with multimem.red to the arena's multicast
      alias (the switch reduces; every rank's replica ends up with the sum) and multicast-stores the
      masked colour gradient + camera position into its slot of every replica.  No collective call: what
      is left per step is a memset of the other buffer of the double-buffered arena, one device-side
      barrier and the local SH rebuild.  Falls back to "factorized_sh" when multicast is unavailable.

With `attach()` the backward of the B200-native packages writes its gradients straight into the
flat buffer (no packing copy).  The collectives run on a side stream; `wait()` makes the reduced
gradients visible to the current stream.  Works with any torch.distributed backend: NCCL over NVLink
on the B200 box, gloo on CPU (tests; the CPU path of the SH rebuild below is a plain-torch
restatement used by those tests only).
"""
from collections import OrderedDict

import torch
import torch.distributed as dist

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
        assert mode in ("allreduce", "factorized_sh", "nvls")
        self.group, self.average, self.mode = group, average, mode
        self.nvls = None
        if mode == "nvls":
            self.mode = mode = "factorized_sh"   # layout and SH rebuild are those of the factorized exchange
            self._nvls_requested = True
        else:
            self._nvls_requested = False
        self.is_cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.is_cuda else None
        self._work, self._done, self._attached = None, None, None
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

    # ---- NVLS (in-switch) exchange -----------------------------------------------------------------
    def _setup_nvls(self, device):
        """Allocate the double-buffered symmetric arena; on any failure stay on the NCCL path."""
        self.nvls_note = None
        try:
            if not (self.is_cuda and dist.is_available() and dist.is_initialized()) or self._world() < 2:
                raise RuntimeError("needs CUDA and an initialised process group with world_size > 1")
            import torch.distributed._symmetric_memory as symm_mem
            world, rank = self._world(), dist.get_rank(self.group)
            group = self.group if self.group is not None else dist.group.WORLD
            numel = world * self.head + 11 * self.P
            arenas, handles = [], []
            for _ in range(2):
                t = symm_mem.empty(numel, dtype=torch.float32, device=torch.device(device))
                h = symm_mem.rendezvous(t, group)
                if not int(h.multicast_ptr):
                    raise RuntimeError("no multicast support on this system")
                t.zero_()
                arenas.append(t)
                handles.append(h)
            torch.cuda.synchronize(device)
            handles[0].barrier(channel=0, timeout_ms=30000)
            torch.cuda.synchronize(device)
            off = world * self.head
            red = OrderedDict()
            for name, n, shape in (("means3D", 3 * self.P, (self.P, 3)), ("opacities", self.P, (self.P, 1)),
                                   ("scales", 3 * self.P, (self.P, 3)), ("rotations", 4 * self.P, (self.P, 4))):
                red[name] = (off, n, shape)
                off += n
            self.nvls = dict(arenas=arenas, handles=handles, cur=0, done=None, world=world, rank=rank, slices=red)
            self.mode = "nvls"
        except Exception as e:  # noqa: BLE001 - any failure means: use the NCCL exchange
            self.nvls = None
            self.nvls_note = "nvls unavailable (%s: %s); using factorized_sh" % (type(e).__name__, e)

    def _nvls_register(self):
        n = self.nvls
        self._attached._C.set_grad_arena_nvls(n["arenas"][n["cur"]], int(n["handles"][n["cur"]].multicast_ptr),
                                              n["rank"], n["world"])

    def _nvls_exchange(self):
        """Runs on the side stream after the backward: zero the other buffer, barrier, rebuild dL/dsh."""
        n = self.nvls
        cur, nxt = n["cur"], 1 - n["cur"]
        n["arenas"][nxt].zero_()
        n["handles"][cur].barrier(channel=0, timeout_ms=30000)
        gathered = n["arenas"][cur][:n["world"] * self.head].view(n["world"], self.head)
        self.sh_sum = self._attached._C.sh_grad_from_views(self.means3D.detach(), gathered, self.sh_degree, self.M)
        if self.average:
            n["arenas"][cur][n["world"] * self.head:].div_(n["world"])
            self.sh_sum.div_(n["world"])
        n["done"] = cur
        n["cur"] = nxt

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
        if self.mode == "nvls":
            self._nvls_register()
        else:
            fn(self.flat, self.mode == "factorized_sh")
        return True

    def detach(self):
        if self._attached is not None:
            self._attached._C.set_grad_arena(torch.Tensor(), False)
            self._attached = None

    def views(self):
        """Per-parameter reduced gradients (valid after wait())."""
        if self.mode == "nvls":
            n = self.nvls
            buf = n["arenas"][n["done"] if n["done"] is not None else n["cur"]]
            out = {k: buf[o:o + m].view(shape) for k, (o, m, shape) in n["slices"].items()}
            out["shs"] = self.sh_sum
            return out
        out = {k: self.flat[o:o + n].view(shape) for k, (o, n, shape) in self.slices.items()}
        if self.mode == "factorized_sh":
            out["shs"] = self.sh_sum
        return out

    def pack(self, grads, masked_color=None, campos=None):
        for k, (o, n, _shape) in self.slices.items():
            g = grads.get(k)
            self.flat[o:o + n].copy_(g.reshape(-1) if g is not None else 0)
        if self.mode == "factorized_sh":
            self.flat[:3 * self.P].copy_(masked_color.reshape(-1))
            self.flat[3 * self.P:3 * self.P + 3].copy_(campos.reshape(-1)[:3])

    def _aliases_flat(self, grads):
        if self.mode == "nvls":
            n = self.nvls
            lo = n["arenas"][n["cur"]].data_ptr()
            return all(grads.get(k) is not None and grads[k].data_ptr() == lo + 4 * o
                       for k, (o, _m, _s) in n["slices"].items())
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
        exchange.  Returns immediately on CUDA; call wait() before reading views()."""
        if self.mode == "nvls":
            if grads is not None and not self._aliases_flat(grads):
                raise RuntimeError("nvls exchange: the gradients were not produced into the symmetric arena "
                                   "(attach() the rasterizer package and call wait() after every reduce_async())")
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self._nvls_exchange()
                self._done = torch.cuda.Event()
                self._done.record(self.stream)
            self._nvls_register()   # the next backward writes into the other buffer
            return
        if grads is not None and not (self._attached is not None and self._aliases_flat(grads)):
            self.pack(grads, masked_color, campos)
        if self.is_cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self._exchange()
                self._done = torch.cuda.Event()
                self._done.record(self.stream)
        else:
            self._exchange()

    def wait(self):
        if self.is_cuda and self._done is not None:
            torch.cuda.current_stream().wait_event(self._done)
            self._done = None
        return self.views()

    def bytes_per_step(self):
        """fp32 bytes this rank contributes to the exchange per step."""
        if self.mode == "nvls":
            return (14 * self.P + 4) * 4
        return self.numel * 4
