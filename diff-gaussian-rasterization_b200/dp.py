"""View-level data parallelism for the rasterizer (new work: the reference has no multi-GPU code).

One process per GPU; every rank holds the same replicated scene parameters and renders its own
camera view(s) forward+backward.  The only exchange step is ONE all-reduce (sum) of the
scene-parameter gradients, packed into a single flat fp32 buffer

    [ dL/dmeans3D (3P) | dL/dsh (3MP) | dL/dopacity (P) | dL/dscales (3P) | dL/drotations (4P) ]

= 59 floats = 236 bytes per Gaussian at M = 16.  Per-view results (images, radii,
dL/dviewmatrix, dL/dmeans2D) are NOT reduced.  The all-reduce runs on a side stream so that the
caller can overlap it with the next view's forward; `wait()` makes the reduced gradients visible
to the current stream.

Works with any torch.distributed backend: NCCL over NVLink on the B200 box, gloo on CPU (tests).
"""
from collections import OrderedDict

import torch
import torch.distributed as dist

PARAM_ORDER = ("means3D", "shs", "opacities", "scales", "rotations")


def shard_views(num_views, rank, world_size):
    """Views assigned to `rank`: view v -> rank v mod world_size."""
    return list(range(rank, num_views, world_size))


class SceneGradReducer:
    def __init__(self, shapes, device, group=None, average=False):
        """shapes: mapping name -> shape for the entries of PARAM_ORDER that exist."""
        self.group = group
        self.average = average
        self.slices = OrderedDict()
        off = 0
        for name in PARAM_ORDER:
            if name in shapes and shapes[name] is not None:
                n = 1
                for s in shapes[name]:
                    n *= int(s)
                self.slices[name] = (off, n, tuple(int(s) for s in shapes[name]))
                off += n
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.is_cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.is_cuda else None
        self._work = None
        self._done = None

    def views(self):
        """Per-parameter views into the flat buffer (valid after wait())."""
        return {k: self.flat[o:o + n].view(shape) for k, (o, n, shape) in self.slices.items()}

    def pack(self, grads):
        for k, (o, n, _shape) in self.slices.items():
            g = grads[k]
            self.flat[o:o + n].copy_(g.reshape(-1) if g is not None else 0)

    def attach(self, rasterizer_module):
        """Register the flat buffer as the gradient arena of a B200-native rasterizer package
        (`_C.set_grad_arena`): its backward then writes the five scene gradients straight into
        the buffer and reduce_async() needs no packing copy.  Returns False (and changes nothing)
        for packages without that extension, e.g. the reference build."""
        fn = getattr(getattr(rasterizer_module, "_C", None), "set_grad_arena", None)
        if fn is None or not self.is_cuda:
            return False
        fn(self.flat)
        self._attached = rasterizer_module
        return True

    def detach(self):
        mod = getattr(self, "_attached", None)
        if mod is not None:
            mod._C.set_grad_arena(torch.Tensor())
            self._attached = None

    def _aliases_flat(self, grads):
        lo = self.flat.data_ptr()
        for k, (o, _n, _shape) in self.slices.items():
            g = grads.get(k)
            if g is None or g.data_ptr() != lo + 4 * o:
                return False
        return True

    def reduce_async(self, grads=None):
        """Pack (unless the gradients already live in the flat buffer, see attach()) and launch
        the single all-reduce.  Returns immediately; call wait() before reading views()."""
        if grads is not None and not self._aliases_flat(grads):
            self.pack(grads)
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            self._work = None
            return
        if self.is_cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                if self.average:
                    self.flat.div_(dist.get_world_size(self.group))
                self._done = torch.cuda.Event()
                self._done.record(self.stream)
        else:
            self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def wait(self):
        if self.is_cuda:
            if self._done is not None:
                torch.cuda.current_stream().wait_event(self._done)
                self._done = None
        elif self._work is not None:
            self._work.wait()
            self._work = None
            if self.average:
                self.flat.div_(dist.get_world_size(self.group))
        return self.views()

    def bytes_per_step(self):
        return self.numel * 4
