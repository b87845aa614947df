"""Design-level data parallelism: shard designs across ranks, all-reduce gradients.

The reference trains one design at a time on one device (train.py:461); designs are independent
samples, so the only exchange is one gradient SUM per step (11.6 MB of fp32), posted in two buckets:
head + fusion (9 MB) as soon as those gradients exist, so that the NCCL transfer over NVLink overlaps the
remaining backward kernels, and GNN + U-Net (2.5 MB) as one exchange after the last kernel.

``FlatBucket`` owns ONE persistent flat fp32 buffer per group, allocated on first use.  Every step:

1. the freshly computed gradients are copied into the buffer's per-parameter views by one
   multi-tensor copy (``torch._foreach_copy_``) -- no ``torch.cat``, no allocation;
2. the all-reduce (SUM) and the 1/world scaling run on the communication stream, fenced by events;
3. ``param.grad`` IS the view, so nothing is scattered back.

Because the buffers are persistent the whole exchange can be captured inside the step's CUDA graph
(``DesignStep.capture``), where it overlaps the backward exactly as in the eager schedule.
Parameters without a gradient (``fc_net_drive``, ``fc_attn2``: never used, D12) are skipped --
consistently on every rank because the set is structural.  CPU tensors (gloo) are supported for tests.
"""
import torch
import torch.distributed as dist


def shard_designs(designs, rank, world):
    """Round-robin: rank r takes designs r, r+world, ..."""
    return list(designs[rank::world])


class FlatBucket:
    def __init__(self, params, world, group=None):
        self.params = list(params)
        self.world, self.group = world, group
        self.flat = None
        self.views = None
        self.ready = None            # CUDA event: the averaged gradients are in the buffer

    def _ensure(self, grads):
        if self.flat is not None:
            return
        n = sum(g.numel() for g in grads)
        self.flat = torch.zeros(n, dtype=torch.float32, device=grads[0].device)
        self.views, off = [], 0
        for p, g in zip(self.params, grads):
            self.views.append(self.flat[off:off + g.numel()].view(p.shape))
            off += g.numel()

    def post(self, comm_stream=None):
        """Pack ``param.grad`` of every parameter into the flat buffer on the current stream, then all-reduce
        and scale on ``comm_stream`` (None: current stream / CPU).  ``param.grad`` is re-pointed at the views."""
        grads = [p.grad for p in self.params]
        if not grads:
            return
        self._ensure(grads)
        torch._foreach_copy_(self.views, [g.reshape(v.shape) for g, v in zip(grads, self.views)])
        for p, v in zip(self.params, self.views):
            p.grad = v
        if comm_stream is None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)
            return
        cur = torch.cuda.current_stream()
        comm_stream.wait_stream(cur)
        with torch.cuda.stream(comm_stream):
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)

    def finish(self, comm_stream=None):
        """Make the current stream wait for the exchange posted on ``comm_stream``."""
        if comm_stream is not None and self.flat is not None:
            torch.cuda.current_stream().wait_stream(comm_stream)


# kept for callers of the first interface (tests): flatten / all-reduce / scatter in one call
class GradBucket(FlatBucket):
    def __init__(self, params, world, group=None):
        super().__init__([p for p in params if p.grad is not None], world, group)

    def post(self, comm_stream=None):          # noqa: D102
        super().post(comm_stream)
        return comm_stream

    def finish(self, work=None):               # noqa: D102
        super().finish(work if isinstance(work, torch.cuda.Stream) else None)
