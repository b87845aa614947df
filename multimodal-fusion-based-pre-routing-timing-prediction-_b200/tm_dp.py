"""Design-level data parallelism: shard designs across ranks, all-reduce gradients.

The reference trains one design at a time on one device (train.py:461); designs are independent
samples, so the only exchange is one gradient SUM per step (11.6 MB of fp32).  ``GradBucket``
flattens a group of gradients into one buffer, posts the NCCL all-reduce asynchronously on a side
stream (gloo on CPU for tests) and scatters the averaged values back into ``param.grad``.
Parameters without a gradient (``fc_net_drive``, ``fc_attn2``: never used, D12) are skipped --
consistently on every rank because the set is structural.
"""
import torch
import torch.distributed as dist


def shard_designs(designs, rank, world):
    """Round-robin: rank r takes designs r, r+world, ..."""
    return list(designs[rank::world])


class GradBucket:
    def __init__(self, params, world, group=None):
        self.params = [p for p in params if p.grad is not None]
        self.world, self.group = world, group
        self.flat = None

    def post(self, comm_stream):
        """Flatten on the current stream, all-reduce on ``comm_stream`` (None: current / CPU)."""
        if not self.params:
            return None
        self.flat = torch.cat([p.grad.reshape(-1) for p in self.params])
        if comm_stream is None:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ev)
            self.flat.record_stream(comm_stream)
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self, work):
        if work is None:
            return
        work.wait()
        self.flat.mul_(1.0 / self.world)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].reshape(p.shape)
            off += n
