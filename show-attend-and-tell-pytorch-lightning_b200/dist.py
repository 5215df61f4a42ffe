"""Data-parallel plumbing for the training step: one process per GPU, gradients of all trainable
parameters averaged with NCCL all-reduce over NVLink/NVSwitch (what PL's implicit DDP does for the
reference, train.py:272).  Gradients are laid out in a few large flat buckets so each all-reduce is
bandwidth- rather than launch-bound; decoder parameters come first (their grads are final before
the encoder backward starts), so their bucket can be reduced while cuDNN is still busy.
"""
import torch
import torch.distributed as dist


class FlatGradBuckets:
    """Re-homes `.grad` of the given parameters into contiguous fp32 buckets and all-reduces them."""

    def __init__(self, params, bucket_bytes=64 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []
        cur, cur_n = [], 0
        for p in self.params:
            n = p.numel()
            if cur and (cur_n + n) * 4 > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += n
        if cur:
            self.buckets.append(cur)
        self.flat = []
        for b in self.buckets:
            dev = b[0].device
            flat = torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=dev)
            off = 0
            for p in b:
                p.grad = self._view(flat, off, p)
                off += p.numel()
            self.flat.append(flat)

    @staticmethod
    def _view(flat, off, p):
        """bucket slice with the parameter's own strides (channels_last conv weights keep their layout, so autograd's
        gradient-layout contract holds and no per-step transposes are inserted)."""
        try:
            if p.is_contiguous() or not p.is_non_overlapping_and_dense():
                return flat[off:off + p.numel()].view_as(p)
            return flat[off:off + p.numel()].as_strided(p.size(), p.stride())
        except Exception:
            return flat[off:off + p.numel()].view_as(p)

    def zero(self):
        for f in self.flat:
            f.zero_()

    def rebind(self):
        """make sure every param's .grad is still the bucket view (autograd accumulates in place)."""
        for b, flat in zip(self.buckets, self.flat):
            off = 0
            for p in b:
                v = self._view(flat, off, p)
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    if p.grad is not None:
                        v.copy_(p.grad)
                    p.grad = v
                off += p.numel()

    def allreduce_mean(self, world_size, group=None):
        if world_size <= 1:
            return
        self.rebind()
        handles = [dist.all_reduce(f, op=dist.ReduceOp.SUM, group=group, async_op=True) for f in self.flat]
        for h in handles:
            h.wait()
        inv = 1.0 / world_size
        for f in self.flat:
            f.mul_(inv)
