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
        # exactly the parameter's strides, also where they are ambiguous (a channels_last 1x1 conv weight is "contiguous"
        # with strides [Ci,1,Ci,Ci]): the fused optimizers require params, grads and state to have identical strides
        expect = 1
        for st, sz in sorted((st, sz) for sz, st in zip(p.shape, p.stride()) if sz != 1):
            if st != expect:                               # not a permutation of a dense block: plain contiguous view
                return flat[off:off + p.numel()].view_as(p)
            expect *= sz
        return flat[off:off + p.numel()].as_strided(p.size(), p.stride())

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


class OverlappedGradReducer:
    """Gradient averaging overlapped with the backward pass (what PL's DDP does for the reference, train.py:272).

    Bucket 0 holds the decoder parameters: their gradients are all final when the fused decoder backward returns, i.e.
    BEFORE the encoder backward starts, so its all-reduce travels over NVLink while cuDNN runs the trunk's backward.  The
    encoder parameters follow in REVERSE order (autograd produces the last layers' gradients first) in buckets of
    ~`bucket_bytes`, tapering to 16 / 4 / 1 MB at the stem end (only the last bucket's all-reduce cannot hide under backward).
    A post-accumulate-grad hook per parameter counts its bucket down; when the last gradient has landed the bucket's gradient
    tensors are gathered into its flat fp32 buffer by ONE multi-tensor copy, `.grad` of every parameter becomes a view (with
    the parameter's strides) into that buffer, and ONE asynchronous all-reduce (ReduceOp.AVG on NCCL: no scaling pass) is issued.

    Two designs were measured and dropped (BASELINE configs[1] / [2], N=2): (a) pre-assigned, pre-zeroed bucket views as `.grad`
    -- autograd then launches one accumulation kernel per parameter instead of handing its result over (+1.1 ms per 32 ms
    step for ~180 parameters); (b) no flat buffers, a coalesced NCCL group over the raw gradient tensors -- fastest in steady
    state (+0.4 ms) but with ~330 operations per step at configs[2] the host sporadically stalls 100-400 ms inside NCCL.

        reducer.prepare(); loss.backward(); reducer.finish(); optimizer.step()
    """

    def __init__(self, dec_params, enc_params, bucket_bytes=32 << 20, group=None):
        self.group = group
        dec = [p for p in dec_params if p.requires_grad]
        enc = [p for p in enc_params if p.requires_grad][::-1]
        self.buckets = [dec] if dec else []
        caps = [min(c << 20, bucket_bytes) for c in (1, 4, 16)]
        rev, cur, cur_n = [], [], 0
        for p in enc[::-1]:                                   # forward (registration) order: stem first
            n = p.numel()
            cap = caps[len(rev)] if len(rev) < len(caps) else bucket_bytes
            if cur and (cur_n + n) * 4 > cap:
                rev.append(cur[::-1])
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += n
        if cur:
            rev.append(cur[::-1])
        self.buckets += rev[::-1]                             # backward order again: last layers first, stem last
        self.flat, self.views = [], []
        for b in self.buckets:
            flat = torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device)
            views, off = [], 0
            for p in b:
                views.append(FlatGradBuckets._view(flat, off, p))
                off += p.numel()
            self.flat.append(flat)
            self.views.append(views)
        self._pending = [0] * len(self.buckets)
        self._handles = [None] * len(self.buckets)
        self._hooks = []
        self._armed = False
        backend = dist.get_backend(group) if dist.is_initialized() else "none"
        self._avg = backend == "nccl"
        self._world = dist.get_world_size(group) if dist.is_initialized() else 1
        for bi, b in enumerate(self.buckets):
            for p in b:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _make_hook(self, bi):
        def hook(param):
            if not self._armed:
                return
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch(bi)
        return hook

    def _launch(self, bi):
        if self._handles[bi] is not None:
            return
        src, dst = [], []
        for p, v in zip(self.buckets[bi], self.views[bi]):
            if p.grad is None:
                v.zero_()                                     # no gradient this step (must be so on every rank)
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)                    # one multi-tensor kernel per bucket
        for p, v in zip(self.buckets[bi], self.views[bi]):
            if p.grad is not None:
                p.grad = v
        if self._world <= 1:
            self._handles[bi] = ()
            return
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._handles[bi] = (dist.all_reduce(self.flat[bi], op=op, group=self.group, async_op=True),)

    def prepare(self):
        """before backward(): gradients back to None (autograd hands over its result tensors) and arm the hooks"""
        for bi, b in enumerate(self.buckets):
            for p in b:
                p.grad = None
            self._pending[bi] = len(b)
            self._handles[bi] = None
        self._armed = True

    def finish(self):
        """after backward(): reduce buckets whose hooks did not all fire (parameters without a gradient this step), wait for
        every all-reduce"""
        self._armed = False
        for bi in range(len(self.buckets)):
            if self._handles[bi] is None:
                self._launch(bi)
        for bi, hs in enumerate(self._handles):
            for h in hs or ():
                h.wait()
            if hs and not self._avg:
                self.flat[bi].mul_(1.0 / self._world)

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
