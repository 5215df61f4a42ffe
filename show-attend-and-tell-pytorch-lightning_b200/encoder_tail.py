"""Encoder tail: the resize layer readme.md:118-121 appends to the encoder (nn.Upsample((s,s), mode="bilinear",
align_corners=False)) as ONE pass of the library over the channels_last map -- the result is physically the [B,L,D] annotation
array the decoder kernels stream.  Under bf16 autocast the stock layer runs as cast-to-fp32 + upsample_bilinear2d<float> + a later
cast to bf16 (three passes, fp32 intermediate), and its backward scatters with atomics; this one interpolates the stored values in
fp32, rounds once and has a deterministic gather-form backward (csrc/sat_encoder_tail.cu).  No parameters: state_dict unchanged."""
import torch
from torch import nn

from . import _lib


class _ResizeNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H2, W2):
        n, D, h, w = x.shape
        y = torch.empty((n, D, H2, W2), dtype=x.dtype, device=x.device).contiguous(memory_format=torch.channels_last)
        _lib.check(_lib.lib().sat_resize_nhwc_fwd(x.data_ptr(), y.data_ptr(), n, h, w, H2, W2, D, _lib.dtype_code(x.dtype), _lib.stream_ptr()),
                   "sat_resize_nhwc_fwd")
        ctx.shape = (n, D, h, w, H2, W2)
        ctx.dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        n, D, h, w, H2, W2 = ctx.shape
        if dy.dtype != ctx.dtype:
            dy = dy.to(ctx.dtype)
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((n, D, h, w), dtype=dy.dtype, device=dy.device).contiguous(memory_format=torch.channels_last)
        _lib.check(_lib.lib().sat_resize_nhwc_bwd(dy.data_ptr(), dx.data_ptr(), n, h, w, H2, W2, D, _lib.dtype_code(dy.dtype), _lib.stream_ptr()),
                   "sat_resize_nhwc_bwd")
        return dx, None, None


class ResizeBilinearNHWC(nn.Upsample):
    """nn.Upsample(size, mode="bilinear", align_corners=False) whose CUDA channels_last bf16 / fp32 inputs take the library's
    one-pass kernels; anything else (CPU, NCHW, other dtypes, channel counts that are not 16-byte vectors) the stock layer."""

    def forward(self, x):
        vec = 8 if x.dtype == torch.bfloat16 else 4
        if (x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float32) and x.shape[1] % vec == 0 and x.numel() > 0
                and x.is_contiguous(memory_format=torch.channels_last) and self.mode == "bilinear" and not self.align_corners
                and self.size is not None):
            H2, W2 = (self.size, self.size) if isinstance(self.size, int) else tuple(self.size)
            return _ResizeNHWC.apply(x, int(H2), int(W2))
        return super().forward(x)
