"""Encoder boundary: batch-norm layers of the torchvision trunk on cuDNN's NHWC (persistent) kernels for bf16 / fp16.

The CNN encoder stays on cuDNN through PyTorch (it is the boundary of the hot path, not the target).  PyTorch's dispatcher,
however, does not hand bf16 batch-norm to cuDNN: a channels_last bf16 ResNet runs ATen's
`batch_norm_*_channels_last_kernel`s, which are 57 % of the whole BASELINE configs[1] training step on B200.  cuDNN's
`cudnnBatchNormalization{ForwardTraining,Backward}Ex` in CUDNN_BATCHNORM_SPATIAL_PERSISTENT mode accepts NHWC bf16 and is
3-4x faster on the same tensors (measured with tools/cudnn_bn_probe.py: forward 160 vs 635 us, backward 198 vs 585 us at
[128,256,56,56]).  This module calls that library entry point directly (ctypes on the libcudnn that PyTorch itself loaded):

    convert_batchnorm(trunk)   # nn.BatchNorm2d -> CudnnBatchNorm2d, same parameters / buffers / state_dict keys

Training-mode forward and backward of 4-D channels_last half-precision inputs go to cuDNN; everything else (eval mode,
fp32, NCHW) falls back to the stock nn.BatchNorm2d forward.  Running statistics follow nn.BatchNorm2d (momentum, unbiased
running variance, num_batches_tracked).
"""
import ctypes as C
import glob
import os
import site

import torch
from torch import nn

_NHWC, _FLOAT, _HALF, _BF16 = 1, 0, 2, 9
_PERSISTENT, _OPS_BN = 2, 0
_vp = C.c_void_p


class _Cudnn:
    """libcudnn handle + descriptor cache of this process (one process per GPU)."""
    lib = None
    handles = {}
    plans = {}
    failed = False

    @classmethod
    def load(cls):
        if cls.lib is not None or cls.failed:
            return cls.lib
        cands = []
        for sp in list(site.getsitepackages()) + [site.getusersitepackages()]:
            cands += glob.glob(os.path.join(sp, "nvidia", "cudnn", "lib", "libcudnn.so.9"))
        cands += ["libcudnn.so.9", "libcudnn.so"]
        for f in cands:
            try:
                cls.lib = C.CDLL(f, mode=C.RTLD_GLOBAL)
                break
            except OSError:
                continue
        if cls.lib is None:
            cls.failed = True
            return None
        cls.lib.cudnnGetErrorString.restype = C.c_char_p
        return cls.lib

    @classmethod
    def check(cls, rc, what):
        if rc != 0:
            raise RuntimeError("cuDNN %s failed: %d %s" % (what, rc, cls.lib.cudnnGetErrorString(rc).decode()))

    @classmethod
    def handle(cls, device):
        h = cls.handles.get(device.index)
        if h is None:
            h = _vp()
            cls.check(cls.lib.cudnnCreate(C.byref(h)), "cudnnCreate")
            cls.handles[device.index] = h
        cls.check(cls.lib.cudnnSetStream(h, _vp(torch.cuda.current_stream(device).cuda_stream)), "cudnnSetStream")
        return h

    @classmethod
    def plan(cls, h, shape, dtype, device):
        """tensor descriptors and workspace sizes of one (N,C,H,W,dtype) configuration"""
        key = (device.index, tuple(shape), dtype)
        p = cls.plans.get(key)
        if p is not None:
            return p
        lib = cls.lib
        n, c, hh, ww = shape
        xd, bd = _vp(), _vp()
        cls.check(lib.cudnnCreateTensorDescriptor(C.byref(xd)), "cudnnCreateTensorDescriptor")
        cls.check(lib.cudnnSetTensor4dDescriptor(xd, _NHWC, _BF16 if dtype == torch.bfloat16 else _HALF, n, c, hh, ww), "cudnnSetTensor4dDescriptor")
        cls.check(lib.cudnnCreateTensorDescriptor(C.byref(bd)), "cudnnCreateTensorDescriptor")
        cls.check(lib.cudnnDeriveBNTensorDescriptor(bd, xd, _PERSISTENT), "cudnnDeriveBNTensorDescriptor")
        wf, wb, rs = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        cls.check(lib.cudnnGetBatchNormalizationForwardTrainingExWorkspaceSize(h, _PERSISTENT, _OPS_BN, xd, None, xd, bd, None, C.byref(wf)),
                  "ForwardTrainingExWorkspaceSize")
        cls.check(lib.cudnnGetBatchNormalizationBackwardExWorkspaceSize(h, _PERSISTENT, _OPS_BN, xd, None, xd, None, xd, bd, None, C.byref(wb)),
                  "BackwardExWorkspaceSize")
        cls.check(lib.cudnnGetBatchNormalizationTrainingExReserveSpaceSize(h, _PERSISTENT, _OPS_BN, None, xd, C.byref(rs)), "ReserveSpaceSize")
        p = (xd, bd, int(wf.value), int(wb.value), int(rs.value))
        cls.plans[key] = p
        return p


_ONE, _ZERO = C.c_float(1.0), C.c_float(0.0)


class _CudnnBNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps):
        dev = x.device
        lib = _Cudnn.lib
        h = _Cudnn.handle(dev)
        xd, bd, wf, wb, rs = _Cudnn.plan(h, x.shape, x.dtype, dev)
        y = torch.empty_like(x)                         # channels_last, like x
        C_ = x.shape[1]
        save_mean = torch.empty(C_, dtype=torch.float32, device=dev)
        save_invstd = torch.empty(C_, dtype=torch.float32, device=dev)
        ws = torch.empty(max(wf, 16), dtype=torch.uint8, device=dev)
        reserve = torch.empty(max(rs, 16), dtype=torch.uint8, device=dev)
        rm = running_mean.data_ptr() if running_mean is not None else None
        rv = running_var.data_ptr() if running_var is not None else None
        _Cudnn.check(lib.cudnnBatchNormalizationForwardTrainingEx(
            h, _PERSISTENT, _OPS_BN, C.byref(_ONE), C.byref(_ZERO), xd, _vp(x.data_ptr()), None, None, xd, _vp(y.data_ptr()), bd,
            _vp(weight.data_ptr()), _vp(bias.data_ptr()), C.c_double(momentum), _vp(rm), _vp(rv), C.c_double(eps),
            _vp(save_mean.data_ptr()), _vp(save_invstd.data_ptr()), None, _vp(ws.data_ptr()), C.c_size_t(wf), _vp(reserve.data_ptr()),
            C.c_size_t(rs)), "cudnnBatchNormalizationForwardTrainingEx")
        ctx.save_for_backward(x, weight, save_mean, save_invstd, reserve)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, save_mean, save_invstd, reserve = ctx.saved_tensors
        dev = x.device
        lib = _Cudnn.lib
        h = _Cudnn.handle(dev)
        xd, bd, wf, wb, rs = _Cudnn.plan(h, x.shape, x.dtype, dev)
        if dy.dtype != x.dtype or not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        dw = torch.empty_like(weight)
        db = torch.empty_like(weight)
        ws = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
        _Cudnn.check(lib.cudnnBatchNormalizationBackwardEx(
            h, _PERSISTENT, _OPS_BN, C.byref(_ONE), C.byref(_ZERO), C.byref(_ONE), C.byref(_ZERO), xd, _vp(x.data_ptr()), None, None, xd,
            _vp(dy.data_ptr()), None, None, xd, _vp(dx.data_ptr()), bd, _vp(weight.data_ptr()), None, _vp(dw.data_ptr()), _vp(db.data_ptr()),
            C.c_double(ctx.eps), _vp(save_mean.data_ptr()), _vp(save_invstd.data_ptr()), None, _vp(ws.data_ptr()), C.c_size_t(wb),
            _vp(reserve.data_ptr()), C.c_size_t(rs)), "cudnnBatchNormalizationBackwardEx")
        return dx, dw, db, None, None, None, None


class CudnnBatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d whose training-mode forward / backward of channels_last bf16 / fp16 inputs runs on cuDNN's NHWC
    persistent batch-norm.  Parameters, buffers and state_dict keys are those of nn.BatchNorm2d."""

    def forward(self, x):
        use = (self.training and x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float16) and self.affine
               and self.weight.dtype == torch.float32 and x.shape[1] % 4 == 0 and x.numel() > 0
               and x.is_contiguous(memory_format=torch.channels_last) and _Cudnn.load() is not None)
        if not use:
            return super().forward(x)
        if self.momentum is None:                        # cumulative moving average
            if self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
                factor = 1.0 / float(self.num_batches_tracked)
            else:
                factor = 0.0
        else:
            factor = self.momentum
            if self.track_running_stats and self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
        rm = self.running_mean if self.track_running_stats else None
        rv = self.running_var if self.track_running_stats else None
        if not torch.is_grad_enabled():                  # no graph: still the cuDNN forward, nothing saved
            with torch.no_grad():
                return _CudnnBNFunction.apply(x, self.weight, self.bias, rm, rv, factor, self.eps)
        return _CudnnBNFunction.apply(x, self.weight, self.bias, rm, rv, factor, self.eps)


def convert_batchnorm(module):
    """Replaces every nn.BatchNorm2d under `module` by a CudnnBatchNorm2d sharing its parameters and buffers (in place).
    Returns the number of layers converted."""
    n = 0
    for name, child in list(module.named_children()):
        if type(child) is nn.BatchNorm2d:
            new = CudnnBatchNorm2d(child.num_features, eps=child.eps, momentum=child.momentum, affine=child.affine,
                                   track_running_stats=child.track_running_stats)
            new.weight, new.bias = child.weight, child.bias
            new.running_mean, new.running_var, new.num_batches_tracked = child.running_mean, child.running_var, child.num_batches_tracked
            new.train(child.training)
            setattr(module, name, new)
            n += 1
        else:
            n += convert_batchnorm(child)
    return n
