"""Encoder boundary: batch-norm (+ReLU, +residual add) and max-pool layers of the torchvision trunk on cuDNN's NHWC kernels
for bf16 / fp16.

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

The same cuDNN entry point also fuses what follows the normalisation in a residual block (bnOps BN_ACTIVATION and
BN_ADD_ACTIVATION): y = relu(bn(x)) and y = relu(bn(x) + z) cost what the plain batch-norm costs (tools/cudnn_fused_probe.py:
412 us against 1537 us for ATen's bn + add + relu at [256,256,56,56], backward 455 against 1501 us), which removes the ReLU
forward / backward and residual-add kernels of the eager trunk (5.9 ms of the 36.8 ms BASELINE configs[1] step, profiles/
r02_full_train_step_launches.txt).  fuse_residual_blocks(trunk) switches torchvision's Bottleneck / BasicBlock (and the
conv-bn-relu-maxpool stem) to these calls; the stem's max-pool goes to cudnnPooling{Forward,Backward} (249 / 1484 us against
ATen's 748 / 2005 us at [256,64,112,112]).  In inference (eval mode under no_grad) the same blocks fold the batch-norm into the
convolution weights and call cuDNN's fused convolution + bias (+ residual) + ReLU (torch.cudnn_convolution_relu /
cudnn_convolution_add_relu; tools/conv_fuse_probe.py: 61-160 us against 285-820 us for conv + bn + add + relu at the 56x56 layers),
so the captioning trunk runs without a single element-wise pass.  Parameters, buffers, state_dict keys and fp32 behaviour are
unchanged.
"""
import ctypes as C
import glob
import os
import site

import torch
from torch import nn

_NHWC, _FLOAT, _HALF, _BF16 = 1, 0, 2, 9
_PERSISTENT, _OPS_BN, _OPS_BN_ACT, _OPS_BN_ADD_ACT = 2, 0, 1, 2
_vp = C.c_void_p


class _Cudnn:
    """libcudnn handle + descriptor cache of this process (one process per GPU)."""
    lib = None
    handles = {}
    plans = {}
    failed = False
    act = None

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
    def tensor_desc(cls, shape, dtype):
        n, c, hh, ww = shape
        d = _vp()
        cls.check(cls.lib.cudnnCreateTensorDescriptor(C.byref(d)), "cudnnCreateTensorDescriptor")
        cls.check(cls.lib.cudnnSetTensor4dDescriptor(d, _NHWC, _BF16 if dtype == torch.bfloat16 else _HALF, n, c, hh, ww), "cudnnSetTensor4dDescriptor")
        return d

    @classmethod
    def relu_desc(cls):
        if cls.act is None:
            a = _vp()
            cls.check(cls.lib.cudnnCreateActivationDescriptor(C.byref(a)), "cudnnCreateActivationDescriptor")
            cls.check(cls.lib.cudnnSetActivationDescriptor(a, 1, 0, C.c_double(0.0)), "cudnnSetActivationDescriptor")   # RELU, NOT_PROPAGATE_NAN
            cls.act = a
        return cls.act

    @classmethod
    def plan(cls, h, shape, dtype, device, ops=_OPS_BN):
        """tensor descriptors and workspace sizes of one (N,C,H,W,dtype,bnOps) configuration"""
        key = (device.index, tuple(shape), dtype, ops)
        p = cls.plans.get(key)
        if p is not None:
            return p
        lib = cls.lib
        xd, bd = cls.tensor_desc(shape, dtype), _vp()
        cls.check(lib.cudnnCreateTensorDescriptor(C.byref(bd)), "cudnnCreateTensorDescriptor")
        cls.check(lib.cudnnDeriveBNTensorDescriptor(bd, xd, _PERSISTENT), "cudnnDeriveBNTensorDescriptor")
        act = cls.relu_desc() if ops != _OPS_BN else None
        zd = xd if ops == _OPS_BN_ADD_ACT else None
        yd = xd if ops != _OPS_BN else None                    # backward of the fused ops reads the output
        wf, wb, rs = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        cls.check(lib.cudnnGetBatchNormalizationForwardTrainingExWorkspaceSize(h, _PERSISTENT, ops, xd, zd, xd, bd, act, C.byref(wf)),
                  "ForwardTrainingExWorkspaceSize")
        cls.check(lib.cudnnGetBatchNormalizationBackwardExWorkspaceSize(h, _PERSISTENT, ops, xd, yd, xd, zd, xd, bd, act, C.byref(wb)),
                  "BackwardExWorkspaceSize")
        cls.check(lib.cudnnGetBatchNormalizationTrainingExReserveSpaceSize(h, _PERSISTENT, ops, act, xd, C.byref(rs)), "ReserveSpaceSize")
        p = (xd, bd, int(wf.value), int(wb.value), int(rs.value), act)
        cls.plans[key] = p
        return p

    @classmethod
    def pool_plan(cls, xshape, yshape, dtype, device, k, pad, stride):
        key = ("pool", device.index, tuple(xshape), dtype, k, pad, stride)
        p = cls.plans.get(key)
        if p is None:
            pd = _vp()
            cls.check(cls.lib.cudnnCreatePoolingDescriptor(C.byref(pd)), "cudnnCreatePoolingDescriptor")
            cls.check(cls.lib.cudnnSetPooling2dDescriptor(pd, 3, 0, k[0], k[1], pad[0], pad[1], stride[0], stride[1]),     # MAX_DETERMINISTIC
                      "cudnnSetPooling2dDescriptor")
            p = cls.plans[key] = (pd, cls.tensor_desc(xshape, dtype), cls.tensor_desc(yshape, dtype))
        return p


_ONE, _ZERO = C.c_float(1.0), C.c_float(0.0)


class _CudnnBNFunction(torch.autograd.Function):
    """y = bn(x) (ops 0), relu(bn(x)) (ops 1) or relu(bn(x) + z) (ops 2), training mode, NHWC half precision"""

    @staticmethod
    def forward(ctx, x, z, weight, bias, running_mean, running_var, momentum, eps, ops):
        dev = x.device
        lib = _Cudnn.lib
        h = _Cudnn.handle(dev)
        xd, bd, wf, wb, rs, act = _Cudnn.plan(h, x.shape, x.dtype, dev, ops)
        y = torch.empty_like(x)                         # channels_last, like x
        C_ = x.shape[1]
        save_mean = torch.empty(C_, dtype=torch.float32, device=dev)
        save_invstd = torch.empty(C_, dtype=torch.float32, device=dev)
        ws = torch.empty(max(wf, 16), dtype=torch.uint8, device=dev)
        reserve = torch.empty(max(rs, 16), dtype=torch.uint8, device=dev)
        rm = running_mean.data_ptr() if running_mean is not None else None
        rv = running_var.data_ptr() if running_var is not None else None
        zd, zp = (xd, _vp(z.data_ptr())) if ops == _OPS_BN_ADD_ACT else (None, None)
        _Cudnn.check(lib.cudnnBatchNormalizationForwardTrainingEx(
            h, _PERSISTENT, ops, C.byref(_ONE), C.byref(_ZERO), xd, _vp(x.data_ptr()), zd, zp, xd, _vp(y.data_ptr()), bd,
            _vp(weight.data_ptr()), _vp(bias.data_ptr()), C.c_double(momentum), _vp(rm), _vp(rv), C.c_double(eps),
            _vp(save_mean.data_ptr()), _vp(save_invstd.data_ptr()), act, _vp(ws.data_ptr()), C.c_size_t(wf), _vp(reserve.data_ptr()),
            C.c_size_t(rs)), "cudnnBatchNormalizationForwardTrainingEx")
        if ops == _OPS_BN:
            ctx.save_for_backward(x, weight, bias, save_mean, save_invstd, reserve)
        else:
            ctx.save_for_backward(x, weight, bias, save_mean, save_invstd, reserve, y)
        ctx.eps, ctx.ops = eps, ops
        return y

    @staticmethod
    def backward(ctx, dy):
        ops = ctx.ops
        if ops == _OPS_BN:
            x, weight, bias, save_mean, save_invstd, reserve = ctx.saved_tensors
            y = None
        else:
            x, weight, bias, save_mean, save_invstd, reserve, y = ctx.saved_tensors
        dev = x.device
        lib = _Cudnn.lib
        h = _Cudnn.handle(dev)
        xd, bd, wf, wb, rs, act = _Cudnn.plan(h, x.shape, x.dtype, dev, ops)
        if dy.dtype != x.dtype or not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        dz = torch.empty_like(x) if ops == _OPS_BN_ADD_ACT else None
        dw = torch.empty_like(weight)
        db = torch.empty_like(weight)
        ws = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
        yd, yp = (xd, _vp(y.data_ptr())) if y is not None else (None, None)
        zd, zp = (xd, _vp(dz.data_ptr())) if dz is not None else (None, None)
        _Cudnn.check(lib.cudnnBatchNormalizationBackwardEx(
            h, _PERSISTENT, ops, C.byref(_ONE), C.byref(_ZERO), C.byref(_ONE), C.byref(_ZERO), xd, _vp(x.data_ptr()), yd, yp, xd,
            _vp(dy.data_ptr()), zd, zp, xd, _vp(dx.data_ptr()), bd, _vp(weight.data_ptr()), _vp(bias.data_ptr()) if y is not None else None,
            _vp(dw.data_ptr()), _vp(db.data_ptr()), C.c_double(ctx.eps), _vp(save_mean.data_ptr()), _vp(save_invstd.data_ptr()), act,
            _vp(ws.data_ptr()), C.c_size_t(wb), _vp(reserve.data_ptr()), C.c_size_t(rs)), "cudnnBatchNormalizationBackwardEx")
        return dx, dz, dw, db, None, None, None, None, None


class _CudnnMaxPoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, pad, stride):
        dev = x.device
        n, c, hh, ww = x.shape
        ho, wo = (hh + 2 * pad[0] - k[0]) // stride[0] + 1, (ww + 2 * pad[1] - k[1]) // stride[1] + 1
        y = torch.empty((n, c, ho, wo), dtype=x.dtype, device=dev).contiguous(memory_format=torch.channels_last)
        h = _Cudnn.handle(dev)
        pd, xd, yd = _Cudnn.pool_plan(x.shape, y.shape, x.dtype, dev, k, pad, stride)
        _Cudnn.check(_Cudnn.lib.cudnnPoolingForward(h, pd, C.byref(_ONE), xd, _vp(x.data_ptr()), C.byref(_ZERO), yd, _vp(y.data_ptr())),
                     "cudnnPoolingForward")
        ctx.save_for_backward(x, y)
        ctx.cfg = (k, pad, stride)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dev = x.device
        if dy.dtype != x.dtype or not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.to(x.dtype).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        h = _Cudnn.handle(dev)
        pd, xd, yd = _Cudnn.pool_plan(x.shape, y.shape, x.dtype, dev, *ctx.cfg)
        _Cudnn.check(_Cudnn.lib.cudnnPoolingBackward(h, pd, C.byref(_ONE), yd, _vp(y.data_ptr()), yd, _vp(dy.data_ptr()), xd, _vp(x.data_ptr()),
                                                     C.byref(_ZERO), xd, _vp(dx.data_ptr())), "cudnnPoolingBackward")
        return dx, None, None, None


def _half_nhwc(x):
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float16) and x.numel() > 0
            and x.is_contiguous(memory_format=torch.channels_last))


class CudnnBatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d whose training-mode forward / backward of channels_last bf16 / fp16 inputs runs on cuDNN's NHWC
    persistent batch-norm.  Parameters, buffers and state_dict keys are those of nn.BatchNorm2d.
    forward(x, z=None, relu=None) optionally fuses the residual add and the ReLU that follow the layer in a residual block:
    relu(bn(x) + z); `fused_relu` makes relu=True the default (a bn whose nn.ReLU successor was replaced by nn.Identity)."""

    fused_relu = False

    def forward(self, x, z=None, relu=None):
        relu = self.fused_relu if relu is None else relu
        use = (self.training and _half_nhwc(x) and self.affine and self.weight.dtype == torch.float32 and x.shape[1] % 4 == 0
               and (z is None or (relu and z.shape == x.shape and z.dtype == x.dtype and z.is_cuda)) and _Cudnn.load() is not None)
        if not use:
            y = super().forward(x)
            if z is not None:
                y = y + z
            return torch.relu(y) if relu else y
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
        self.__dict__["_stats_gen"] = self.__dict__.get("_stats_gen", 0) + 1      # invalidates the inference fold (see _fold)
        ops = _OPS_BN if not relu else (_OPS_BN_ACT if z is None else _OPS_BN_ADD_ACT)
        if z is not None and not z.is_contiguous(memory_format=torch.channels_last):
            z = z.contiguous(memory_format=torch.channels_last)
        return _CudnnBNFunction.apply(x, z, self.weight, self.bias, rm, rv, factor, self.eps, ops)


class CudnnMaxPool2d(nn.MaxPool2d):
    """nn.MaxPool2d whose channels_last bf16 / fp16 forward and backward run on cudnnPooling{Forward,Backward} (deterministic
    max); anything else (fp32, NCHW, dilation, ceil_mode, return_indices) takes the stock path."""

    def forward(self, x):
        two = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        if (_half_nhwc(x) and two(self.dilation) == (1, 1) and not self.ceil_mode and not self.return_indices and _Cudnn.load() is not None):
            return _CudnnMaxPoolFunction.apply(x, two(self.kernel_size), two(self.padding), two(self.stride if self.stride is not None else self.kernel_size))
        return super().forward(x)


# ---- inference: batch-norm folded into the convolution, bias / residual add / ReLU in cuDNN's fused convolution ------------
def _foldable(conv, bn):
    return (isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d) and bn.track_running_stats and bn.running_mean is not None
            and bn.affine and conv.groups == 1 and tuple(conv.dilation) == (1, 1) and conv.padding_mode == "zeros"
            and not isinstance(conv.padding, str))


def _fold(conv, bn, dtype):
    """(weight, bias) of the convolution with the eval-mode batch-norm folded in: w * g/sqrt(var+eps), beta + (b - mean) * g/sqrt(var+eps);
    computed in fp32, stored once in `dtype` (channels_last) and rebuilt when any source tensor changes (version counters)."""
    src = (conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    # (_stats_gen: cuDNN updates the running statistics through raw pointers, which torch's version counters do not see)
    key = tuple((t.data_ptr(), t._version) if t is not None else None for t in src) + (dtype, bn.__dict__.get("_stats_gen", 0))
    cached = bn.__dict__.get("_folded")
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    with torch.no_grad():
        sc = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
        w = (conv.weight.float() * sc.view(-1, 1, 1, 1)).to(dtype).contiguous(memory_format=torch.channels_last)
        b0 = conv.bias.float() if conv.bias is not None else 0.0
        b = (bn.bias.float() + (b0 - bn.running_mean.float()) * sc).to(dtype)
    bn.__dict__["_folded"] = (key, w, b)
    return w, b


def _conv_relu(x, conv, bn, z=None, extra_bias=None):
    w, b = _fold(conv, bn, x.dtype)
    if extra_bias is not None:
        b = b + extra_bias
    if z is None:
        return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    return torch.cudnn_convolution_add_relu(x, w, z, 1.0, b, conv.stride, conv.padding, conv.dilation, conv.groups)


def _identity_eval(self, x):
    """residual branch in inference: x, or the downsample convolution with its folded batch-norm's bias deferred to the block's
    last fused call (both biases are added before the ReLU)"""
    ds = self.downsample
    if ds is None:
        return x, None
    w, b = _fold(ds[0], ds[1], x.dtype)
    return torch.nn.functional.conv2d(x, w, None, ds[0].stride, ds[0].padding, ds[0].dilation, ds[0].groups), b


def _block_can_fold(self, x, pairs):
    if self.training or not _half_nhwc(x) or torch.is_grad_enabled():
        return False
    ds = self.downsample
    if ds is not None and not (isinstance(ds, nn.Sequential) and len(ds) == 2 and _foldable(ds[0], ds[1])):
        return False
    return all(_foldable(getattr(self, c), getattr(self, b)) for c, b in pairs)


def _bottleneck_forward(self, x):
    """torchvision Bottleneck.forward with the ReLUs and the residual add folded into the batch-norm calls (training) or, with
    the batch-norms folded into the weights, into cuDNN's fused convolutions (inference under no_grad)"""
    if _block_can_fold(self, x, (("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3"))):
        out = _conv_relu(x, self.conv1, self.bn1)
        out = _conv_relu(out, self.conv2, self.bn2)
        identity, bias = _identity_eval(self, x)
        return _conv_relu(out, self.conv3, self.bn3, z=identity, extra_bias=bias)
    out = self.bn1(self.conv1(x), relu=True)
    out = self.bn2(self.conv2(out), relu=True)
    identity = x if self.downsample is None else self.downsample(x)
    return self.bn3(self.conv3(out), z=identity, relu=True)


def _basicblock_forward(self, x):
    if _block_can_fold(self, x, (("conv1", "bn1"), ("conv2", "bn2"))):
        out = _conv_relu(x, self.conv1, self.bn1)
        identity, bias = _identity_eval(self, x)
        return _conv_relu(out, self.conv2, self.bn2, z=identity, extra_bias=bias)
    out = self.bn1(self.conv1(x), relu=True)
    identity = x if self.downsample is None else self.downsample(x)
    return self.bn2(self.conv2(out), z=identity, relu=True)


_FUSED_CLASSES = {}


def _fused_class(base, fwd):
    """subclass of a torchvision block with the fused forward, registered in this module so that modules holding it can be
    pickled (torch.save(model)) and deep-copied"""
    if base not in _FUSED_CLASSES:
        cls = type("Fused" + base.__name__, (base,), {"forward": fwd, "__module__": __name__})
        globals()[cls.__name__] = cls
        _FUSED_CLASSES[base] = cls
    return _FUSED_CLASSES[base]


try:  # created at import when torchvision is present, so that a pickled module can be loaded in a fresh process
    from torchvision.models.resnet import BasicBlock as _BasicBlock, Bottleneck as _Bottleneck
except Exception:  # pragma: no cover
    _BasicBlock = _Bottleneck = None


def fuse_residual_blocks(module):
    """After convert_batchnorm: switches every torchvision Bottleneck / BasicBlock under `module` to the fused forward above,
    and a (Conv2d, CudnnBatchNorm2d, ReLU, MaxPool2d) run of an nn.Sequential (the ResNet stem) to bn+relu in one call and
    cuDNN max-pooling.  Module names, parameters and buffers do not change.  Returns the number of blocks switched."""
    from torchvision.models.resnet import BasicBlock, Bottleneck
    n = 0
    for m in module.modules():
        for base, fwd, bns in ((Bottleneck, _bottleneck_forward, ("bn1", "bn2", "bn3")), (BasicBlock, _basicblock_forward, ("bn1", "bn2"))):
            if type(m) is base and type(m.relu) is nn.ReLU and all(type(getattr(m, b)) is CudnnBatchNorm2d for b in bns):
                m.__class__ = _fused_class(base, fwd)
                n += 1
        if isinstance(m, nn.Sequential):
            kids = list(m.named_children())
            for i in range(len(kids) - 2):
                (_, a), (nb, b), (nr, r) = kids[i], kids[i + 1], kids[i + 2]
                if isinstance(a, nn.Conv2d) and type(b) is CudnnBatchNorm2d and type(r) is nn.ReLU:
                    b.fused_relu = True
                    setattr(m, nr, nn.Identity())
                    if i + 3 < len(kids) and type(kids[i + 3][1]) is nn.MaxPool2d:
                        p = kids[i + 3][1]
                        setattr(m, kids[i + 3][0], CudnnMaxPool2d(p.kernel_size, p.stride, p.padding, p.dilation, p.return_indices, p.ceil_mode))
                    n += 1
    return n


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


if _Bottleneck is not None:
    _fused_class(_Bottleneck, _bottleneck_forward)
    _fused_class(_BasicBlock, _basicblock_forward)
