"""Encoder-boundary probe: does cuDNN's NHWC persistent batch-norm accept the fused ops (BN+ReLU, BN+Add+ReLU) in bf16 on this box,
are the results those of the unfused PyTorch composite, and how fast are they?  Also cuDNN max-pooling (NHWC bf16) next to ATen's
max_pool2d channels_last kernels (3x3/s2 on [256,64,112,112], the ResNet stem)."""
import ctypes as C, glob, os, site
import torch
lib = None
for sp in site.getsitepackages():
    for f in glob.glob(os.path.join(sp, "nvidia", "cudnn", "lib", "libcudnn.so.9")):
        lib = C.CDLL(f, mode=C.RTLD_GLOBAL)
if lib is None:
    lib = C.CDLL("libcudnn.so.9")
vp = C.c_void_p
lib.cudnnGetErrorString.restype = C.c_char_p
def chk(rc, what):
    if rc != 0:
        raise RuntimeError("%s -> %d %s" % (what, rc, lib.cudnnGetErrorString(rc)))
h = vp()
chk(lib.cudnnCreate(C.byref(h)), "create")
chk(lib.cudnnSetStream(h, vp(torch.cuda.current_stream().cuda_stream)), "stream")
NHWC, HALF, BF16, PERSIST = 1, 2, 9, 2
one, zero = C.c_float(1.0), C.c_float(0.0)
def tdesc(dt, n, c, hh, ww):
    d = vp()
    chk(lib.cudnnCreateTensorDescriptor(C.byref(d)), "ctd")
    chk(lib.cudnnSetTensor4dDescriptor(d, NHWC, dt, n, c, hh, ww), "set4d")
    return d
def t(fn, iters=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
act = vp()
chk(lib.cudnnCreateActivationDescriptor(C.byref(act)), "cad")
chk(lib.cudnnSetActivationDescriptor(act, 1, 0, C.c_double(0.0)), "sad")      # RELU, NOT_PROPAGATE_NAN

def run_bn(dtype, cd, ops, N, Cc, Hh, Ww):
    x = torch.randn(N, Cc, Hh, Ww, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    z = torch.randn(N, Cc, Hh, Ww, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    w = torch.rand(Cc, device="cuda") + 0.5
    b = torch.randn(Cc, device="cuda") * 0.5
    xd = tdesc(cd, N, Cc, Hh, Ww)
    zd = xd if ops == 2 else None
    bd = vp(); chk(lib.cudnnCreateTensorDescriptor(C.byref(bd)), "ctd")
    chk(lib.cudnnDeriveBNTensorDescriptor(bd, xd, PERSIST), "derive")
    ws, rs, wsb = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    chk(lib.cudnnGetBatchNormalizationForwardTrainingExWorkspaceSize(h, PERSIST, ops, xd, zd, xd, bd, act, C.byref(ws)), "wsz")
    chk(lib.cudnnGetBatchNormalizationTrainingExReserveSpaceSize(h, PERSIST, ops, act, xd, C.byref(rs)), "rsz")
    chk(lib.cudnnGetBatchNormalizationBackwardExWorkspaceSize(h, PERSIST, ops, xd, xd, xd, zd, xd, bd, act, C.byref(wsb)), "wszb")
    wsp = torch.empty(max(ws.value, wsb.value, 16), dtype=torch.uint8, device="cuda")
    rsp = torch.empty(max(rs.value, 16), dtype=torch.uint8, device="cuda")
    y = torch.empty_like(x); rm = torch.zeros(Cc, device="cuda"); rv = torch.ones(Cc, device="cuda")
    sm = torch.empty(Cc, device="cuda"); si = torch.empty(Cc, device="cuda")
    zp = vp(z.data_ptr()) if ops == 2 else None
    def fwd():
        chk(lib.cudnnBatchNormalizationForwardTrainingEx(h, PERSIST, ops, C.byref(one), C.byref(zero), xd, vp(x.data_ptr()), zd, zp, xd, vp(y.data_ptr()),
            bd, vp(w.data_ptr()), vp(b.data_ptr()), C.c_double(0.1), vp(rm.data_ptr()), vp(rv.data_ptr()), C.c_double(1e-5),
            vp(sm.data_ptr()), vp(si.data_ptr()), act, vp(wsp.data_ptr()), C.c_size_t(ws.value), vp(rsp.data_ptr()), C.c_size_t(rs.value)), "fwd")
    dy = torch.randn_like(x); dx = torch.empty_like(x); dz = torch.empty_like(x); dw = torch.empty(Cc, device="cuda"); db = torch.empty(Cc, device="cuda")
    dzp = vp(dz.data_ptr()) if ops == 2 else None
    def bwd():
        chk(lib.cudnnBatchNormalizationBackwardEx(h, PERSIST, ops, C.byref(one), C.byref(zero), C.byref(one), C.byref(zero), xd, vp(x.data_ptr()), xd, vp(y.data_ptr()),
            xd, vp(dy.data_ptr()), zd, dzp, xd, vp(dx.data_ptr()), bd, vp(w.data_ptr()), vp(b.data_ptr()), vp(dw.data_ptr()), vp(db.data_ptr()), C.c_double(1e-5),
            vp(sm.data_ptr()), vp(si.data_ptr()), act, vp(wsp.data_ptr()), C.c_size_t(wsb.value), vp(rsp.data_ptr()), C.c_size_t(rs.value)), "bwd")
    fwd(); bwd(); torch.cuda.synchronize()
    xr = x.float().requires_grad_(True); zr = z.float().requires_grad_(True)
    yr = torch.nn.functional.batch_norm(xr, None, None, w, b, True, 0.1, 1e-5)
    if ops == 2:
        yr = yr + zr
    yr = torch.relu(yr)
    yr.backward(dy.float())
    msg = "  ops=%d [%d,%d,%d,%d]: y err %.3e  dx relerr %.3e" % (ops, N, Cc, Hh, Ww, float((y.float() - yr).abs().max()),
                                                                float((dx.float() - xr.grad).abs().max() / xr.grad.abs().max()))
    if ops == 2:
        msg += "  dz relerr %.3e" % float((dz.float() - zr.grad).abs().max() / zr.grad.abs().max())
        own = ((dz != 0) == (y > 0)) | (dy == 0)                 # cuDNN's mask against its OWN output
        flips = int(((y > 0) != (yr > 0)).sum())                  # ... against the fp32 composite (sign of a value that rounds near 0)
        msg += "  mask self-consistent: %s, sign flips vs fp32 composite: %d of %d (|pre-activation| there <= %.2e)" % (
            bool(own.all()), flips, y.numel(), float((torch.nn.functional.batch_norm(x.float(), None, None, w, b, True, 0.1, 1e-5) + z.float())[(y > 0) != (yr > 0)].abs().max()) if flips else 0.0)
    print(msg)
    bn = torch.nn.BatchNorm2d(Cc).cuda()
    xx = x.clone().requires_grad_(True); zz = z.clone().requires_grad_(True)
    def af():
        o = torch.nn.functional.batch_norm(xx, None, None, bn.weight, bn.bias, True, 0.1, 1e-5)
        if ops == 2:
            o = o + zz
        return torch.relu(o)
    yy = af()
    def ab():
        yy.backward(dy, retain_graph=True)
    print("     cudnn fused fwd %.1f us  bwd %.1f us   |  torch (aten bn + add + relu) fwd %.1f us  bwd %.1f us   (ws %d/%d, reserve %d)" % (t(fwd), t(bwd), t(af), t(ab), ws.value, wsb.value, rs.value))

import sys
for name, dt, cd in (("bf16", torch.bfloat16, BF16),):
    for ops in (2,):
        for shape in ((256, 256, 56, 56), (256, 256, 56, 56), (128, 256, 56, 56), (256, 512, 28, 28)):
            print(name, "bnOps", ops, flush=True)
            try:
                run_bn(dt, cd, ops, *shape)
            except Exception as e:
                print("  FAILED:", e)

# ---- max pooling ----
def run_pool(dtype, cd, mode, N=256, Cc=64, Hh=112, Ww=112):
    x = torch.randn(N, Cc, Hh, Ww, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    Ho, Wo = (Hh + 2 - 3) // 2 + 1, (Ww + 2 - 3) // 2 + 1
    y = torch.empty(N, Cc, Ho, Wo, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    xd, yd = tdesc(cd, N, Cc, Hh, Ww), tdesc(cd, N, Cc, Ho, Wo)
    pd = vp(); chk(lib.cudnnCreatePoolingDescriptor(C.byref(pd)), "cpd")
    chk(lib.cudnnSetPooling2dDescriptor(pd, mode, 0, 3, 3, 1, 1, 2, 2), "spd")
    def fwd():
        chk(lib.cudnnPoolingForward(h, pd, C.byref(one), xd, vp(x.data_ptr()), C.byref(zero), yd, vp(y.data_ptr())), "pool fwd")
    dy = torch.randn_like(y); dx = torch.empty_like(x)
    def bwd():
        chk(lib.cudnnPoolingBackward(h, pd, C.byref(one), yd, vp(y.data_ptr()), yd, vp(dy.data_ptr()), xd, vp(x.data_ptr()), C.byref(zero), xd, vp(dx.data_ptr())), "pool bwd")
    fwd(); bwd(); torch.cuda.synchronize()
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.max_pool2d(xr, 3, 2, 1)
    yr.backward(dy)
    print("  pool mode %d: y equal %s, dx max diff %.3e (ties may be attributed differently)" % (mode, bool(torch.equal(y, yr)), float((dx.float() - xr.grad.float()).abs().max())))
    yy = torch.nn.functional.max_pool2d(xr, 3, 2, 1)
    print("     cudnn fwd %.1f us  bwd %.1f us  |  aten fwd %.1f us  bwd %.1f us" % (t(fwd), t(bwd), t(lambda: torch.nn.functional.max_pool2d(xr, 3, 2, 1)),
                                                                                      t(lambda: yy.backward(dy, retain_graph=True))))
for mode in ():
    try:
        run_pool(torch.bfloat16, BF16, mode)
    except Exception as e:
        print("  pool FAILED:", e)
