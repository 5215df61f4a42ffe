"""Does cuDNN's NHWC (persistent) batch-norm accept bf16 on this box, and how fast is it next to ATen's channels_last kernels?
Exploration for the encoder boundary (the trunk stays on cuDNN through PyTorch; torch's dispatcher skips cuDNN BN for bf16)."""
import ctypes as C, glob, os, site, sys
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
NHWC, FLOAT, HALF, BF16 = 1, 0, 2, 9
SPATIAL, PERSIST = 1, 2
def tdesc(fmt, dt, n, c, hh, ww):
    d = vp()
    chk(lib.cudnnCreateTensorDescriptor(C.byref(d)), "ctd")
    chk(lib.cudnnSetTensor4dDescriptor(d, fmt, dt, n, c, hh, ww), "set4d")
    return d
def run(dtype, cd, mode, N=128, Cc=256, Hh=56, Ww=56, iters=20):
    x = torch.randn(N, Cc, Hh, Ww, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    w = torch.rand(Cc, device="cuda") + 0.5
    b = torch.randn(Cc, device="cuda")
    xd = tdesc(NHWC, cd, N, Cc, Hh, Ww)
    bd = vp(); chk(lib.cudnnCreateTensorDescriptor(C.byref(bd)), "ctd")
    chk(lib.cudnnDeriveBNTensorDescriptor(bd, xd, mode), "derive")
    ws, rs, wsb = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    chk(lib.cudnnGetBatchNormalizationForwardTrainingExWorkspaceSize(h, mode, 0, xd, None, xd, bd, None, C.byref(ws)), "wsz")
    chk(lib.cudnnGetBatchNormalizationTrainingExReserveSpaceSize(h, mode, 0, None, xd, C.byref(rs)), "rsz")
    chk(lib.cudnnGetBatchNormalizationBackwardExWorkspaceSize(h, mode, 0, xd, None, xd, None, xd, bd, None, C.byref(wsb)), "wszb")
    wsp = torch.empty(max(ws.value, wsb.value, 16), dtype=torch.uint8, device="cuda")
    rsp = torch.empty(max(rs.value, 16), dtype=torch.uint8, device="cuda")
    y = torch.empty_like(x); rm = torch.zeros(Cc, device="cuda"); rv = torch.ones(Cc, device="cuda")
    sm = torch.empty(Cc, device="cuda"); si = torch.empty(Cc, device="cuda")
    one, zero = C.c_float(1.0), C.c_float(0.0)
    def fwd():
        chk(lib.cudnnBatchNormalizationForwardTrainingEx(h, mode, 0, C.byref(one), C.byref(zero), xd, vp(x.data_ptr()), None, None, xd, vp(y.data_ptr()),
            bd, vp(w.data_ptr()), vp(b.data_ptr()), C.c_double(0.1), vp(rm.data_ptr()), vp(rv.data_ptr()), C.c_double(1e-5),
            vp(sm.data_ptr()), vp(si.data_ptr()), None, vp(wsp.data_ptr()), C.c_size_t(ws.value), vp(rsp.data_ptr()), C.c_size_t(rs.value)), "fwd")
    dy = torch.randn_like(x); dx = torch.empty_like(x); dw = torch.empty(Cc, device="cuda"); db = torch.empty(Cc, device="cuda")
    def bwd():
        chk(lib.cudnnBatchNormalizationBackwardEx(h, mode, 0, C.byref(one), C.byref(zero), C.byref(one), C.byref(zero), xd, vp(x.data_ptr()), None, None,
            xd, vp(dy.data_ptr()), None, None, xd, vp(dx.data_ptr()), bd, vp(w.data_ptr()), None, vp(dw.data_ptr()), vp(db.data_ptr()), C.c_double(1e-5),
            vp(sm.data_ptr()), vp(si.data_ptr()), None, vp(wsp.data_ptr()), C.c_size_t(wsb.value), vp(rsp.data_ptr()), C.c_size_t(rs.value)), "bwd")
    fwd(); bwd(); torch.cuda.synchronize()
    xr = x.float().requires_grad_(True)
    yr = torch.nn.functional.batch_norm(xr, None, None, w, b, True, 0.1, 1e-5)
    yr.backward(dy.float())
    print("  fwd err %.3e  dx err %.3e  dw err %.3e" % (float((y.float() - yr).abs().max()), float((dx.float() - xr.grad).abs().max() / xr.grad.abs().max()),
          float((dw - (dy.float() * ((x.float() - x.float().mean((0, 2, 3), keepdim=True)) * si.view(1, -1, 1, 1))).sum((0, 2, 3))).abs().max() / dw.abs().max())))
    def t(fn):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3
    tf, tb = t(fwd), t(bwd)
    bn = torch.nn.BatchNorm2d(Cc).cuda().to(memory_format=torch.channels_last)
    xx = x.clone().requires_grad_(True)
    def af():
        return bn(xx)
    yy = af()
    def ab():
        yy.backward(dy, retain_graph=True)
    print("  cudnn fwd %.1f us  bwd %.1f us   |  torch fwd %.1f us  bwd %.1f us   (ws %d, reserve %d)" % (tf, tb, t(af), t(ab), ws.value, rs.value))
for name, dt, cd in (("fp16", torch.float16, HALF), ("bf16", torch.bfloat16, BF16)):
    for mname, mode in (("SPATIAL", SPATIAL), ("PERSISTENT", PERSIST)):
        print(name, mname, flush=True)
        try:
            run(dt, cd, mode)
            run(dt, cd, mode, N=256, Cc=2048, Hh=7, Ww=7)
        except Exception as e:
            print("  FAILED:", e)
