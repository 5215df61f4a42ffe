"""GPU: the encoder-boundary cuDNN modules (sat_b200/cudnn_bn.py).  The trunk is third-party torchvision code that stays on
cuDNN / ATen; these tests pin that swapping nn.BatchNorm2d / ReLU / residual add / MaxPool2d for the cuDNN NHWC calls does not
change what the trunk computes: same state_dict, same outputs and gradients as the stock modules to bf16 rounding, identical
running statistics, and the stock path in eval mode / fp32."""
import copy
import warnings

import pytest
import torch
from torch import nn

warnings.filterwarnings("ignore")
pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _trunks(arch):
    from torchvision import models
    from sat_b200.cudnn_bn import convert_batchnorm, fuse_residual_blocks
    torch.manual_seed(0)
    m = models.__dict__[arch](weights=None)
    ref = nn.Sequential(*list(m.children())[:-2]).cuda().to(memory_format=torch.channels_last)
    fused = copy.deepcopy(ref)
    assert convert_batchnorm(fused) > 0
    assert fuse_residual_blocks(fused) > 0
    return ref, copy.deepcopy(ref), fused


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_fused_trunk_is_as_close_to_fp32_as_the_stock_bf16_trunk(arch):
    """A randomly initialised deep ResNet amplifies bf16 rounding (stock bf16 autocast against fp32: output cosine 0.999 for
    resnet18 but 0.84 for resnet50, gradient cosines 0.93 / 0.15), so the fused trunk cannot be compared element-wise with the
    stock bf16 trunk.  The yardstick is the fp32 trunk: the fused bf16 trunk must sit as close to it as the stock bf16 trunk
    does (output cosine, median gradient cosine), with identical state_dict keys and batch-norm bookkeeping.  The layer-level
    test below pins the fused calls themselves tightly."""
    stock, fp32, fused = _trunks(arch)
    assert list(stock.state_dict().keys()) == list(fused.state_dict().keys())
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(16, 3, 128, 128, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    res = {}
    for name, net in (("stock", stock), ("fp32", fp32), ("fused", fused)):
        net.train()
        if name == "fp32":
            y = net(x.clone())
        else:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = net(x.clone())
        dy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)).to(y.dtype)
        y.backward(dy)
        res[name] = (y.detach().flatten().double(), {k: p.grad.flatten().double() for k, p in net.named_parameters()})
        for k, p in net.named_parameters():
            assert torch.isfinite(p.grad).all(), (name, k)
    cos = lambda a, b: float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))
    med = lambda v: sorted(v)[len(v) // 2]
    y0, G0 = res["fp32"]
    out_stock, out_fused = cos(res["stock"][0], y0), cos(res["fused"][0], y0)
    g_stock = med([cos(res["stock"][1][k], G0[k]) for k in G0])
    g_fused = med([cos(res["fused"][1][k], G0[k]) for k in G0])
    print("%s: output cosine vs fp32 stock %.4f fused %.4f; median gradient cosine stock %.4f fused %.4f" % (arch, out_stock, out_fused, g_stock, g_fused))
    assert out_fused > out_stock - 0.01
    assert g_fused > g_stock - 0.03
    sd_s, sd_f = stock.state_dict(), fused.state_dict()
    for k in sd_s:
        if "num_batches_tracked" in k:
            assert int(sd_f[k]) == int(sd_s[k]) == 1
    # first layers (before the rounding differences have been amplified): running statistics agree tightly
    first_bn = [k for k in sd_s if "running_mean" in k][0]
    assert relerr(sd_f[first_bn], sd_s[first_bn]) < 1e-2


def test_fused_ops_match_composite_exactly_where_they_should():
    """one layer: relu(bn(x) + z) against the stock composite in fp32 math on the same bf16 inputs"""
    from sat_b200.cudnn_bn import CudnnBatchNorm2d
    torch.manual_seed(0)
    bn = CudnnBatchNorm2d(64).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.5)
    x = torch.randn(32, 64, 14, 14, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    z = torch.randn(32, 64, 14, 14, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    dy = torch.randn(32, 64, 14, 14, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    for use_z, relu in ((False, False), (False, True), (True, True)):
        for t in (x, z, bn.weight, bn.bias):
            t.grad = None
        y = bn(x, z if use_z else None, relu=relu)
        y.backward(dy)
        got = (y.float(), x.grad.float(), z.grad.float() if use_z else None, bn.weight.grad.clone(), bn.bias.grad.clone())
        xr, zr = x.detach().float().requires_grad_(True), z.detach().float().requires_grad_(True)
        w, b = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
        yr = torch.nn.functional.batch_norm(xr, None, None, w, b, True, 0.1, bn.eps)
        if use_z:
            yr = yr + zr
        if relu:
            yr = torch.relu(yr)
        yr.backward(dy.float())
        assert relerr(got[0], yr) < 1e-2
        assert relerr(got[1], xr.grad) < 2e-2
        if use_z:
            assert relerr(got[2], zr.grad) < 1e-2
        assert relerr(got[3], w.grad) < 2e-2 and relerr(got[4], b.grad) < 2e-2
    # eval mode and fp32 take the stock path (bit-identical to nn.BatchNorm2d followed by add / relu)
    bn.eval()
    ref = nn.BatchNorm2d(64).cuda().eval()
    ref.load_state_dict(bn.state_dict())
    with torch.no_grad():
        assert torch.equal(bn(x, z, relu=True), torch.relu(ref(x) + z))
        assert torch.equal(bn(x.float()), ref(x.float()))


def test_cudnn_maxpool_matches_stock():
    from sat_b200.cudnn_bn import CudnnMaxPool2d
    p, ref = CudnnMaxPool2d(3, 2, 1), nn.MaxPool2d(3, 2, 1)
    x = torch.randn(8, 64, 56, 56, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = p(xa), ref(xb)
    assert torch.equal(ya, yb)
    dy = torch.randn_like(ya)
    ya.backward(dy)
    yb.backward(dy)
    assert relerr(xa.grad.float(), xb.grad.float()) < 1e-2       # ties inside a window may be attributed to a different element
    assert torch.equal(p(x.float()), ref(x.float()))              # fp32: stock path


def test_sat_module_uses_the_fused_trunk_and_keeps_the_reference_state_dict():
    from oracle import ref_harness as rh
    from sat_b200.cudnn_bn import CudnnBatchNorm2d, CudnnMaxPool2d
    from sat_b200.model import SAT
    hp = rh.default_hparams(encoder_arch="resnet18", encoder_dim=64, attention_dim=32, embed_dim=32, decoder_dim=64, vocab_size=128,
                            input_size=64, precision="bf16")
    torch.manual_seed(0)
    m = SAT(**hp)
    hp2 = dict(hp, cudnn_batchnorm=False)
    torch.manual_seed(0)
    m2 = SAT(**hp2)
    assert list(m.state_dict().keys()) == list(m2.state_dict().keys())
    assert any(type(x) is CudnnBatchNorm2d for x in m.encoder.modules()) and any(type(x) is CudnnMaxPool2d for x in m.encoder.modules())
    assert any(type(x).__name__ == "FusedBasicBlock" for x in m.encoder.modules())
    assert not any(type(x).__name__.startswith("Fused") for x in m2.encoder.modules())
    m, m2 = m.cuda(), m2.cuda()
    m.encoder.to(memory_format=torch.channels_last)
    m2.encoder.to(memory_format=torch.channels_last)
    img = torch.rand(8, 3, 64, 64, device="cuda")
    m.eval()
    with torch.no_grad():
        before = m.encode(img.clone())                                           # builds the folded inference weights once
    m.train(); m2.train()
    a, b = m.encode(img.clone()), m2.encode(img.clone())
    assert a.shape == b.shape and relerr(a.float(), b.float()) < 5e-2
    m2.load_state_dict(m.state_dict())                                           # same running statistics
    m.eval(); m2.eval()
    with torch.no_grad():
        a, b = m.encode(img.clone()), m2.encode(img.clone())                     # inference: folded batch-norm + fused cuDNN convolutions
    assert relerr(a.float(), b.float()) < 5e-2
    assert not torch.equal(a, before)                                            # the fold followed the new running statistics
    with torch.enable_grad():                                                    # eval mode WITH autograd: the stock (differentiable) path
        x = img.clone().requires_grad_(True)
        y = m.encoder(x.contiguous(memory_format=torch.channels_last))
    assert y.requires_grad


@pytest.mark.parametrize("shape", [(4, 64, 7, 7, 14, 14), (3, 72, 8, 8, 14, 14), (2, 32, 5, 10, 16, 12), (2, 16, 9, 9, 6, 7)])
def test_resize_layer_matches_upsample(shape):
    """encoder tail (readme.md:118-121): the library's one-pass NHWC bilinear resize against nn.Upsample(mode="bilinear",
    align_corners=False) -- fp32 to 1e-6 (forward and backward), bf16 = the fp32 interpolation of the stored bf16 values rounded
    once; the gather-form backward is bit-reproducible"""
    from sat_b200.encoder_tail import ResizeBilinearNHWC
    n, D, h, w, H2, W2 = shape
    layer, ref = ResizeBilinearNHWC((H2, W2), mode="bilinear", align_corners=False), nn.Upsample((H2, W2), mode="bilinear", align_corners=False)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, D, h, w, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(n, D, H2, W2, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = layer(xa), ref(xb)
    assert ya.is_contiguous(memory_format=torch.channels_last) and ya.shape == yb.shape
    assert relerr(ya, yb) < 1e-6
    ya.backward(dy)
    yb.backward(dy)
    assert relerr(xa.grad, xb.grad) < 1e-6
    # bf16
    xh = x.bfloat16()
    xc = xh.clone().requires_grad_(True)
    yh = layer(xc)
    assert yh.dtype == torch.bfloat16
    want = ref(xh.float())
    assert relerr(yh.float(), want) < 8e-3                      # one bf16 rounding of the fp32 interpolation
    yh.backward(dy.bfloat16())
    g1 = xc.grad.clone()
    xr = xh.float().requires_grad_(True)
    ref(xr).backward(dy.bfloat16().float())
    assert relerr(g1.float(), xr.grad) < 8e-3
    xc.grad = None
    layer(xc).backward(dy.bfloat16())
    assert torch.equal(xc.grad, g1)                              # deterministic
    # NCHW input: stock path
    assert torch.equal(layer(x.contiguous()), ref(x.contiguous()))


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_inference_trunk_with_folded_batchnorm(arch):
    """eval mode under no_grad: the residual blocks fold the batch-norm into the convolution and call cuDNN's fused
    convolution + bias (+ residual) + ReLU.  Yardstick as above: the fp32 eval trunk; the folded bf16 trunk must be as close to it
    as the stock bf16 eval trunk; the folded weights follow parameter updates (version counters)."""
    stock, fp32, fused = _trunks(arch)
    g = torch.Generator(device="cuda").manual_seed(3)
    with torch.no_grad():                                   # non-trivial running statistics and affine parameters, same in all three
        for net in (stock,):
            for mod in net.modules():
                if isinstance(mod, nn.BatchNorm2d):
                    mod.running_mean.normal_(0, 0.2, generator=g)
                    mod.running_var.uniform_(0.6, 1.6, generator=g)
                    mod.weight.uniform_(0.7, 1.3, generator=g)
                    mod.bias.normal_(0, 0.2, generator=g)
    fp32.load_state_dict(stock.state_dict())
    fused.load_state_dict(stock.state_dict())
    x = torch.rand(8, 3, 128, 128, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    cos = lambda a, b: float((a.flatten().double() @ b.flatten().double()) / (a.double().norm() * b.double().norm()).clamp_min(1e-30))
    for net in (stock, fp32, fused):
        net.eval()
    with torch.no_grad():
        y0 = fp32(x.clone())
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ys, yf = stock(x.clone()), fused(x.clone())
        assert yf.dtype == torch.bfloat16 and yf.shape == y0.shape
        cs, cf = cos(ys.float(), y0), cos(yf.float(), y0)
        print("%s inference: output cosine vs fp32 stock %.5f folded %.5f" % (arch, cs, cf))
        assert cf > cs - 5e-3 and cf > 0.99
        # the fold follows the parameters
        for mod in fused.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.weight.mul_(0.5)
        for mod in fp32.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.weight.mul_(0.5)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yf2 = fused(x.clone())
        assert cos(yf2.float(), fp32(x.clone())) > 0.99 and not torch.equal(yf2, yf)


@pytest.mark.parametrize("arch", ["resnext50_32x4d", "densenet121", "shufflenet_v2_x0_5", "mobilenet_v3_small", "squeezenet1_0"])
def test_other_reference_trunks_run_through_the_swapped_modules(arch):
    """the reference accepts every torchvision family in model.py:27-43; the cuDNN-routed modules must cope with their shapes
    (grouped convolutions, concatenated feature maps, ReLU6 / Hardswish activations that are NOT fused): train step and
    inference produce finite annotations / gradients of the reference's shape"""
    from oracle import ref_harness as rh
    from sat_b200.model import SAT
    hp = rh.default_hparams(encoder_arch=arch, encoder_dim=64, attention_dim=32, embed_dim=32, decoder_dim=64, vocab_size=128,
                            input_size=96, precision="bf16")
    torch.manual_seed(0)
    m = SAT(**hp).cuda()
    m.encoder.to(memory_format=torch.channels_last)
    img = torch.rand(4, 3, 96, 96, device="cuda")
    m.train()
    a = m.encode(img.clone())
    assert a.shape[:2] == (4, 64) and torch.isfinite(a.float()).all()
    a.float().square().mean().backward()
    grads = [p.grad for p in m.encoder.parameters() if p.requires_grad]
    assert grads and all(g is not None and torch.isfinite(g).all() for g in grads)
    m.eval()
    with torch.no_grad():
        b = m.encode(img.clone())
    assert b.shape == a.shape and torch.isfinite(b.float()).all()
