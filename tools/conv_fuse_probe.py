"""Probe (encoder boundary, inference): does PyTorch's cuDNN fused convolution (+bias +residual +ReLU) accept bf16 channels_last here,
does it match the unfused eval-mode conv -> batch-norm -> (add) -> relu with the batch-norm folded into the weights, and how fast is it?"""
import torch
import torch.nn.functional as F

def t(fn, iters=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

torch.manual_seed(0)
for (N, Ci, Co, H, k, s) in ((256, 64, 64, 56, 3, 1), (256, 256, 64, 56, 1, 1), (256, 64, 256, 56, 1, 1), (256, 512, 2048, 7, 1, 1), (256, 512, 512, 7, 3, 1)):
    conv = torch.nn.Conv2d(Ci, Co, k, s, k // 2, bias=False).cuda().to(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(Co).cuda().eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.3); bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5)
    x = torch.randn(N, Ci, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    z = torch.randn(N, Co, H // s, H // s, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    sc = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    wf = (conv.weight * sc.view(-1, 1, 1, 1)).bfloat16().contiguous(memory_format=torch.channels_last)
    bf = (bn.bias - bn.running_mean * sc).bfloat16()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        def stock(): return torch.relu(bn(conv(x)))
        def stock_add(): return torch.relu(bn(conv(x)) + z)
        ref, ref_add = stock(), stock_add()
        try:
            def fused(): return torch.cudnn_convolution_relu(x, wf, bf, (s, s), (k // 2, k // 2), (1, 1), 1)
            def fused_add(): return torch.cudnn_convolution_add_relu(x, wf, z, 1.0, bf, (s, s), (k // 2, k // 2), (1, 1), 1)
            y, ya = fused(), fused_add()
            e1 = float((y.float() - ref.float()).abs().max() / ref.float().abs().max())
            e2 = float((ya.float() - ref_add.float()).abs().max() / ref_add.float().abs().max())
            print("N%d %d->%d %dx%d k%d: relerr relu %.3e add_relu %.3e | stock %.1f / %.1f us  fused %.1f / %.1f us  (conv alone %.1f us) cl=%s" % (
                N, Ci, Co, H, H, k, e1, e2, t(stock), t(stock_add), t(fused), t(fused_add), t(lambda: conv(x)), y.is_contiguous(memory_format=torch.channels_last)))
        except Exception as e:
            print("FAILED", (N, Ci, Co, H, k), repr(e)[:300])
