"""Which batch-norm kernels does the torchvision trunk run under bf16 / fp16 autocast + channels_last, and how long does a
ResNet-50 forward+backward take under each setting?  (exploration for the encoder boundary; not product code)"""
import sys, time
import torch, torchvision

def run(tag, dtype, cl, B=128, steps=4, prof=False, cudnn_bn=True):
    torch.manual_seed(0)
    m = torchvision.models.resnet50(weights=None).cuda()
    m = torch.nn.Sequential(*list(m.children())[:-2])
    if cl:
        m = m.to(memory_format=torch.channels_last)
    m.train()
    x = torch.rand(B, 3, 224, 224, device="cuda")
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    def step():
        with torch.autocast("cuda", dtype=dtype):
            y = m(x)
        y.float().mean().backward()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("%-40s %8.2f ms/step" % (tag, e0.elapsed_time(e1) / steps), flush=True)
    if prof:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as p:
            step()
            torch.cuda.synchronize()
        rows = sorted(p.key_averages(), key=lambda r: -r.device_time_total)[:12]
        for r in rows:
            print("    %8.2f ms  n=%4d  %s" % (r.device_time_total / 1e3, r.count, r.key[:110]))

print(torch.__version__, torch.backends.cudnn.version(), torch.cuda.get_device_name())
run("bf16 channels_last", torch.bfloat16, True, prof=True)
run("fp16 channels_last", torch.float16, True, prof=True)
run("bf16 NCHW", torch.bfloat16, False, prof=True)
torch.backends.cudnn.benchmark = True
run("bf16 channels_last cudnn.benchmark", torch.bfloat16, True)
run("fp16 channels_last cudnn.benchmark", torch.float16, True)
