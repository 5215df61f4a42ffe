"""Where does the end-to-end train step lose time against the device-resident step?  (exploration)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sat_b200.model import SAT
c = bench.CFG["train"]; dev = torch.device("cuda")
torch.manual_seed(0)
model = SAT(**bench.hparams(c)).to(dev); model.encoder.to(memory_format=torch.channels_last); model.train()
opt = model.configure_optimizers()
B = 256
img_h, caps_h, lens_h = bench.synth_batch(B, 20, 6400, 1, pin=True)
img_d, caps_d, lens_d = img_h.to(dev), caps_h.to(dev), lens_h.to(dev)
def ev(): return torch.cuda.Event(enable_timing=True)
def timeit(fn, n=8, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0, e1 = ev(), ev(); e0.record()
    for _ in range(n): fn()
    e1.record(); t_issue = time.perf_counter() - t0; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * t_issue / n
print("H2D 154MB copy: %.2f ms (gpu), issue %.2f ms" % timeit(lambda: img_h.to(dev, non_blocking=True)))
def step_dev():
    opt.zero_grad(set_to_none=True)
    loss, aux = model.fused_loss((img_d.clone(), caps_d, lens_d)); loss.backward(); opt.step()
print("device step: %.2f ms gpu, %.2f ms cpu issue" % timeit(step_dev))
def step_h2d_serial():
    opt.zero_grad(set_to_none=True)
    loss, aux = model.fused_loss((img_h.to(dev, non_blocking=True), caps_d, lens_d)); loss.backward(); opt.step()
print("serial H2D + step: %.2f ms gpu, %.2f ms cpu issue" % timeit(step_h2d_serial))
cs = torch.cuda.Stream(); pend = {}
def prefetch():
    with torch.cuda.stream(cs):
        pend["b"] = img_h.to(dev, non_blocking=True); pend["e"] = torch.cuda.Event(); pend["e"].record(cs)
def step_pref():
    if "b" not in pend: prefetch()
    cur = torch.cuda.current_stream(); cur.wait_event(pend["e"]); img = pend.pop("b"); img.record_stream(cur); prefetch()
    opt.zero_grad(set_to_none=True)
    loss, aux = model.fused_loss((img, caps_d, lens_d)); loss.backward(); opt.step()
print("prefetched H2D + step: %.2f ms gpu, %.2f ms cpu issue" % timeit(step_pref))
buf = [torch.empty_like(img_d) for _ in range(2)]; st = {"i": 0}
def prefetch2():
    i = st["i"] & 1
    with torch.cuda.stream(cs):
        buf[i].copy_(img_h, non_blocking=True); pend["e"] = torch.cuda.Event(); pend["e"].record(cs)
    pend["b"] = buf[i]; st["i"] += 1
def step_pref2():
    if "b" not in pend: prefetch2()
    cur = torch.cuda.current_stream(); cur.wait_event(pend["e"]); img = pend.pop("b")
    done = torch.cuda.Event()
    opt.zero_grad(set_to_none=True)
    loss, aux = model.fused_loss((img, caps_d, lens_d))
    cs.wait_stream(cur)        # the other buffer is free again once this step's encoder forward has consumed... (conservative: whole fwd)
    prefetch2()
    loss.backward(); opt.step()
print("double-buffered H2D + step: %.2f ms gpu, %.2f ms cpu issue" % timeit(step_pref2))
