"""Micro-benchmark of the fused attention-step forward kernel through the C ABI (sat_attention_step_fwd).
    SAT_ATT_MODE=0|1|2|4|8 python tools/att_bench.py [--batch 256] [--fp32] [--flush]
Times `--iters` launches with CUDA events; --flush rewrites a 256 MB buffer between launches (cold L2)."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from sat_b200 import _lib, decoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--L", type=int, default=196)
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--A", type=int, default=128)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--fp32", action="store_true")
ap.add_argument("--flush", action="store_true")
ap.add_argument("--reps", type=int, default=1, help="launches per timed region (amortises the host launch gap)")
args = ap.parse_args()
dt = torch.float32 if args.fp32 else torch.bfloat16
B, L, D, A = args.batch, args.L, args.D, args.A
g = torch.Generator(device="cuda").manual_seed(0)
ann = torch.randn(B, L, D, device="cuda", generator=g).to(dt)
P = torch.randn(B, L, A, device="cuda", generator=g).to(dt)
hp = torch.randn(B, A + D, device="cuda", generator=g)
wf = torch.randn(A, device="cuda", generator=g)
alpha = torch.empty(B, L, device="cuda")
z = torch.empty(B, D, device="cuda", dtype=dt)
gz = torch.empty_like(z)
beta = torch.empty_like(z)
d = decoder.make_dims(B, B, L, dict(D=D, A=A, E=8, H=8, V=8), 1, dt, args.fp32, False)
lib = _lib.lib()
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda") if args.flush else None


def run():
    _lib.check(lib.sat_attention_step_fwd(C.byref(d), _lib.ptr(ann), _lib.ptr(P), _lib.ptr(wf), _lib.ptr(hp), A + D, None, 0,
                                          _lib.ptr(alpha), L, _lib.ptr(z), _lib.ptr(gz), _lib.ptr(beta), D, _lib.stream_ptr()), "att")


for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(args.iters):
    if flush is not None:
        flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3 / args.reps)
ts.sort()
s = 4 if args.fp32 else 2
nbytes = B * (L * (A + D) * s + 2 * D * s + 4 * L + 4 * (A + D))
med = ts[len(ts) // 2]
print("mode=%s B=%d L=%d D=%d %s flush=%d: median %.1f us  min %.1f us  -> %.0f GB/s (median)" % (
    os.environ.get("SAT_ATT_MODE", "auto"), B, L, D, "fp32" if args.fp32 else "bf16", int(args.flush), med, ts[0], nbytes / med / 1e3))
