"""Decoder-only workload at BASELINE configs[1] dims (B=256, L=196, D=512, A=128, E=256, H=512, V=6400, T=20, bf16)
for ncu captures: N iterations of fused forward + loss + BPTT (+ parameter-gradient GEMMs) on synthetic annotations.
    python tools/decoder_step.py [--iters 3] [--fp32] [--decode K]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--fp32", action="store_true")
ap.add_argument("--no-tc", action="store_true")
ap.add_argument("--decode", type=int, default=0, help="run batched decode with this beam width instead of training")
ap.add_argument("--dims", type=str, default="", help="D,A,E,H,V,T,L override, e.g. 2048,128,256,1024,6400,20,196 (configs[2])")
ap.add_argument("--rand-start", action="store_true", help="random word instead of <START> in column 0 (no heavy word)")
ap.add_argument("--no-fuse-ce", action="store_true", help="unfused vocabulary GEMM + ce_rows_kernel")
ap.add_argument("--profile", type=int, default=0, help="also report the in-situ per-launch time of kernel kind 1 (att fwd) / 2 (att bwd) / 3 (vocab GEMM)")
args = ap.parse_args()

from oracle import sat_oracle as O  # noqa: E402  (weights generator only)
from sat_b200 import decode, decoder  # noqa: E402
from sat_b200.packing import PackedWeights  # noqa: E402

D, A, E, H, V, T, L = 512, 128, 256, 512, 6400, 20, 196
if args.dims:
    D, A, E, H, V, T, L = (int(x) for x in args.dims.split(","))
dtype = torch.float32 if args.fp32 else torch.bfloat16
W = O.random_weights(D, A, E, H, V, seed=0)
g = torch.Generator(device="cuda").manual_seed(0)
B = args.batch
ann = torch.randn(B, L, D, device="cuda", generator=g).to(dtype)
caps = torch.randint(1, V - 3, (B, T + 1), device="cuda", generator=g)
if not args.rand_start:
    caps[:, 0] = V - 2
lens = torch.full((B,), T, device="cuda")
use_tc = (not args.fp32) and (not args.no_tc)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
times = []
if args.decode:
    dw = decode.DecodeWeights(W, dtype, torch.device("cuda"), args.fp32, use_tc)
    vocab = dict(PAD=0, UNK=V - 3, START=V - 2, END=V - 1)
    for it in range(args.iters):
        ev0.record()
        t = decode.decode_annotations(dw, ann, args.decode, 30, 1.0, None, 0.5, vocab)
        ev1.record()
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1))
    if args.profile:
        from sat_b200 import _lib
        _lib.profile_begin(args.profile)
        t = decode.decode_annotations(dw, ann, args.decode, 30, 1.0, None, 0.5, vocab)
        torch.cuda.synchronize()
        ms, n = _lib.profile_end()
        print("kind %d: %d launches, %.2f us per launch" % (args.profile, n, 1e3 * ms / max(n, 1)))
    times = sorted(times[2:] or times)
    print("decode k=%d B=%d: median %.3f ms (min %.3f) over %d iters" % (args.decode, B, times[len(times) // 2], times[0], len(times)))
else:
    pw = PackedWeights(W, dtype=dtype, device="cuda", backward=True)
    import time
    cpu = []
    for it in range(args.iters):
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        ev0.record()
        buf = decoder.train_forward(pw, ann, caps, lens, 0.0, 1.0, exact=args.fp32, use_tc=use_tc, backward=True, fuse_ce=not args.no_fuse_ce)
        G, d_ann = decoder.train_backward(pw, buf)
        ev1.record()
        cpu.append(1e3 * (time.perf_counter() - c0))      # host time to ISSUE the step (no synchronize inside)
        torch.cuda.synchronize()
        times.append(ev0.elapsed_time(ev1))
    cpu = sorted(cpu[2:] or cpu)
    print("host issue time per step: median %.3f ms (min %.3f)" % (cpu[len(cpu) // 2], cpu[0]))
    if args.profile:
        from sat_b200 import _lib
        _lib.profile_begin(args.profile)
        for it in range(3):
            buf = decoder.train_forward(pw, ann, caps, lens, 0.0, 1.0, exact=args.fp32, use_tc=use_tc, backward=True, fuse_ce=not args.no_fuse_ce)
            G, d_ann = decoder.train_backward(pw, buf)
        torch.cuda.synchronize()
        ms, n = _lib.profile_end()
        print("kind %d: %d launches, %.2f us per launch" % (args.profile, n, 1e3 * ms / max(n, 1)))
    times = sorted(times[2:] or times)
    print("train fwd+bwd B=%d: median %.3f ms (min %.3f) over %d iters  loss %.4f" % (B, times[len(times) // 2], times[0], len(times), float(buf.t["out"][0])))
