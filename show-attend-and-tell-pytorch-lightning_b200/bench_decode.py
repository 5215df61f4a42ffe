"""Decode workloads of bench.py (BASELINE.json configs[3] greedy and configs[4] beam search).

  greedy: resnet50 encoder (1x1 conv to D=512, encoder_size 14 -> L=196), H=512, V=6400, max_gen_length 30, batch 1024 / GPU
  beam  : wide_resnet101_2 encoder (D=2048 native, encoder_size 16 -> L=256), H=512, V=10000, k=5, max_gen_length 30, batch 256 / GPU

`value` = captions/s of the decode itself with annotations resident in HBM (sat_decode: 31 steps of the fused decoder
+ device-side beam bookkeeping); `e2e` = SAT.caption(images on pinned host memory) -> the reference's four Python lists
(encoder, H2D of the images, D2H of tokens / scores / alphas included).  Decode shards by image: ranks are independent.
"""
import json
import os
import statistics
import time

import torch

METRIC = "captions_per_sec"

CFG = {
    "greedy": dict(arch="resnet50", D=512, size=14, L=196, H=512, A=128, E=256, V=6400, k=1, S=30, B=1024),
    "beam": dict(arch="wide_resnet101_2", D=2048, size=16, L=256, H=512, A=128, E=256, V=10000, k=5, S=30, B=256),
}


def _vocab(V):
    stoi = {"<PAD>": 0}
    for i in range(1, V - 3):
        stoi["w%d" % i] = i
    stoi["<UNK>"], stoi["<START>"], stoi["<END>"] = V - 3, V - 2, V - 1
    return stoi, {v: k for k, v in stoi.items()}


def _hparams(c, precision):
    stoi, itos = _vocab(c["V"])
    return dict(encoder_arch=c["arch"], pretrained=False, input_size=224, encoder_dim=c["D"], encoder_size=c["size"],
                mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], embed_dim=c["E"], embed_norm=None,
                attention_dim=c["A"], decoder_dim=c["H"], decoder_layers=1, dropout=0.0, embedding_dropout=0.0,
                label_smoothing=0.0, weight_tying=False, deep_output=True, vocab_size=c["V"], vocab_stoi=stoi,
                vocab_itos=itos, pretrained_embedding=None, att_gamma=1.0, decoder_tf="always", precision=precision)


def _config(args, c):
    return {"workload": "%s decode: %s encoder, L=%d, D=%d, H=%d, V=%d, beamk=%d, max_gen_length=%d, batch %d images per GPU"
                        % (args.workload, c["arch"], c["L"], c["D"], c["H"], c["V"], c["k"], c["S"], c["B"]),
            "global_batch": c["B"] * args.gpus, "parallelism": "independent shards x%d" % args.gpus,
            "l2": "annotations + P (%.0f MB) exceed or rival L2; logits [rows,V] are rewritten every step"
                  % (c["B"] * c["L"] * (c["D"] + c["A"]) * 2 / 1e6)}


def reference_line(args):
    """CPU arm: oracle port of SAT.caption (one image at a time, like the reference) on a bounded sample."""
    from oracle import sat_oracle as O
    c = CFG[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = 4 if args.workload == "beam" else 8
    W = O.random_weights(c["D"], c["A"], c["E"], c["H"], c["V"], seed=0)
    enc = O.build_encoder(c["arch"], c["D"], c["size"]).eval()
    g = torch.Generator().manual_seed(1)
    img = torch.rand(n, 3, 224, 224, generator=g)
    vocab = dict(PAD=0, UNK=c["V"] - 3, START=c["V"] - 2, END=c["V"] - 1)
    times = []
    with torch.no_grad():
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            ann = enc(img.clone())
            O.caption(W, ann, vocab, beamk=c["k"], max_gen_length=c["S"])
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    v = n * len(times) / sum(times)
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": "captions/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _config(args, c),
            "cpu_baseline": {"value": v, "unit": "captions/s", "cores": cores, "kind": "port",
                             "sample": "oracle port of SAT.caption (encoder + per-image beam loop) on %d images per step" % n},
            "e2e": {"value": v, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run(args):
    import torch.distributed as dist
    from . import _lib, decode, decoder
    from .model import SAT
    c = CFG[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = SAT(**_hparams(c, args.precision)).to(dev).eval()
    if args.precision == "bf16":
        model.encoder.to(memory_format=torch.channels_last)
    B = c["B"] if args.batch == 256 else args.batch          # --batch overrides only when given
    g = torch.Generator().manual_seed(100 + rank)
    img_h = torch.rand(B, 3, 224, 224, generator=g).pin_memory()
    dw = decode.inference_weights(model)
    with torch.no_grad():
        chunks = [model.encode(img_h[i:i + 128].to(dev)) for i in range(0, B, 128)]
        ann = torch.cat(chunks, 0)
    bld = decoder.annotations_as_bld(ann, dw.pw.dtype)
    vocab = dict(PAD=0, UNK=c["V"] - 3, START=c["V"] - 2, END=c["V"] - 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return decode.decode_annotations(dw, bld, c["k"], c["S"], 1.0, None, 0.5, vocab)

    def step_e2e():
        # public bulk-captioning call: pinned host images in chunks of 256 (encoder activation memory); copies, kernels
        # and the host-side list assembly of consecutive chunks overlap inside caption_stream
        chunks = (img_h[i:i + 256] for i in range(0, B, 256))
        return list(model.caption_stream(chunks, beamk=c["k"], max_gen_length=c["S"]))

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            fn()
            evs[i + 1].record()
        barrier()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        tot = evs[0].elapsed_time(evs[steps])
        if world > 1:
            t = torch.tensor([tot], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot = float(t.item())
        return tot, per

    from bench import ClockSampler
    clocks = ClockSampler(local)
    l0 = _lib.launch_count()
    clocks.start()
    tot_ms, per = timed(step_device, args.steps, args.warmup)
    clk = clocks.stop()
    launches = (_lib.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    value = B * world * args.steps / (tot_ms * 1e-3)
    t0 = time.perf_counter()
    e2e_ms, _ = timed(step_e2e, max(2, args.steps // 3), 1)
    e2e_steps = max(2, args.steps // 3)
    e2e_value = B * world * e2e_steps / (e2e_ms * 1e-3)

    _lib.profile_begin(1)
    for _ in range(2):
        step_device()
    att_ms, att_n = _lib.profile_end()
    s = 2 if dw.pw.dtype == torch.bfloat16 else 4
    R = B * c["k"]
    att_bytes = B * c["L"] * (c["A"] + c["D"]) * s + R * ((c["H"] + 2 * c["D"]) * s + 4 * c["L"])   # per image-step + per row
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = att_bytes / (att_ms / max(att_n, 1) * 1e-3) / 1e9 if att_n else None
    line = {
        "metric": METRIC, "value": value, "unit": "captions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_ms / args.steps, "step_p50_ms": statistics.median(per), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": _config(args, dict(c, B=B)), "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "captions/s", "h2d_bytes_per_step": img_h.numel() * 4,
                "d2h_bytes_per_step": int(B * (c["S"] + 1) * 4 * 2 + B * c["S"] * c["L"] * 4), "ms_per_step": e2e_ms / e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "attention_step_fwd_group_kernel" if c["k"] > 1 else "attention_step_fwd_pipe_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": None, "launches_timed": att_n,
                     "avg_launch_us": 1e3 * att_ms / max(att_n, 1), "algorithmic_bytes_per_launch": att_bytes,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
    }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            a2 = type("A", (), dict(workload=args.workload, steps=2, warmup=1, gpus=1))()
            line["cpu_baseline"] = reference_line(a2)["cpu_baseline"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
