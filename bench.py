#!/usr/bin/env python
"""Benchmark of the SAT decoder hot path on B200 (BASELINE.json metric: captions/sec, step p50 ms).

    python bench.py --gpus N --steps K --warmup W [--workload train|greedy|beam] [--impl reference]

N=1 default workload = BASELINE.json configs[1]: SAT resnet50 encoder, L=196, D=512, hidden 512,
vocab 6400, batch 256, bf16 training step (encoder fwd + decoder fwd + loss + full backward +
optimizer step).  For N>1 launch with torch.distributed.run; each rank keeps batch 256 (weak scaling),
gradients are averaged with NCCL all-reduce.

`--impl reference` times the reference's algorithm on the host CPU (oracle port of model.py, the
reference itself is Python and /root/reference does not travel to the GPU box) on a bounded sample.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "captions_per_sec"
DIMS = dict(L=196, D=512, A=128, E=256, H=512, V=6400, T=20)


def vocab(V):
    stoi = {"<PAD>": 0}
    for i in range(1, V - 3):
        stoi["w%d" % i] = i
    stoi["<UNK>"], stoi["<START>"], stoi["<END>"] = V - 3, V - 2, V - 1
    return stoi, {v: k for k, v in stoi.items()}


def hparams(arch="resnet50", precision="bf16", **over):
    stoi, itos = vocab(DIMS["V"])
    hp = dict(encoder_arch=arch, pretrained=False, input_size=224, encoder_dim=DIMS["D"], encoder_size=14,
              mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], embed_dim=DIMS["E"], embed_norm=None,
              attention_dim=DIMS["A"], decoder_dim=DIMS["H"], decoder_layers=1, dropout=0.0, embedding_dropout=0.0,
              label_smoothing=0.0, weight_tying=False, deep_output=True, vocab_size=DIMS["V"], vocab_stoi=stoi,
              vocab_itos=itos, pretrained_embedding=None, att_gamma=1.0, decoder_tf="always", precision=precision,
              opt="adam", decoder_lr=4e-4, embedding_lr=4e-4, encoder_lr=1e-4, weight_decay=0.0,
              encoder_finetune_after=-1)
    hp.update(over)
    return hp


def synth_batch(B, T, V, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, 224, 224, generator=g)
    caps = torch.randint(1, V - 3, (B, 1, T + 1), generator=g)
    caps[:, :, 0] = V - 2
    caps[:, :, T] = V - 1
    lens = torch.full((B, 1), T, dtype=torch.long)
    if pin:
        img, caps, lens = img.pin_memory(), caps.pin_memory(), lens.pin_memory()
    if device != "cpu":
        img, caps, lens = img.to(device), caps.to(device), lens.to(device)
    return img, caps, lens


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's train step on host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_baseline(steps, warmup, sample_B=8, arch="resnet50"):
    from oracle import sat_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    enc = O.build_encoder(arch, DIMS["D"], 14)
    enc.train()
    W = O.random_weights(DIMS["D"], DIMS["A"], DIMS["E"], DIMS["H"], DIMS["V"], seed=0)
    W = {k: v.requires_grad_(True) for k, v in W.items()}
    params = list(enc.parameters()) + list(W.values())
    opt = torch.optim.Adam(params, lr=1e-4)
    img, caps, lens = synth_batch(sample_B, DIMS["T"], DIMS["V"], seed=1)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        ann = enc(img.clone())
        r = O.train_loss(W, ann, caps, lens, 0.0, 1.0)
        r["loss"].backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    tot = sum(times)
    return dict(value=sample_B * len(times) / tot, ms_per_step=1e3 * tot / len(times), cores=cores, sample_B=sample_B,
                p50_ms=1e3 * statistics.median(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload != "train":
        from sat_b200 import bench_decode
        print(json.dumps(bench_decode.reference_line(args)), flush=True)
        return
    sample_B = 8
    r = cpu_train_baseline(args.steps, args.warmup, sample_B)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "captions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "step_p50_ms": r["p50_ms"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, 256),
        "cpu_baseline": {"value": r["value"], "unit": "captions/s", "cores": r["cores"], "kind": "port",
                         "sample": "oracle port of model.py train step (resnet50 encoder fwd+bwd, decoder fwd+loss+bwd, Adam) "
                                   "at batch %d per step, fp32, torch CPU" % sample_B},
        "e2e": {"value": r["value"], "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def config_dict(args, B):
    return {"workload": "train_step: SAT resnet50 encoder (pretrained=False, encoder_size=14 -> L=196), D=512, A=128, E=256, "
                        "H=512, V=6400, T=20 targets, batch %d per GPU, teacher-forced fwd+loss+bwd+Adam" % B,
            "global_batch": B * args.gpus, "caption_len": DIMS["T"], "parallelism": "dp%d" % args.gpus,
            "l2": "working set > L2: encoder activations of a batch-256 ResNet-50 step (GBs) are rewritten every step"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_train(args):
    import torch.distributed as dist
    from sat_b200 import _lib, decoder
    from sat_b200.dist import FlatGradBuckets
    from sat_b200.model import SAT
    from sat_b200.packing import PackedWeights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, V = args.batch, DIMS["T"], DIMS["V"]
    torch.manual_seed(0)                       # identical replicas
    model = SAT(**hparams(precision=args.precision)).to(dev)
    if args.precision == "bf16":
        model.encoder.to(memory_format=torch.channels_last)
    model.train()
    opt = model.configure_optimizers()
    enc_params = [p for p in model.encoder.parameters() if p.requires_grad]
    dec_params = [p for n, p in model.named_parameters() if not n.startswith("encoder.") and p.requires_grad]
    buckets = FlatGradBuckets(dec_params + enc_params) if world > 1 else None
    img_d, caps_d, lens_d = synth_batch(B, T, V, seed=100 + rank, device=dev)
    img_h, caps_h, lens_h = synth_batch(B, T, V, seed=100 + rank, pin=True)

    def step_device():
        if buckets is not None:
            buckets.zero()
        else:
            opt.zero_grad(set_to_none=True)
        loss, aux = model.fused_loss((img_d.clone(), caps_d, lens_d))
        loss.backward()
        if buckets is not None:
            buckets.allreduce_mean(world)
        opt.step()
        return loss

    # End-to-end step = what a training loop around the public API does: the NEXT batch's host->device copy is issued on a
    # copy stream while the current step computes (input prefetch), and each step's loss is read back through a pinned
    # buffer one step late, so neither copy stalls the launch queue.  Every timed step still issues one H2D copy of a
    # full batch and one D2H read of a loss; the last loss is drained before the closing event.
    copy_stream = torch.cuda.Stream()
    pending = {}
    loss_pin = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "last": None}

    def prefetch():
        with torch.cuda.stream(copy_stream):
            pending["batch"] = (img_h.to(dev, non_blocking=True), caps_h.to(dev, non_blocking=True), lens_h.to(dev, non_blocking=True))
            pending["ev"] = torch.cuda.Event()
            pending["ev"].record(copy_stream)

    def step_e2e():
        if "batch" not in pending:
            prefetch()
        cur = torch.cuda.current_stream()
        cur.wait_event(pending["ev"])
        img, caps, lens = pending.pop("batch")
        for x in (img, caps, lens):
            x.record_stream(cur)
        prefetch()                                  # next batch's H2D overlaps this step's compute
        if buckets is not None:
            buckets.zero()
        else:
            opt.zero_grad(set_to_none=True)
        m = model.training_step((img, caps, lens), 0)
        m["loss"].backward()
        if buckets is not None:
            buckets.allreduce_mean(world)
        opt.step()
        i = e2e_state["i"]
        loss_pin[i & 1].copy_(m["loss"].detach().reshape(1).float(), non_blocking=True)     # device -> host read of the step's result
        loss_ev[i & 1].record()
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
            e2e_state["last"] = float(loss_pin[(i - 1) & 1][0])
        e2e_state["i"] = i + 1
        return e2e_state["last"]

    def drain_e2e():
        i = e2e_state["i"]
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
            e2e_state["last"] = float(loss_pin[(i - 1) & 1][0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, per_step=False, drain=None):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            fn()
            if drain is not None and i == steps - 1:
                drain()
            evs[i + 1].record()
        barrier()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        tot = evs[0].elapsed_time(evs[steps])
        if world > 1:
            t = torch.tensor([tot], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot = float(t.item())
        return tot, per

    clocks = ClockSampler(local)
    l0 = _lib.launch_count()
    clocks.start()
    tot_ms, per = timed(step_device, args.steps, args.warmup)
    clk = clocks.stop()
    launches = (_lib.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    value = B * world * args.steps / (tot_ms * 1e-3)
    e2e_ms, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2), drain=drain_e2e)
    e2e_value = B * world * args.steps / (e2e_ms * 1e-3)
    h2d = img_h.numel() * 4 + caps_h.numel() * 8 + lens_h.numel() * 8

    # decoder-only (annotations resident): fwd + loss + BPTT + parameter-gradient GEMMs
    with torch.no_grad():
        ann = model.encode(img_d.clone())
    cfg = model._cfg()
    W = {n: p for n, p in zip(__import__("sat_b200.packing", fromlist=["PARAM_NAMES"]).PARAM_NAMES, model.decoder_weights())
         if p is not None}
    bld = decoder.annotations_as_bld(ann, cfg["dtype"])

    def dec_step():
        pw = PackedWeights(W, dtype=cfg["dtype"], device=dev, backward=True)
        buf = decoder.train_forward(pw, bld, caps_d, lens_d, 0.0, 1.0, exact=cfg["exact"], use_tc=cfg["use_tc"], backward=True)
        decoder.train_backward(pw, buf)

    dec_ms, dec_per = timed(dec_step, args.steps, args.warmup)

    # roofline of the dominant decoder kernel (fused attention step, HBM-bound): device time of its launches
    _lib.profile_begin(1)
    for _ in range(3):
        dec_step()
    att_ms, att_n = _lib.profile_end()
    s = 2 if cfg["dtype"] == torch.bfloat16 else 4
    att_bytes = B * (DIMS["L"] * (DIMS["A"] + DIMS["D"]) * s + (DIMS["H"] + 2 * DIMS["D"]) * s + 4 * DIMS["L"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = att_bytes / (att_ms / max(att_n, 1) * 1e-3) / 1e9 if att_n else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "attention_fwd_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    roof = {"kernel": "attention_step_fwd_pipe_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "launches_timed": att_n,
            "avg_launch_us": 1e3 * att_ms / max(att_n, 1), "algorithmic_bytes_per_launch": att_bytes,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}

    # the backward attention kernel is the largest single decoder item by time; report it next to the forward kernel
    _lib.profile_begin(2)
    for _ in range(3):
        dec_step()
    attb_ms, attb_n = _lib.profile_end()
    attb_bytes = B * (DIMS["L"] * (DIMS["A"] + DIMS["D"]) * s + 2 * DIMS["D"] * s + 8 * DIMS["L"])
    attb = attb_bytes / (attb_ms / max(attb_n, 1) * 1e-3) / 1e9 if attb_n else None
    roof["other_kernels"] = [{"kernel": "attention_step_bwd_pipe_kernel", "bound": "hbm", "achieved": attb, "peak": peak,
                              "unit": "GB/s", "frac": (attb / peak) if attb else None, "launches_timed": attb_n,
                              "avg_launch_us": 1e3 * attb_ms / max(attb_n, 1), "algorithmic_bytes_per_launch": attb_bytes}]

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "captions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_ms / args.steps, "step_p50_ms": statistics.median(per), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": config_dict(args, B), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "captions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "decoder_only": {"value": B * args.steps / (dec_ms * 1e-3), "unit": "captions/s", "ms_per_step": dec_ms / args.steps,
                             "p50_ms": statistics.median(dec_per), "what": "decoder fwd+loss+BPTT+param-grad GEMMs, annotations resident"},
            "roofline": roof,
        }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_train_baseline(steps=2, warmup=1, sample_B=8)
            line["cpu_baseline"] = {"value": r["value"], "unit": "captions/s", "cores": r["cores"], "kind": "port",
                                    "sample": "oracle port of the reference train step at batch 8 (same dims), 2 timed steps after 1 warm-up, fp32 torch CPU"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "greedy", "beam"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "train":
        return run_train(args)
    from sat_b200 import bench_decode
    return bench_decode.run(args)


if __name__ == "__main__":
    main()
