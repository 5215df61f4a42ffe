"""Drop-in counterpart of the reference's model.py: get_encoder, InitLSTM, SoftAttention, DeepOutput
and the SAT (Lightning)Module with the same constructor kwargs, attribute / parameter names
(state_dict compatible) and method signatures -- with the per-timestep decoder
hot path running in libsat_b200.so (hand-written sm_100a kernels) instead of ~136 ATen ops / step.

Reference spans mirrored (citations are file:line into the reference repository):
  get_encoder model.py:16-63 (+ readme.md:118-121 encoder_size) | InitLSTM model.py:66-81 |
  SoftAttention model.py:84-109 | DeepOutput model.py:112-131 | SAT model.py:134-817.
The CNN trunk stays on cuDNN through PyTorch (it is the boundary, not the target).
There is no CPU fallback: decoder methods need a CUDA device and the built library.
"""
import inspect
import math

import torch
from torch import nn
from torch.nn.utils.rnn import pack_padded_sequence

from . import _lib, decoder
from .packing import PackedWeights, layers_of, param_names

_HELPER_STREAMS = {}      # device -> (copy stream, result read-back stream) of SAT.caption_stream

try:  # PyTorch-Lightning is optional (not installed in the build image)
    import pytorch_lightning as pl
    _Base = pl.LightningModule
    _HAVE_PL = True
except Exception:  # pragma: no cover - exercised in this image
    pl = None
    _HAVE_PL = False

    class _HParams(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

    class _Base(nn.Module):
        """Minimal stand-in for pl.LightningModule (hparams, device, no-op logging hooks)."""

        def __init__(self):
            super().__init__()
            self.current_epoch = 0
            self.global_step = 0
            self.logger = None
            self.trainer = None

        def save_hyperparameters(self):
            frame = inspect.currentframe().f_back
            object.__setattr__(self, "_hparams", _HParams(frame.f_locals.get("kwargs", {})))

        @property
        def hparams(self):
            return self._hparams

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def optimizers(self):
            return getattr(self, "_optimizer", None)

        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False
            self.eval()

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **overrides):
            """Loads a PyTorch-Lightning checkpoint file (what `SAT.load_from_checkpoint` does in evaluate.ipynb cell 2,
            visualize.ipynb cell 1, temperature_scaling.py:17): a torch-pickled dict with `hyper_parameters` (the
            constructor kwargs captured by save_hyperparameters) and `state_dict`."""
            ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
            hp = dict(ckpt.get("hyper_parameters", ckpt.get("hparams", {})))
            hp.update(overrides)
            hp["pretrained"] = False                      # the weights come from the checkpoint, not from a download
            model = cls(**hp)
            model.load_state_dict(ckpt["state_dict"], strict=strict)
            model.current_epoch = int(ckpt.get("epoch", 0))
            model.global_step = int(ckpt.get("global_step", 0))
            return model


def _hp(args, name, default=None):
    try:
        return getattr(args, name)
    except AttributeError:
        return default


# ------------------------------------------------------------------------------------------------
# encoder (boundary; cuDNN through PyTorch)
# ------------------------------------------------------------------------------------------------
class _Normalize(nn.Module):
    """In-place channel normalisation of [0,1] images: layer 0 of the encoder, as in model.py:59
    (the caller's image tensor is modified)."""

    def __init__(self, mean, std):
        super().__init__()
        self.mean, self.std = list(mean), list(std)
        self._cache = {}            # (device, dtype) -> (mean, std) device tensors: one upload, not a host->device copy per call

    def forward(self, x):
        key = (x.device, x.dtype)
        if key not in self._cache:
            self._cache[key] = (torch.as_tensor(self.mean, dtype=x.dtype, device=x.device).view(1, -1, 1, 1),
                                torch.as_tensor(self.std, dtype=x.dtype, device=x.device).view(1, -1, 1, 1))
        mean, std = self._cache[key]
        return x.sub_(mean).div_(std)


_TRUNK_CUT = (("resnet", -2), ("resnext", -2), ("shufflenet", -1), ("squeezenet", -1), ("densenet", -1),
              ("mobilenet_v2", -1), ("mobilenet_v3", -2), ("mnasnet", -1))


def get_encoder(args):
    """torchvision trunk without pooling/classifier -> annotations [B, encoder_dim, h, w]
    (model.py:16-63).  Writes the trunk width back into args.encoder_dim when no 1x1 conv is
    added (model.py:56).  Optional `encoder_size` appends the resize of readme.md:118-121."""
    from torchvision import models
    arch = args.encoder_arch
    ctor = models.__dict__.get(arch, None)
    if not callable(ctor):
        raise ValueError("Unknown model arg: {}".format(arch))
    pretrained = bool(_hp(args, "pretrained", False))
    try:
        m = ctor(weights="DEFAULT" if pretrained else None)
    except TypeError:
        m = ctor(pretrained=pretrained)
    if pretrained:
        for p in m.parameters():
            p.requires_grad = False
    cut = None
    for key, c in _TRUNK_CUT:
        if key in arch:
            cut = c
            break
    if cut is None:
        raise ValueError("Encoder not supported : {}".format(arch))
    layers = list(m.children())[:cut]
    with torch.no_grad():
        probe = nn.Sequential(*layers)(torch.zeros(1, 3, args.input_size, args.input_size))
    final_dim, final_size = probe.shape[1], probe.shape[2]
    if args.encoder_dim is not None and args.encoder_dim != final_dim:
        layers.append(nn.Conv2d(final_dim, args.encoder_dim, kernel_size=1, stride=1, bias=True))
    else:
        args.encoder_dim = final_dim
    size = _hp(args, "encoder_size", None)
    if size is not None:
        if size < final_size:
            layers.append(nn.AdaptiveAvgPool2d((size, size)))
        elif size > final_size:
            from .encoder_tail import ResizeBilinearNHWC          # nn.Upsample subclass: one library pass for CUDA channels_last maps
            layers.append(ResizeBilinearNHWC((size, size), mode="bilinear", align_corners=False))
    return nn.Sequential(_Normalize(args.mean, args.std), *layers)


# ------------------------------------------------------------------------------------------------
# decoder modules: parameter containers with the reference's names; forward() runs the CUDA kernels
# ------------------------------------------------------------------------------------------------
def _as_bld(annotations, dtype):
    if annotations.dim() == 3:          # [B,L,D] accepted as well
        x = annotations
        return (x if x.dtype == dtype else x.to(dtype)).contiguous(), None
    return decoder.annotations_as_bld(annotations, dtype), annotations.shape[2:]


def _pad8(x, dims):
    """zero-pad the given dims of x up to multiples of 8 (the kernels' storage granularity)."""
    pads = []
    for dim in range(x.dim() - 1, -1, -1):
        n = x.shape[dim]
        pads += [0, ((n + 7) // 8 * 8 - n) if (dim in dims or dim - x.dim() in dims) else 0]
    return torch.nn.functional.pad(x, pads) if any(pads) else x


def _linear(x, weight, bias=None):
    """x [M,K] fp32 cuda, weight [N,K] -> [M,N] through libsat_b200's GEMM core (sat_linear); any K, N (zero-padded to 8s)."""
    import ctypes as C
    N0 = weight.shape[0]
    x = _pad8(x.float(), (-1,)).contiguous()
    w = _pad8(weight.detach().float(), (0, 1)).contiguous()
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device=x.device)
    b = _pad8(bias.detach().float(), (0,)).contiguous() if bias is not None else None
    _lib.check(_lib.lib().sat_linear(_lib.ptr(x), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(out), N, M, N, K,
                                     _lib.SAT_F32, 1, 0, _lib.stream_ptr()), "sat_linear")
    return out[:, :N0] if N0 != N else out


class InitLSTM(nn.Module):
    """model.py:66-81: mean over locations -> factorize -> init -> [B,2H] read back as [2*layers,B,H]."""

    def __init__(self, args, bias=True):
        super().__init__()
        self.decoder_dim = args.decoder_dim
        self.decoder_layers = args.decoder_layers
        self.factorize = nn.Linear(args.encoder_dim, args.embed_dim, bias=bias)
        self.init = nn.Linear(args.embed_dim, 2 * args.decoder_dim * args.decoder_layers, bias=bias)
        self.dropout = nn.Dropout(p=args.dropout)

    @torch.no_grad()
    def forward(self, annotations):
        bld, _ = _as_bld(annotations, torch.float32)
        mean = self.dropout(bld.mean(1))
        out = _linear(_linear(mean, self.factorize.weight, self.factorize.bias), self.init.weight, self.init.bias)
        st = out.reshape(2 * self.decoder_layers, mean.shape[0], self.decoder_dim)
        return st[:self.decoder_layers], st[self.decoder_layers:]


class SoftAttention(nn.Module):
    """model.py:84-109 (additive attention, L^-0.5 score scale, no biases)."""

    def __init__(self, args):
        super().__init__()
        self.encoder_att = nn.Linear(args.encoder_dim, args.attention_dim, bias=False)
        self.decoder_att = nn.Linear(args.decoder_dim, args.attention_dim, bias=False)
        self.f_att = nn.Linear(args.attention_dim, 1, bias=False)

    @torch.no_grad()
    def forward(self, annotations, decoder_hidden):
        """-> (z [B,D], alpha [B,h,w])  (alpha [B,L] for 3-D annotations)."""
        import ctypes as C
        bld, hw = _as_bld(annotations, torch.float32)
        D0 = bld.shape[2]
        bld = _pad8(bld, (-1,)).contiguous()                      # storage dims are multiples of 8: zero channels / units add nothing
        B, L, D = bld.shape
        Wa = _pad8(self.encoder_att.weight.detach().float(), (0, 1))
        A = Wa.shape[0]
        P = _linear(bld.reshape(B * L, D), Wa).reshape(B, L, A)
        hp = torch.zeros(B, A + D, dtype=torch.float32, device=bld.device)
        hp[:, :self.decoder_att.weight.shape[0]] = _linear(decoder_hidden, self.decoder_att.weight)
        up8 = lambda n: (n + 7) // 8 * 8
        d = decoder.make_dims(B, B, L, dict(D=D, A=A, E=8, H=up8(decoder_hidden.shape[1]), V=8), 1, torch.float32, True, False)
        alpha = torch.empty(B, L, dtype=torch.float32, device=bld.device)
        z = torch.empty(B, D, dtype=torch.float32, device=bld.device)
        gz = torch.empty_like(z)
        wf = _pad8(self.f_att.weight.detach().reshape(-1).float(), (0,)).contiguous()
        _lib.check(_lib.lib().sat_attention_step_fwd(C.byref(d), _lib.ptr(bld), _lib.ptr(P), _lib.ptr(wf), _lib.ptr(hp),
                                                     A + D, None, 0, _lib.ptr(alpha), L, _lib.ptr(z), _lib.ptr(gz), None, D,
                                                     _lib.stream_ptr()), "sat_attention_step_fwd")
        z = z[:, :D0] if D0 != D else z
        return z, (alpha.reshape(B, *hw) if hw is not None else alpha)


class DeepOutput(nn.Module):
    """model.py:112-131."""

    def __init__(self, args):
        super().__init__()
        self.deep = args.deep_output
        self.dropout = nn.Dropout(p=args.dropout)
        self.hidden = nn.Linear(args.decoder_dim, args.embed_dim, bias=False)
        if self.deep:
            self.context = nn.Linear(args.encoder_dim, args.embed_dim, bias=False)
        self.output = nn.Linear(args.embed_dim, args.vocab_size, bias=(not args.weight_tying))

    @torch.no_grad()
    def forward(self, prev_embed, hidden, context):
        if self.deep:
            x = torch.tanh(prev_embed.float() + _linear(torch.cat([hidden.float(), context.float()], 1),
                                                        torch.cat([self.hidden.weight, self.context.weight], 1)))
        else:
            x = _linear(hidden, self.hidden.weight)
        return _linear(self.dropout(x), self.output.weight, self.output.bias)


class LabelSmoothing(nn.Module):
    """util.py:91-112 (kept as `criterion` for callers that score packed logits themselves)."""

    def __init__(self, smoothing=0.0):
        super().__init__()
        self.confidence = 1.0 - smoothing
        self.smoothing = smoothing

    def forward(self, x, target):
        lp = torch.log_softmax(x, dim=-1)
        nll = -lp.gather(-1, target.unsqueeze(1)).squeeze(1)
        return (self.confidence * nll + self.smoothing * (-lp.mean(-1))).mean()


# ------------------------------------------------------------------------------------------------
# autograd bridges
# ------------------------------------------------------------------------------------------------
def _packed_for(cfg, W, device):
    """Persistent PackedWeights of the owning module (cfg["owner"]), repacked from the current parameters with one kernel;
    a throw-away object when there is no owner or the parameter set changed shape."""
    owner = cfg.get("owner")
    key = (cfg["dtype"], str(device), tuple((n, tuple(p.shape)) for n, p in sorted(W.items())))
    if owner is not None:
        cached = getattr(owner, "_packed_train", None)
        if cached is not None and cached[0] == key:
            return cached[1].repack(W)
    pw = PackedWeights(W, dtype=cfg["dtype"], device=device, backward=True)
    if owner is not None:
        object.__setattr__(owner, "_packed_train", (key, pw))
    return pw


class _FusedTrainLoss(torch.autograd.Function):
    """loss (+ aux) of one teacher-forced step; backward = hand-written BPTT (sat_train_backward)."""

    @staticmethod
    def forward(ctx, ann, caps, lens, cfg, *params):
        W = {n: p for n, p in zip(param_names(layers_of(len(params))), params) if p is not None}
        pw = _packed_for(cfg, W, ann.device)
        bld = decoder.annotations_as_bld(ann, cfg["dtype"])
        buf = decoder.train_forward(pw, bld, caps, lens, cfg["label_smoothing"], cfg["att_gamma"], exact=cfg["exact"],
                                    use_tc=cfg["use_tc"], logits_f32=False, backward=True, keep_logits=False,
                                    sampled=cfg.get("sampled"), dropout=cfg.get("dropout", (0.0, 0.0, 0)), fuse_ce=True)
        ctx.pw, ctx.buf, ctx.cfg = pw, buf, cfg
        ctx.ann_shape, ctx.ann_dtype = ann.shape, ann.dtype
        ctx.have = [p is not None for p in params]
        out = buf.t["out"]
        aux = out.detach().clone()
        ctx.mark_non_differentiable(aux)
        return out[0].clone(), aux

    @staticmethod
    def backward(ctx, gloss, _gaux):
        cfg = ctx.cfg
        G, d_ann = decoder.train_backward(ctx.pw, ctx.buf, gloss, pad_idx=cfg["pad_idx"], weight_tying=cfg["weight_tying"])
        Bi, D, h, w = ctx.ann_shape
        d_ann = d_ann.reshape(Bi, h, w, D).permute(0, 3, 1, 2).to(ctx.ann_dtype)
        grads = [G.get(n) if have else None for n, have in zip(param_names(layers_of(len(ctx.have))), ctx.have)]
        ctx.buf = None
        return (d_ann, None, None, None, *grads)


class _TrainLogits(torch.autograd.Function):
    """API path of train_batch: returns padded fp32 logits [B,T,V] and alphas [B,T,L]; backward takes
    arbitrary upstream grads for both (the caller computes its own loss, e.g. temperature_scaling.py:38)."""

    @staticmethod
    def forward(ctx, ann, caps, lens, cfg, *params):
        W = {n: p for n, p in zip(param_names(layers_of(len(params))), params) if p is not None}
        pw = _packed_for(cfg, W, ann.device)
        bld = decoder.annotations_as_bld(ann, cfg["dtype"])
        buf = decoder.train_forward(pw, bld, caps, lens, 0.0, 0.0, exact=cfg["exact"], use_tc=cfg["use_tc"],
                                    logits_f32=True, backward=True, keep_logits=True, sampled=cfg.get("sampled"),
                                    dropout=cfg.get("dropout", (0.0, 0.0, 0)))
        ctx.pw, ctx.buf, ctx.cfg = pw, buf, cfg
        ctx.ann_shape, ctx.ann_dtype = ann.shape, ann.dtype
        ctx.have = [p is not None for p in params]
        V0 = pw.dims0["V"]
        logits = buf.t["logits"][:, :, :V0].permute(1, 0, 2).contiguous()          # [B,T,V] fp32 (model.py:504)
        return logits, buf.t["alphas"].clone()

    @staticmethod
    def backward(ctx, glogits, galphas):
        cfg, buf = ctx.cfg, ctx.buf
        V0 = ctx.pw.dims0["V"]
        if V0 != ctx.pw.dims["V"]:
            buf.t["dlogits"].zero_()
        buf.t["dlogits"][:, :, :V0].copy_(glogits.permute(1, 0, 2))
        saved_gamma = buf.c.att_gamma
        buf.c.att_gamma = 0.0
        G, d_ann = decoder.train_backward(ctx.pw, buf, None, pad_idx=cfg["pad_idx"], weight_tying=cfg["weight_tying"],
                                          dalpha_ext=galphas)
        buf.c.att_gamma = saved_gamma
        Bi, D, h, w = ctx.ann_shape
        d_ann = d_ann.reshape(Bi, h, w, D).permute(0, 3, 1, 2).to(ctx.ann_dtype)
        grads = [G.get(n) if have else None for n, have in zip(param_names(layers_of(len(ctx.have))), ctx.have)]
        ctx.buf = None
        return (d_ann, None, None, None, *grads)


# ------------------------------------------------------------------------------------------------
# SAT
# ------------------------------------------------------------------------------------------------
class SAT(_Base):
    """Show, Attend and Tell (model.py:134-817) with the decoder on libsat_b200.so.

    Extra (optional) hparams: encoder_size (readme.md:118-121), precision in {"fp32","bf16"}
    (fp32 = parity mode with exact transcendental paths; bf16 = tcgen05 tensor-core mode).
    """

    def __init__(self, **kwargs):
        super().__init__()
        self.save_hyperparameters()
        hp = self.hparams
        self.scheduler = None
        self.opt_init_lr = None
        for k, v in (("decoder_layers", 1), ("dropout", 0.0), ("embedding_dropout", 0.0), ("label_smoothing", 0.0),
                     ("weight_tying", False), ("deep_output", False), ("embed_norm", None), ("pretrained_embedding", None),
                     ("att_gamma", 1.0), ("decoder_tf", None), ("encoder_size", None), ("precision", "fp32"),
                     ("encoder_finetune_after", -1), ("lr_warmup_steps", 0), ("scheduler", None), ("val_beamk", 3), ("val_max_len", 32),
                     ("save_monitor", None), ("early_stop_monitor", None), ("plateau_monitor", None)):
            if k not in hp:
                hp[k] = v
        assert 0 <= hp.label_smoothing < (hp.vocab_size - 1) / hp.vocab_size
        if not 1 <= hp.decoder_layers <= _lib.SAT_MAX_LAYERS:
            raise NotImplementedError("decoder_layers=%d: the kernels support 1..%d stacked LSTM layers" % (hp.decoder_layers, _lib.SAT_MAX_LAYERS))
        self.criterion = LabelSmoothing(hp.label_smoothing)
        self.special_idxs = [self.stoi("<PAD>"), self.stoi("<START>"), self.stoi("<END>")]
        # module construction order follows model.py:154-195 so that a seeded default init is identical
        self.encoder = get_encoder(hp)
        if self._dtype() == torch.bfloat16 and hp.get("cudnn_batchnorm", True):
            # encoder boundary: PyTorch does not route bf16 batch-norm to cuDNN (ATen's channels_last kernels are 3-4x slower
            # on B200); same parameters / buffers / state_dict keys, training-mode NHWC half-precision inputs go to cuDNN
            from .cudnn_bn import convert_batchnorm, fuse_residual_blocks
            convert_batchnorm(self.encoder)
            if hp.get("cudnn_fuse_blocks", True):
                # ... and the ReLU / residual add that follow a batch-norm in torchvision's ResNet blocks ride in the same cuDNN
                # call (bnOps BN_ACTIVATION / BN_ADD_ACTIVATION); the stem's max-pool goes to cudnnPooling*
                fuse_residual_blocks(self.encoder)
        self.embedding = nn.Embedding(num_embeddings=hp.vocab_size, embedding_dim=hp.embed_dim, max_norm=hp.embed_norm,
                                      padding_idx=self.stoi("<PAD>"))
        self.embedding_dropout = nn.Dropout(p=hp.embedding_dropout)
        if hp.pretrained_embedding is not None:
            import numpy as np
            self.embedding.weight = nn.Parameter(torch.tensor(np.load(hp.pretrained_embedding), dtype=torch.float32))
        self.init_lstm = InitLSTM(hp, bias=True)
        self.lstm = nn.LSTM(input_size=hp.embed_dim + hp.encoder_dim, hidden_size=hp.decoder_dim,
                            num_layers=hp.decoder_layers, bias=True)      # parameter container (names/shapes/init)
        self.attention = SoftAttention(hp)
        self.beta = nn.Sequential(nn.Linear(hp.decoder_dim, hp.encoder_dim, bias=True), nn.Sigmoid())
        fan_in = self.beta[0].weight.shape[1]
        self.beta[0].bias.data.fill_(1 / fan_in)                            # model.py:191-192
        self.output = DeepOutput(hp)
        if hp.weight_tying and hp.deep_output:
            self.output.output.weight = self.embedding.weight
        self._packed_infer = None

    # ---- vocabulary helpers (model.py:202-212) ------------------------------------------------
    def stoi(self, s):
        return int(self.hparams.vocab_stoi.get(s, self.hparams.vocab_stoi["<UNK>"]))

    def itos(self, i):
        return str(self.hparams.vocab_itos.get(int(i), "<UNK>"))

    def decode_seq(self, seq, remove_special=False):
        return [str(self.itos(t)) for t in seq if not (remove_special and t in self.special_idxs)]

    # ---- plumbing -----------------------------------------------------------------------------
    def _dtype(self):
        # the reference trains in fp32 or under PL's 16-bit AMP (train.py --precision 16).  The 16-bit mode here is bf16: same
        # operand width, fp32 accumulation and master weights, no loss scaling needed (fp32's exponent range), and it is what
        # the tcgen05 kernels take.  PL's spellings of mixed precision all select it.
        half = ("bf16", "bfloat16", "bf16-mixed", "16", "16-mixed", "fp16", "half")
        return torch.bfloat16 if str(self.hparams.precision).lower() in half else torch.float32

    def _cfg(self):
        dt = self._dtype()
        return dict(owner=self, dtype=dt, exact=(dt == torch.float32), use_tc=(dt == torch.bfloat16),
                    label_smoothing=float(self.hparams.label_smoothing), att_gamma=float(self.hparams.att_gamma),
                    pad_idx=self.stoi("<PAD>"), weight_tying=bool(self.hparams.weight_tying and self.hparams.deep_output))

    def _dropout_cfg(self):
        """(dropout p, embedding_dropout p, seed): active in train() mode only, like nn.Dropout.  The masks are a pure function
        of (seed, element index) inside the kernels (sat_b200.h); a fresh seed is drawn from torch's CPU generator per step."""
        if not self.training:
            return (0.0, 0.0, 0)
        p, pe = float(self.hparams.dropout), float(self.hparams.embedding_dropout)
        if p <= 0.0 and pe <= 0.0:
            return (0.0, 0.0, 0)
        return (p, pe, int(torch.randint(0, 2 ** 62, (1,)).item()))

    def _renorm_embedding(self, token_ids=None):
        """nn.Embedding(max_norm=embed_norm) renormalises the looked-up rows IN PLACE at every forward (model.py:158-163).
        The kernels gather from a packed copy, so the same side effect is applied here before packing: to the rows of
        `token_ids`, or to every row when the ids are produced on the device (scheduled sampling)."""
        mn = self.hparams.embed_norm
        if mn is None:
            return
        w = self.embedding.weight
        with torch.no_grad():
            ids = torch.arange(w.shape[0], device=w.device) if token_ids is None else token_ids.reshape(-1).to(w.device).unique()
            torch.embedding_renorm_(w, ids, float(mn), 2.0)

    def decoder_weights(self):
        """reference-named decoder parameters in packing.param_names(decoder_layers) order (None where absent)."""
        sd = dict(self.named_parameters())
        if "output.output.weight" not in sd:           # tied: shares embedding.weight
            sd["output.output.weight"] = self.embedding.weight
        return [sd.get(n) for n in param_names(self.hparams.decoder_layers)]

    def encode(self, img):
        """images -> annotations [B,D,h,w]; bf16 mode runs the trunk channels_last under autocast so the
        result is physically the [B,L,D] array the kernels read (zero-copy)."""
        if self._dtype() == torch.bfloat16 and img.is_cuda:
            img = img.contiguous(memory_format=torch.channels_last)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.encoder(img)
        return self.encoder(img)

    # ---- training (model.py:474-628) ----------------------------------------------------------
    @staticmethod
    def _sampling_plan(lengths, T, epsilon):
        """Host-side schedule of model.py:518: steps 0..2 are teacher-forced; for every later step that still has an
        active caption one torch.rand(1) is drawn and the step feeds back argmax(logits[t-1]) when it exceeds epsilon."""
        eps = float(epsilon)
        if eps >= 1.0:
            return None
        max_len = int(lengths.max())
        plan = [0] * T
        for step in range(min(T, max_len)):
            if step <= 2:
                continue
            if not bool(torch.rand(1) <= eps):
                plan[step] = 1
        return plan

    def train_batch(self, batch, epsilon=0):
        """-> (PackedSequence logits fp32, PackedSequence targets, alphas [B,T,L])  (model.py:557)."""
        img, encoded_captions, lengths = batch
        ann = self.encode(img)
        cfg = self._cfg()
        cfg["sampled"] = self._sampling_plan(lengths, encoded_captions.size(2) - 1, epsilon)
        cfg["dropout"] = self._dropout_cfg()
        self._renorm_embedding(None if cfg["sampled"] else encoded_captions[..., :-1])
        logits, alphas = _TrainLogits.apply(ann, encoded_captions, lengths, cfg, *self.decoder_weights())
        caps = encoded_captions.reshape(-1, encoded_captions.size(2))
        lens = lengths.reshape(-1).tolist()
        logits_packed = pack_padded_sequence(logits, lens, batch_first=True, enforce_sorted=False)
        targets_packed = pack_padded_sequence(caps[:, 1:], lens, batch_first=True, enforce_sorted=False)
        return logits_packed, targets_packed, alphas

    def fused_loss(self, batch, epsilon=1.0):
        """Fused forward + loss of one training step: (loss with grad, aux[8] = loss, ce, reg,
        accuracy, 1/ntok, ntok).  This is what training_step runs; epsilon < 1 enables scheduled sampling."""
        img, encoded_captions, lengths = batch
        ann = self.encode(img)
        cfg = self._cfg()
        cfg["sampled"] = self._sampling_plan(lengths, encoded_captions.size(2) - 1, epsilon)
        cfg["dropout"] = self._dropout_cfg()
        self._renorm_embedding(None if cfg["sampled"] else encoded_captions[..., :-1])
        return _FusedTrainLoss.apply(ann, encoded_captions, lengths, cfg, *self.decoder_weights())

    def _epsilon(self):
        hp = self.hparams
        tf = hp.decoder_tf
        if tf is None:
            return 0.0
        if tf == "always":
            return 1.0
        if tf == "linear":
            return 1 - (1 - hp.decoder_tf_min) * self.current_epoch / hp.epochs
        if tf == "inv_sigmoid":
            l = -math.log(hp.decoder_tf_min / (1 - hp.decoder_tf_min))
            g = 5.0
            b = (1 / ((l / g) + 1)) * hp.epochs
            return 1 / (1 + math.exp((g / b) * (self.current_epoch - b)))
        if tf == "exp":
            return math.exp(math.log(hp.decoder_tf_min) / hp.epochs) ** self.current_epoch
        raise ValueError("unknown decoder_tf {}".format(tf))

    def training_step(self, batch, batch_idx):
        hp = self.hparams
        epsilon = self._epsilon()
        if self.global_step == hp.encoder_finetune_after and hp.encoder_finetune_after >= 0:
            for p in self.encoder.parameters():
                p.requires_grad = True
        loss, aux = self.fused_loss(batch, epsilon)
        # device scalars: no sync here.  bad_token_ids = 1.0 when a caption held a word id outside [0, vocab_size): nn.Embedding
        # would have raised; the kernels feed <PAD> for it and flag the step (include/sat_b200.h: out[6])
        metrics = {"loss": loss, "accuracy": aux[3], "epsilon_tf": float(epsilon), "bad_token_ids": aux[6]}
        logger = getattr(self, "logger", None)
        if logger is not None and getattr(logger, "experiment", None) is not None:
            for k, v in metrics.items():
                logger.experiment.add_scalar("{}/train".format(k), float(v), global_step=self.global_step)
        self._step_lr_schedule()
        return metrics

    def _trainer_step(self):
        tr = getattr(self, "trainer", None)
        return int(getattr(tr, "global_step", self.global_step)) if tr is not None else int(self.global_step)

    def _step_lr_schedule(self):
        """learning-rate warm-up and the schedulers that step every batch (model.py:608-617; host-side)."""
        opt = self.optimizers() if callable(getattr(self, "optimizers", None)) else None
        if opt is None or self.opt_init_lr is None:
            return
        step, warm = self._trainer_step(), int(self.hparams.lr_warmup_steps or 0)
        if step < warm:
            lr_scale = min(1, float(step + 1) / warm)
            for pg, init_lr in zip(opt.param_groups, self.opt_init_lr):
                pg["lr"] = lr_scale * init_lr
        elif step > 0 and type(self.scheduler) in (torch.optim.lr_scheduler.CosineAnnealingWarmRestarts, torch.optim.lr_scheduler.OneCycleLR):
            self.scheduler.step()

    def training_epoch_end(self, outputs):
        """epoch means to the logger, learning rate, per-epoch schedulers (model.py:621-635)."""
        logger = getattr(self, "logger", None)
        exp = getattr(logger, "experiment", None) if logger is not None else None
        if outputs and exp is not None:
            for k in outputs[0].keys():
                vals = [float(x[k]) for x in outputs]
                exp.add_scalar("{}/train_epoch".format(k), sum(vals) / len(vals) if vals else 0, global_step=self.current_epoch + 1)
            opt = self.optimizers()
            if opt is not None:
                exp.add_scalar("Learning Rate", opt.param_groups[0]["lr"], global_step=self.current_epoch + 1)
        if type(self.scheduler) in (torch.optim.lr_scheduler.MultiStepLR, torch.optim.lr_scheduler.ExponentialLR):
            self.scheduler.step()

    # ---- validation (model.py:646-718) -----------------------------------------------------------
    def score_captions(self, captions, encoded_captions, lengths, perplexities=None):
        """corpus BLEU-1..4 / GLEU of the generated captions against the references and the best cosine similarity of
        mean word embeddings (model.py:646-682).  BLEU / GLEU: sat_b200/metrics.py (nltk is not needed); the embedding
        means are batched on the device instead of one small launch per reference."""
        from . import metrics as M
        enc = encoded_captions.tolist()
        lens = lengths.tolist() if torch.is_tensor(lengths) else lengths
        references = [[c[1:l] for c, l in zip(refs, lens[i])] for i, refs in enumerate(enc)]
        out = {"bleu1": M.corpus_bleu(references, captions, weights=(1, 0, 0, 0)),
               "bleu2": M.corpus_bleu(references, captions, weights=(0.5, 0.5, 0, 0)),
               "bleu3": M.corpus_bleu(references, captions, weights=(0.33, 0.33, 0.33, 0)),
               "bleu4": M.corpus_bleu(references, captions, weights=(0.25, 0.25, 0.25, 0.25))}
        gleu = M.corpus_gleu(references, captions)
        W = self.embedding.weight.detach()
        dev = W.device
        with torch.no_grad():
            def mean_embed(seqs):                       # [n, E]: mean embedding of each id list (nan for an empty one, like .mean(0))
                n = len(seqs)
                mx = max(1, max((len(q) for q in seqs), default=1))
                ids = torch.zeros(n, mx, dtype=torch.long)
                msk = torch.zeros(n, mx)
                for i, q in enumerate(seqs):
                    if len(q):
                        ids[i, :len(q)] = torch.as_tensor(q, dtype=torch.long)
                        msk[i, :len(q)] = 1
                ids, msk = ids.to(dev), msk.to(dev)
                return (W[ids].float() * msk.unsqueeze(-1)).sum(1) / msk.sum(1, keepdim=True)
            cv = mean_embed(list(captions))                                         # [B, E]
            flat = [r for refs in references for r in refs]
            rv = mean_embed(flat).reshape(len(references), -1, W.shape[1])         # [B, ncap, E]
            cos = torch.nn.functional.cosine_similarity(rv, cv.unsqueeze(1), dim=2)
            cosine_similarity = cos.max(dim=1).values.mean()
        out["cosine_similarity"] = cosine_similarity.item()
        out["gleu"] = gleu
        if type(perplexities) == list:
            out["perplexity"] = sum(perplexities) / len(perplexities)
        return out

    def val_batch(self, batch, beamk=3, max_gen_length=32, temperature=0.5, sample_method="beam", sample_topk=3, decoder_noise=None,
                  rescore_method=None, rescore_reward=0.5):
        img, encoded_captions, lengths = batch
        captions, scores, alphas, perplexities = self.caption(img, beamk, max_gen_length, temperature, sample_method, sample_topk,
                                                              decoder_noise, rescore_method, rescore_reward, return_all=False)
        return self.score_captions(captions, encoded_captions, lengths, perplexities)

    def validation_step(self, batch, batch_idx):
        return self.val_batch(batch, beamk=self.hparams.val_beamk, max_gen_length=self.hparams.val_max_len, temperature=1.0,
                              rescore_method="LN")

    def validation_epoch_end(self, outputs):
        hp = self.hparams
        logger = getattr(self, "logger", None)
        exp = getattr(logger, "experiment", None) if logger is not None else None
        plateau_val = None
        for k in (outputs[0].keys() if outputs else []):
            vals = [x[k] for x in outputs]
            try:
                val = sum(vals) / len(vals)
            except Exception:
                val = 0
            if self.current_epoch != 0 and exp is not None:
                exp.add_scalar("{}/val_epoch".format(k), val, global_step=self.current_epoch + 1)
            if k == hp.save_monitor or k == hp.early_stop_monitor:
                self.log(k, val)
            if k == hp.plateau_monitor:
                plateau_val = val
        if self._trainer_step() >= int(hp.lr_warmup_steps or 0) and type(self.scheduler) is torch.optim.lr_scheduler.ReduceLROnPlateau \
                and plateau_val is not None:
            self.scheduler.step(plateau_val)

    # ---- inference (model.py:214-472) ---------------------------------------------------------
    @torch.no_grad()
    def caption(self, img_tensor, beamk=3, max_gen_length=32, temperature=1.0, sample_method="beam", sample_topk=3,
                decoder_noise=None, rescore_method=None, rescore_reward=0.5, return_all=False):
        self.eval()
        return self.forward(img_tensor, beamk, max_gen_length, temperature, sample_method, sample_topk, decoder_noise,
                            rescore_method, rescore_reward, return_all)

    @torch.no_grad()
    def forward(self, img, beamk=3, max_gen_length=32, temperature=1.0, sample_method="beam", sample_topk=3,
                decoder_noise=None, rescore_method=None, rescore_reward=0.5, return_all=False):
        assert sample_method in ["beam", "multinomial", "topk"]
        from . import decode
        ann = self.encode(img)
        # "multinomial" / "topk" sampling and decoder noise run on the device as Gumbel-top-k / stateless Gaussian draws seeded
        # from torch's CPU generator (same distributions as the reference's torch.multinomial / torch.randn, not its RNG stream)
        return decode.caption_from_annotations(self, ann, beamk, max_gen_length, temperature, rescore_method,
                                               rescore_reward, return_all, sample_method=sample_method, sample_topk=sample_topk,
                                               decoder_noise=decoder_noise)

    @torch.no_grad()
    def caption_stream(self, batches, beamk=3, max_gen_length=32, temperature=1.0, rescore_method=None, rescore_reward=0.5,
                       return_all=False):
        """Bulk captioning (extension; the reference captions one call at a time): `batches` yields image tensors
        ([n,3,S,S], pinned host memory or device).  Yields caption()'s four lists per batch, in order.  The host->device
        copy of batch i+1 runs on a copy stream while batch i is computed, and the device work of batch i+1 is queued
        before the host turns batch i's device arrays into Python lists, so copies, kernels and host work overlap."""
        from . import decode
        self.eval()
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        # the two helper streams are created once per device (module-level, so that the model stays deep-copyable / picklable):
        # the caching allocator keeps one pool per stream, so fresh streams per call would cudaMalloc (and implicitly
        # synchronise) the staging buffers again every time
        if dev not in _HELPER_STREAMS:
            _HELPER_STREAMS[dev] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        copy_stream, side = _HELPER_STREAMS[dev]

        def stage(x):
            with torch.cuda.stream(copy_stream):
                y = x.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return y, ev

        vocab = dict(PAD=self.stoi("<PAD>"), START=self.stoi("<START>"), END=self.stoi("<END>"), UNK=self.stoi("<UNK>"))
        dw = decode.inference_weights(self)
        # results of batch i are read back on a side stream that waits only for batch i's decode: the device-side gather of
        # the selected attention maps and the device->host copies run next to batch i+1's kernels, and the host's
        # synchronous reads do not wait for work queued after them
        def finish(t, hw, done):
            with torch.cuda.stream(side):
                side.wait_event(done)
                return decode.assemble(t, hw, return_all=return_all)

        it = iter(batches)
        nxt = next(it, None)
        staged = stage(nxt) if nxt is not None else None
        pending = None
        while staged is not None:
            img, ev = staged
            cur.wait_event(ev)
            img.record_stream(cur)
            nxt = next(it, None)
            staged = stage(nxt) if nxt is not None else None
            ann = self.encode(img)
            bld = decoder.annotations_as_bld(ann, dw.pw.dtype)
            t = decode.decode_annotations(dw, bld, int(beamk), max_gen_length, temperature, rescore_method, rescore_reward, vocab)
            done = torch.cuda.Event()
            done.record(cur)
            if pending is not None:
                yield finish(*pending)
            pending = (t, tuple(ann.shape[2:]), done)
        if pending is not None:
            yield finish(*pending)

    # ---- optimisers (model.py:720-817; host-side, stock torch) ---------------------------------
    def configure_optimizers(self):
        hp = self.hparams

        def groups(modules, wd, lr):
            decay, no_decay = [], []
            for m in modules:
                for _, p in m.named_parameters():
                    if p.requires_grad:
                        (no_decay if p.dim() == 1 else decay).append(p)
            return [{"params": no_decay, "lr": lr, "weight_decay": 0.0}, {"params": decay, "lr": lr, "weight_decay": wd}]

        wd = _hp(hp, "weight_decay", 0.0)
        lr = _hp(hp, "decoder_lr", 1e-3)
        params = groups([self.init_lstm, self.lstm, self.attention, self.beta, self.output], wd, lr)
        if _hp(hp, "embedding_lr", lr) > 0 and not hp.weight_tying:
            params += [{"params": list(self.embedding.parameters()), "lr": _hp(hp, "embedding_lr", lr), "weight_decay": 0.0}]
        # the reference adds the encoder group when encoder_finetune_after > 0 and encoder_lr > 0 (model.py:744); a trunk that
        # trains from scratch (pretrained=False: every parameter requires grad, model.py:23-25) is included as well, otherwise
        # its gradients would be computed and never applied
        if _hp(hp, "encoder_lr", 0.0) > 0 and (_hp(hp, "encoder_finetune_after", -1) > 0 or not hp.pretrained):
            params += groups([self.encoder], wd, hp.encoder_lr)
        opt = _hp(hp, "opt", "adam")
        # same update rule as the reference's torch.optim calls; on CUDA the single-kernel ("fused") implementation of the
        # stock optimizer replaces the default multi-pass one (same arithmetic in fp32, one read / write of p, g, m, v)
        every = [p for g in params for p in g["params"]]
        fused = {"fused": True} if (every and all(p.is_cuda and p.is_floating_point() for p in every)) else {}
        if opt == "sgd":
            optimizer = torch.optim.SGD(params, lr=lr, momentum=_hp(hp, "momentum", 0.9), nesterov=_hp(hp, "nesterov", False))
        elif opt == "adamw":
            optimizer = torch.optim.AdamW(params, lr=lr, betas=(_hp(hp, "adam_b1", 0.9), _hp(hp, "adam_b2", 0.999)), **fused)
        else:
            optimizer = torch.optim.Adam(params, lr=lr, betas=(_hp(hp, "adam_b1", 0.9), _hp(hp, "adam_b2", 0.999)), **fused)
        self.opt_init_lr = [pg["lr"] for pg in optimizer.param_groups]
        self._optimizer = optimizer
        self.scheduler = self._build_scheduler(optimizer)
        return optimizer

    def _build_scheduler(self, optimizer):
        """the five schedules of model.py:765-815 on stock torch schedulers"""
        hp = self.hparams
        S = torch.optim.lr_scheduler
        kind = _hp(hp, "scheduler", None)
        if kind == "step":
            return S.MultiStepLR(optimizer, milestones=hp.milestones, gamma=hp.lr_gamma)
        if kind == "plateau":
            return S.ReduceLROnPlateau(optimizer, mode="max", factor=hp.lr_gamma, patience=hp.plateau_patience, min_lr=hp.min_lr)
        if kind == "exp":
            return S.ExponentialLR(optimizer, gamma=hp.lr_gamma)
        if kind == "cosine":
            # restarts sized so that the last cycle ends with the run (model.py:785-805)
            adj_steps = hp.epochs * hp.train_loader_len - hp.lr_warmup_steps
            t0, tm, acc = hp.cosine_iterations, hp.cosine_multi, _hp(hp, "accumulate", 1)
            if tm != 1:
                restarts = math.floor(math.log(1 - (adj_steps * (1 - tm) / t0)) / math.log(tm))
                t0 = adj_steps + acc if restarts == 0 else math.ceil((adj_steps + acc) / ((1 - tm ** restarts) / (1 - tm)))
            else:
                restarts = math.floor(adj_steps / t0)
                t0 = adj_steps + acc if restarts == 0 else math.ceil((adj_steps + acc) / restarts)
            return S.CosineAnnealingWarmRestarts(optimizer, T_0=int(t0), T_mult=int(tm), eta_min=hp.min_lr)
        if kind == "one_cycle":
            hp.lr_warmup_steps = 0
            return S.OneCycleLR(optimizer, self.opt_init_lr, epochs=hp.epochs, steps_per_epoch=hp.train_loader_len,
                                pct_start=hp.one_cycle_pct, cycle_momentum=False, div_factor=hp.one_cycle_div,
                                final_div_factor=hp.one_cycle_fdiv)
        return None
