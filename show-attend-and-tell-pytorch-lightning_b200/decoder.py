"""Host side of the fused SAT decoder: buffer management and the calls into libsat_b200.so.

Reference spans replaced: SAT.train_batch (model.py:474-557) + the loss of training_step
(model.py:588-597, util.py:105-112) and, through autograd, their backward.
"""
import ctypes as C

import torch

from . import _lib, _redzone
from .packing import PackedWeights, deinterleave_gates


def make_dims(B, Bi, L, D, A, E, H, V, T, dtype, exact, use_tc, plain_output=False):
    d = _lib.SatDims()
    d.plain_output = 1 if plain_output else 0
    d.B, d.Bi, d.ncap = B, Bi, B // Bi
    d.L, d.D, d.A, d.E, d.H, d.V, d.T = L, D, A, E, H, V, T
    d.dtype = _lib.dtype_code(dtype)
    d.exact = 1 if exact else 0
    d.use_tc = 1 if use_tc else 0
    return d


class TrainBuffers:
    """All device buffers of one training step (include/sat_b200.h: SatTrainBuffers)."""

    def __init__(self, d, dtype, device, logits_f32=False, backward=True, keep_logits=False):
        B, Bi, L, D, A, E, H, V, T = d.B, d.Bi, d.L, d.D, d.A, d.E, d.H, d.V, d.T
        NH3 = A + D + 4 * H
        s, f = dtype, torch.float32
        t = self.t = {}
        mk = lambda shape, dt: _redzone.empty(shape, dt, device, "train buffer #%d" % len(t))   # torch.empty unless SAT_REDZONE=1
        t["tok"] = mk((T, B), torch.int32)
        t["P"] = mk((Bi, L, A), s)
        t["meanv"] = mk((Bi, D), s)
        t["f1"] = mk((Bi, E), s)
        t["init_out"] = mk((Bi, 2 * H), f)
        t["Xe"] = mk((T, B, E), s)
        t["Gx"] = mk((T, B, 4 * H), f)
        t["Hs"] = mk((T + 1, B, H), s)
        t["Cs"] = mk((T + 1, B, H), f)
        t["hp"] = mk((B, NH3), f)
        t["Q"] = mk((T, B, A), f)
        t["alphas"] = mk((B, T, L), f)
        t["Z"] = mk((T, B, D), s)
        t["GZ"] = mk((T, B, D), s)
        t["Beta"] = mk((T, B, D), s)
        t["Gates"] = mk((T, B, 4 * H), s)
        t["Xo"] = mk((T, B, E), s)
        t["logits"] = mk((T, B, V), f if logits_f32 else s)
        if backward:
            t["dlogits"] = mk((T, B, V), s) if (keep_logits or logits_f32) else t["logits"]
        t["row_loss"] = mk((T, B), f)
        t["row_argmax"] = mk((T, B), torch.int32)
        t["S"] = mk((B, L), f)
        t["out"] = torch.zeros(8, dtype=f, device=device)
        if backward:
            t["gscale"] = torch.ones(1, dtype=f, device=device)
            t["dpre"] = mk((T, B, E), s)
            t["dHZ"] = mk((T, B, H + D), f)
            t["DY"] = mk((T, B, NH3), s)
            t["dgz"] = mk((16, B, D), f)        # split-K partials (SAT_MAX_SPLITK)
            t["dh"] = mk((16, B, H), f)
            t["dc"] = mk((B, H), f)
            t["dZ"] = mk((T, B, D), s)
            t["dP"] = mk((B, L, A), f)
            if dtype != torch.float32:
                t["dP16"] = mk((B, L, A), s)
            t["dwf_part"] = mk((T, B, A), f)
            t["de"] = mk((T, B, L), f)
            t["dXe"] = mk((T, B, E), f)
            t["d_init_out"] = mk((Bi, 2 * H), f)
            t["df1"] = mk((Bi, E), f)
            if dtype != torch.float32:
                t["d_init_out16"] = mk((Bi, 2 * H), s)
                t["df116"] = mk((Bi, E), s)
            t["dmean"] = mk((Bi, D), f)
            t["d_ann"] = mk((B, L, D), s)          # per caption row; the host sums the ncap rows of an image
        self.c = _lib.SatTrainBuffers()
        for name, typ in _lib.SatTrainBuffers._fields_:
            if typ is C.c_void_p and name in t:
                setattr(self.c, name, _lib.ptr(t[name]))
        self.c.logits_f32 = 1 if logits_f32 else 0
        self.dims = d

    def bind_inputs(self, ann, caps, lens, label_smoothing, att_gamma, sampled=None, dropout=(0.0, 0.0, 0)):
        self.t["ann"], self.t["caps"], self.t["lens"] = ann, caps, lens   # keep alive
        self.c.ann, self.c.caps, self.c.lens = _lib.ptr(ann), _lib.ptr(caps), _lib.ptr(lens)
        if sampled is not None and any(sampled):
            self._sampled = (C.c_int32 * len(sampled))(*[1 if x else 0 for x in sampled])     # host array, kept alive
            self.c.sampled = C.cast(self._sampled, C.c_void_p)
        else:
            self._sampled = None
            self.c.sampled = None
        self.c.label_smoothing = float(label_smoothing)
        self.c.att_gamma = float(att_gamma)
        self.c.dropout_p, self.c.emb_dropout_p, self.c.dropout_seed = float(dropout[0]), float(dropout[1]), int(dropout[2])


def annotations_as_bld(ann, dtype):
    """[Bi,D,h,w] (any memory format) -> contiguous [Bi,L,D] of `dtype`.  Zero-copy when the encoder
    produced channels_last output in `dtype`."""
    Bi, D, h, w = ann.shape
    x = ann.permute(0, 2, 3, 1).reshape(Bi, h * w, D)
    if x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous()


def train_forward(pw, ann_bld, caps, lens, label_smoothing=0.0, att_gamma=1.0, exact=True, use_tc=False,
                  logits_f32=False, backward=True, keep_logits=False, buffers=None, sampled=None, dropout=(0.0, 0.0, 0)):
    """ann_bld [Bi,L,D] (pw.dtype, cuda); caps [Bi,ncap,T+1] or [B,T+1] int; lens [Bi,ncap] or [B].
    Runs sat_train_forward; returns the TrainBuffers (loss etc. in .t['out'])."""
    L_ = _lib.lib()
    dev = ann_bld.device
    Bi, L, D = ann_bld.shape
    caps2 = caps.reshape(-1, caps.shape[-1]).to(device=dev, dtype=torch.int32).contiguous()
    lens2 = lens.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
    B, caplen = caps2.shape
    dm = pw.dims
    assert D == dm["D"], "annotation width %d != encoder_dim %d" % (D, dm["D"])
    assert ann_bld.dtype == pw.dtype and ann_bld.is_contiguous()
    d = make_dims(B, Bi, L, D, dm["A"], dm["E"], dm["H"], dm["V"], caplen - 1, pw.dtype, exact, use_tc, pw.plain_output)
    if buffers is None:
        buffers = TrainBuffers(d, pw.dtype, dev, logits_f32=logits_f32, backward=backward, keep_logits=keep_logits)
    buffers.bind_inputs(ann_bld, caps2, lens2, label_smoothing, att_gamma, sampled, dropout)
    buffers.dims = d
    _lib.check(L_.sat_train_forward(C.byref(d), pw.ref(), C.byref(buffers.c), _lib.stream_ptr()), "sat_train_forward")
    return buffers


def _mm_tn(a, b):
    """a^T @ b with fp32 output (plain library GEMM: cuBLAS through torch).  a [M,N1], b [M,N2]."""
    if a.dtype != b.dtype:
        a, b = a.float(), b.float()
    if a.dtype == torch.float32:
        return a.t() @ b
    try:
        return torch.mm(a.t(), b, out_dtype=torch.float32)
    except TypeError:
        return (a.t() @ b).float()


def train_backward(pw, buf, grad_loss=None, pad_idx=0, weight_tying=False, dalpha_ext=None):
    """Runs sat_train_backward on the buffers of a finished train_forward, then reduces the saved
    per-(t,b) buffers into parameter gradients (reference names / shapes) with plain GEMMs.
    Returns (grads dict, d_ann [Bi,L,D])."""
    L_ = _lib.lib()
    t, d = buf.t, buf.dims
    B, Bi, ncap, L, D, A, E, H, V, T = d.B, d.Bi, d.ncap, d.L, d.D, d.A, d.E, d.H, d.V, d.T
    NH3 = A + D + 4 * H
    M = T * B
    if grad_loss is None:
        t["gscale"].fill_(1.0)
    else:
        t["gscale"].copy_(grad_loss.detach().reshape(1).to(torch.float32))
    if dalpha_ext is not None:
        t["dalpha_ext"] = dalpha_ext.to(torch.float32).contiguous()
        buf.c.dalpha_ext = _lib.ptr(t["dalpha_ext"])
    else:
        buf.c.dalpha_ext = None
    _lib.check(L_.sat_train_backward(C.byref(d), pw.ref(), C.byref(buf.c), _lib.stream_ptr()), "sat_train_backward")
    g = t["gscale"]
    G = {}
    dlog = t["dlogits"].reshape(M, V)
    Xo = t["Xo"].reshape(M, E)
    dWo = _mm_tn(dlog, Xo) * g
    G["output.output.bias"] = dlog.sum(0, dtype=torch.float32) * g
    dpre = t["dpre"].reshape(M, E)
    Hn = t["Hs"][1:].reshape(M, H)
    G["output.hidden.weight"] = _mm_tn(dpre, Hn)
    if not pw.plain_output:
        G["output.context.weight"] = _mm_tn(dpre, t["Z"].reshape(M, D))
    DY = t["DY"].reshape(M, NH3)
    dWh3 = _mm_tn(DY, t["Hs"][:T].reshape(M, H))
    G["attention.decoder_att.weight"] = dWh3[:A]
    G["beta.0.weight"] = dWh3[A:A + D]
    G["lstm.weight_hh_l0"] = deinterleave_gates(dWh3[A + D:])
    dy_sum = DY.sum(0, dtype=torch.float32)       # one pass over DY for both bias gradients
    G["beta.0.bias"] = dy_sum[A:A + D]
    dG = DY[:, A + D:]
    dWihe = _mm_tn(dG, t["Xe"].reshape(M, E))
    dWihz = _mm_tn(dG, t["GZ"].reshape(M, D))
    G["lstm.weight_ih_l0"] = deinterleave_gates(torch.cat([dWihe, dWihz], 1))
    db = deinterleave_gates(dy_sum[A + D:])
    G["lstm.bias_ih_l0"] = db
    G["lstm.bias_hh_l0"] = db.clone()
    G["attention.f_att.weight"] = t["dwf_part"].sum((0, 1)).reshape(1, A)
    ann = t["ann"].reshape(Bi * L, D)
    if ncap == 1 and d.use_tc and "dP16" in t and ann.dtype != torch.float32:
        dPm = t["dP16"].reshape(Bi * L, A)      # operand-dtype copy already written by the kernels
    else:
        dP = t["dP"]
        if ncap > 1:
            dP = dP.reshape(Bi, ncap, L, A).sum(1)
        dPm = dP.reshape(Bi * L, A)
        if ann.dtype != torch.float32:
            dPm = dPm.to(ann.dtype)
    G["attention.encoder_att.weight"] = _mm_tn(dPm, ann)
    dio = t["d_init_out"]
    G["init_lstm.init.weight"] = _mm_tn(dio, t["f1"].float())
    G["init_lstm.init.bias"] = dio.sum(0)
    G["init_lstm.factorize.weight"] = _mm_tn(t["df1"], t["meanv"].float())
    G["init_lstm.factorize.bias"] = t["df1"].sum(0)
    tok = t["tok"].reshape(M).long()            # words actually fed (ground truth or scheduled-sampling feedback)
    dEmb = torch.zeros(V, E, dtype=torch.float32, device=dlog.device)
    dEmb.index_add_(0, tok, t["dXe"].reshape(M, E))
    if pad_idx is not None:
        dEmb[pad_idx].zero_()                                   # nn.Embedding(padding_idx=<PAD>), model.py:162
    if weight_tying:
        dEmb = dEmb + dWo
    else:
        G["output.output.weight"] = dWo
    G["embedding.weight"] = dEmb
    d_ann = t["d_ann"]
    if ncap > 1:
        d_ann = d_ann.reshape(Bi, ncap, L, D).sum(1, dtype=torch.float32)
    else:
        d_ann = d_ann.reshape(Bi, L, D)
    return G, d_ann
