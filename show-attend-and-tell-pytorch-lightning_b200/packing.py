"""Packs the reference-named decoder parameters (SURVEY.md §A.3 / model.py:158-199) into the
device layouts libsat_b200.so consumes (include/sat_b200.h: SatWeights)."""
import ctypes as C

import torch

from . import _lib

PARAM_NAMES = [
    "embedding.weight",
    "init_lstm.factorize.weight", "init_lstm.factorize.bias", "init_lstm.init.weight", "init_lstm.init.bias",
    "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
    "attention.encoder_att.weight", "attention.decoder_att.weight", "attention.f_att.weight",
    "beta.0.weight", "beta.0.bias",
    "output.hidden.weight", "output.context.weight", "output.output.weight", "output.output.bias",
]


def interleave_gates(w):
    """[4H, ...] in torch order (i|f|g|o blocks) -> row 4*j+g = row g*H+j."""
    H = w.shape[0] // 4
    return w.reshape(4, H, *w.shape[1:]).transpose(0, 1).reshape(w.shape).contiguous()


def deinterleave_gates(w):
    """inverse of interleave_gates."""
    H = w.shape[0] // 4
    return w.reshape(H, 4, *w.shape[1:]).transpose(0, 1).reshape(w.shape).contiguous()


class PackedWeights:
    """Device copies of the decoder weights in kernel layout.  `W` maps reference names to tensors
    (any device/dtype); `dtype` is the kernel operand dtype."""

    def __init__(self, W, dtype=torch.float32, device="cuda", backward=True):
        self.dtype = dtype
        dev = torch.device(device)
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32)
        emb = f32(W["embedding.weight"])
        wih, whh = f32(W["lstm.weight_ih_l0"]), f32(W["lstm.weight_hh_l0"])
        wa, wh = f32(W["attention.encoder_att.weight"]), f32(W["attention.decoder_att.weight"])
        wb, bb = f32(W["beta.0.weight"]), f32(W["beta.0.bias"])
        who = f32(W["output.hidden.weight"])
        # DeepOutput(deep=False) has no context projection (model.py:120-121): pack zeros and flag the plain epilogues
        self.plain_output = W.get("output.context.weight", None) is None
        wzo = torch.zeros(who.shape[0], wa.shape[1], device=dev) if self.plain_output else f32(W["output.context.weight"])
        wo = f32(W["output.output.weight"])
        bo = W.get("output.output.bias", None)
        V, E = emb.shape
        H = whh.shape[1]
        D = wa.shape[1]
        A = wa.shape[0]
        self.dims = dict(V=V, E=E, H=H, D=D, A=A)
        s = lambda t: t.to(dtype).contiguous()
        self.t = t = {}
        t["Wa"] = s(wa)
        whcat = torch.cat([wh, wb, interleave_gates(whh), who], 0)
        t["Whcat"] = s(whcat)
        t["bhcat"] = torch.cat([torch.zeros(A, device=dev), bb, torch.zeros(4 * H + E, device=dev)]).contiguous()
        t["Wihz"] = s(interleave_gates(wih[:, E:]))
        t["Wihe"] = s(interleave_gates(wih[:, :E]))
        t["bg"] = interleave_gates(f32(W["lstm.bias_ih_l0"]) + f32(W["lstm.bias_hh_l0"]))
        whozo = torch.cat([who, wzo], 1)
        t["Whozo"] = s(whozo)
        t["Wo"] = s(wo)
        t["bo"] = f32(bo).contiguous() if bo is not None else None
        t["wf"] = f32(W["attention.f_att.weight"]).reshape(-1).contiguous()
        t["Emb"] = s(emb)
        t["Wfact"] = s(f32(W["init_lstm.factorize.weight"]))
        t["bfact"] = f32(W["init_lstm.factorize.bias"]).contiguous()
        t["Winit"] = s(f32(W["init_lstm.init.weight"]))
        t["binit"] = f32(W["init_lstm.init.bias"]).contiguous()
        if backward:
            NH3 = A + D + 4 * H
            t["WoT"] = s(wo.t())
            t["WhozoT"] = s(whozo.t())
            t["WihzT"] = s(interleave_gates(wih[:, E:]).t())
            t["WiheT"] = s(interleave_gates(wih[:, :E]).t())
            t["WhcatT"] = s(whcat[:NH3].t())
            t["WaT"] = s(wa.t())
            t["WinitT"] = s(f32(W["init_lstm.init.weight"]).t())
            t["WfactT"] = s(f32(W["init_lstm.factorize.weight"]).t())
        self.c = _lib.SatWeights()
        for name, _ in _lib.SatWeights._fields_:
            setattr(self.c, name, _lib.ptr(t.get(name)))

    def ref(self):
        return C.byref(self.c)
