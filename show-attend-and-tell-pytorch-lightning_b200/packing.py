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

_LAYER_PARAMS = ("lstm.weight_ih_l%d", "lstm.weight_hh_l%d", "lstm.bias_ih_l%d", "lstm.bias_hh_l%d")
_LAYER_FIELDS = ("w_ih_l", "w_hh_l", "b_ih_l", "b_hh_l")       # SatMasterWeights / SatParamGrads arrays, index l - 1


def param_names(layers=1):
    """reference parameter names of a decoder with `layers` stacked LSTM layers (decoder_layers, model.py:175-180):
    PARAM_NAMES followed by the four nn.LSTM parameters of every layer l >= 1."""
    return PARAM_NAMES + [n % l for l in range(1, int(layers)) for n in _LAYER_PARAMS]


def layers_of(names_or_count):
    """number of LSTM layers of a parameter list in param_names() order (or of a dict keyed by the reference names)"""
    if isinstance(names_or_count, dict):
        n = 1
        while ("lstm.weight_ih_l%d" % n) in names_or_count and names_or_count["lstm.weight_ih_l%d" % n] is not None:
            n += 1
        return n
    return 1 + (int(names_or_count) - len(PARAM_NAMES)) // 4


def interleave_gates(w):
    """[4H, ...] in torch order (i|f|g|o blocks) -> row 4*j+g = row g*H+j."""
    H = w.shape[0] // 4
    return w.reshape(4, H, *w.shape[1:]).transpose(0, 1).reshape(w.shape).contiguous()


def deinterleave_gates(w):
    """inverse of interleave_gates."""
    H = w.shape[0] // 4
    return w.reshape(H, 4, *w.shape[1:]).transpose(0, 1).reshape(w.shape).contiguous()


class PackedWeights:
    """Device copies of the decoder weights in kernel layout (SatWeights), produced by ONE kernel (sat_pack_weights) from
    the fp32 master parameters.  `W` maps reference names to tensors; `dtype` is the kernel operand dtype.  The buffers
    are persistent: call repack(W) after an optimizer step instead of building a new object."""

    _SRC = (("embedding", "embedding.weight"), ("fact_w", "init_lstm.factorize.weight"), ("fact_b", "init_lstm.factorize.bias"),
            ("init_w", "init_lstm.init.weight"), ("init_b", "init_lstm.init.bias"), ("w_ih", "lstm.weight_ih_l0"),
            ("w_hh", "lstm.weight_hh_l0"), ("b_ih", "lstm.bias_ih_l0"), ("b_hh", "lstm.bias_hh_l0"),
            ("enc_att", "attention.encoder_att.weight"), ("dec_att", "attention.decoder_att.weight"),
            ("f_att", "attention.f_att.weight"), ("beta_w", "beta.0.weight"), ("beta_b", "beta.0.bias"),
            ("out_hidden", "output.hidden.weight"), ("out_context", "output.context.weight"),
            ("out_w", "output.output.weight"), ("out_b", "output.output.bias"))

    def __init__(self, W, dtype=torch.float32, device="cuda", backward=True):
        self.dtype = dtype
        dev = self.device = torch.device(device)
        V0, E0 = W["embedding.weight"].shape
        H0 = W["lstm.weight_hh_l0"].shape[1]
        A0, D0 = W["attention.encoder_att.weight"].shape
        nl = self.layers = layers_of(W)
        if nl > _lib.SAT_MAX_LAYERS:
            raise _lib.SatError("decoder_layers=%d: the kernels support up to %d stacked LSTM layers" % (nl, _lib.SAT_MAX_LAYERS))
        # The kernels work on storage dims that are multiples of 8 (16-byte vectors, TMA rows).  A module with other sizes
        # (V = words above min_count + 4, 100/300-d GloVe embeddings ...) is zero-padded here: zero weights keep the padded
        # lanes exactly zero through forward and backward, the padded vocabulary entries get a bias of -inf.
        up8 = lambda n: (int(n) + 7) // 8 * 8
        V, E, H, D, A = up8(V0), up8(E0), up8(H0), up8(D0), up8(A0)
        self.dims = dict(V=V, E=E, H=H, D=D, A=A)
        self.dims0 = dict(V=int(V0), E=int(E0), H=int(H0), D=int(D0), A=int(A0))
        self.padded = self.dims != self.dims0
        # DeepOutput(deep=False) has no context projection (model.py:120-121): its slots stay zero, plain epilogues
        self.plain_output = W.get("output.context.weight", None) is None
        NH3, NH4 = A + D + 4 * H, A + D + 4 * H + E
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        f = torch.float32
        t = self.t = {}
        t["Wa"] = z((A, D), dtype)
        t["Whcat"] = z((NH4, H), dtype)
        t["bhcat"] = z((NH4,), f)
        t["Wihz"] = z((4 * H, D), dtype)
        t["Wihe"] = z((4 * H, E), dtype)
        t["bg"] = z((4 * H,), f)
        t["Whozo"] = z((E, H + D), dtype)
        t["Wo"] = z((V, E), dtype)
        self.has_out_bias = W.get("output.output.bias", None) is not None
        t["bo"] = z((V,), f) if (self.has_out_bias or V != V0) else None
        if t["bo"] is not None and V != V0:
            t["bo"][V0:] = float("-inf")           # padded words never win a max, add nothing to a soft-max
        t["wf"] = z((A,), f)
        t["Emb"] = z((V, E), dtype)
        t["Wfact"] = z((E, D), dtype)
        t["bfact"] = z((E,), f)
        t["Winit"] = z((2 * nl * H, E), dtype)
        t["binit"] = z((2 * nl * H,), f)
        for l in range(1, nl):
            t["Wl%d" % l] = z((4 * H, 2 * H), dtype)
            t["bgl%d" % l] = z((4 * H,), f)
        if backward:
            t["WoT"] = z((E, V), dtype)
            t["WhozoT"] = z((H + D, E), dtype)
            t["WihzT"] = z((D, 4 * H), dtype)
            t["WiheT"] = z((E, 4 * H), dtype)
            t["WhcatT"] = z((H, NH3), dtype)
            t["WaT"] = z((D, A), dtype)
            t["WinitT"] = z((E, 2 * nl * H), dtype)
            t["WfactT"] = z((D, E), dtype)
            for l in range(1, nl):
                t["WlT%d" % l] = z((2 * H, 4 * H), dtype)
        self.c = _lib.SatWeights()
        for name, typ in _lib.SatWeights._fields_:
            if typ is C.c_void_p:
                setattr(self.c, name, _lib.ptr(t.get(name)))
            else:                                       # per-layer arrays (index l - 1)
                arr = getattr(self.c, name)
                for l in range(1, _lib.SAT_MAX_LAYERS):
                    arr[l - 1] = _lib.ptr(t.get("%s%d" % (name, l)))
        self._d = _lib.SatDims()
        self._d.B = self._d.Bi = self._d.ncap = 1
        self._d.L, self._d.D, self._d.A, self._d.E, self._d.H, self._d.V, self._d.T = 1, D, A, E, H, V, 1
        self._d.D0, self._d.A0, self._d.E0, self._d.H0, self._d.V0 = int(D0), int(A0), int(E0), int(H0), int(V0)
        self._d.dtype = _lib.dtype_code(dtype)
        self._d.layers = nl
        self.repack(W)

    def repack(self, W):
        """(re)fill every packed buffer from the current master parameters: one kernel launch."""
        m = _lib.SatMasterWeights()
        keep = []

        def src(p):
            x = p.detach()
            if x.device != self.device or x.dtype != torch.float32 or not x.is_contiguous():
                x = x.to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(x)
            return x.data_ptr()

        for field, name in self._SRC:
            p = W.get(name, None)
            setattr(m, field, None if p is None else src(p))
        for l in range(1, self.layers):
            for field, name in zip(_LAYER_FIELDS, _LAYER_PARAMS):
                getattr(m, field)[l - 1] = src(W[name % l])
        self._keep = keep            # sources must outlive the (asynchronous) kernel
        _lib.check(_lib.lib().sat_pack_weights(C.byref(self._d), C.byref(m), C.byref(self.c), _lib.stream_ptr()), "sat_pack_weights")
        return self

    def ref(self):
        return C.byref(self.c)
