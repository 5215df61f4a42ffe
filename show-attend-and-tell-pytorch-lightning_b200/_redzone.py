"""Guard bands around device buffers (test aid; compute-sanitizer is not available on the GPU pool).

With SAT_REDZONE=1 every work buffer of the training / decode drivers is carved out of a slightly larger allocation whose
256 bytes before and after the payload are filled with a sentinel; `violations()` reports every buffer whose guard bands
were overwritten by a kernel since the buffers were made.  Without the variable `empty()` is torch.empty."""
import os

import torch

GUARD = 256
SENTINEL = 0xA5
_live = []          # (name, raw uint8 tensor, payload bytes)


def enabled():
    return os.environ.get("SAT_REDZONE", "0") == "1"


def empty(shape, dtype, device, name="?"):
    if not enabled():
        return torch.empty(shape, dtype=dtype, device=device)
    n = 1
    for x in shape:
        n *= int(x)
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    pad = (-nbytes) % 16
    raw = torch.full((GUARD + nbytes + pad + GUARD,), SENTINEL, dtype=torch.uint8, device=device)
    _live.append((name, raw, nbytes))
    return raw[GUARD:GUARD + nbytes].view(dtype).reshape(shape)


def violations(clear=True):
    """names of buffers whose guard bands no longer hold the sentinel"""
    bad = []
    for name, raw, nbytes in _live:
        head, tail = raw[:GUARD], raw[GUARD + nbytes:]
        if bool((head != SENTINEL).any()) or bool((tail != SENTINEL).any()):
            bad.append(name)
    if clear:
        _live.clear()
    return bad
