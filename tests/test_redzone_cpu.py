"""The guard-band allocator used by tests/test_redzone_gpu.py, exercised on CPU tensors."""
import torch


def test_redzone_allocator_cpu(monkeypatch):
    from sat_b200 import _redzone
    monkeypatch.delenv("SAT_REDZONE", raising=False)
    x = _redzone.empty((4, 3), torch.float32, "cpu")
    assert x.shape == (4, 3) and not _redzone.enabled()
    monkeypatch.setenv("SAT_REDZONE", "1")
    _redzone.violations()
    a = _redzone.empty((5, 7), torch.bfloat16, "cpu", "a")
    b = _redzone.empty((3,), torch.int32, "cpu", "b")
    assert a.shape == (5, 7) and a.dtype == torch.bfloat16 and a.is_contiguous()
    a.zero_()
    b.fill_(7)
    assert _redzone.violations(clear=False) == []
    torch.as_strided(b, (4,), (1,))[3] = 1          # one element past the end of b
    assert _redzone.violations() == ["b"]
    assert _redzone.violations() == []              # registrations were cleared
