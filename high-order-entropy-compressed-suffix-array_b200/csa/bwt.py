"""Drop-in for the reference's csa/bwt.py on B200: ``bwt_transform(text, suffix_array)``
(csa/bwt.py:3-13) -- bwt[i] = text[SA[i]-1], wrapping to text[n-1] when SA[i] == 0.
Runs libhkcsa K2 (one gather kernel)."""
from hkcsa import views as _views


def bwt_transform(text, suffix_array):
    """text: str / bytes / uint8 tensor; suffix_array: list[int] / DeviceSequence / tensor.

    Returns str for str input (bytes for bytes, a uint8 CUDA tensor for tensor input).
    Like the reference: a suffix array shorter than the text raises IndexError, extra
    entries are ignored, an entry > len(text) raises IndexError (text[pos] out of range),
    and any entry <= 0 reads text[n-1].
    """
    import torch
    from hkcsa import engine
    smap = engine.SymbolMap(text) if isinstance(text, str) else None
    d_text = engine.to_device_u8(smap.encode(text) if smap is not None else text)
    n = d_text.numel()
    if len(suffix_array) < n:
        raise IndexError("list index out of range")
    d_sa = _views.as_device_i32(suffix_array, d_text.device)[:n].contiguous()
    if n:
        if int(d_sa.max().item()) > n:
            raise IndexError("string index out of range")
        d_sa = torch.clamp(d_sa, min=0)          # pos = SA[i]-1 < 0 -> n-1 (csa/bwt.py:9-10)
    out = engine.bwt(d_text, d_sa)
    if isinstance(text, str):
        return smap.decode(out.cpu().numpy().tobytes())
    if isinstance(text, (bytes, bytearray, memoryview)):
        return out.cpu().numpy().tobytes()
    return out
