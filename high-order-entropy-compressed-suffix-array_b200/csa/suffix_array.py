"""Drop-in for the reference's csa/suffix_array.py on B200.

Same names and argument meaning as the reference: ``build_suffix_array(text)``
(csa/suffix_array.py:131-134) and ``ksa(T)`` (:46-129); plus ``optimized_ksa``,
which the reference's csa/csa.py:3 imports but never defines.

All three run the same CUDA builder (libhkcsa K1: prefix doubling over packed
64-bit keys, onesweep LSD radix sort).  The reference's ``ksa`` is a DC3 sketch
that raises TypeError whenever its recursion is needed; wherever
it returns, it returns build_suffix_array's answer, so the alias is a superset.

Deviation: the reference runs ``ksa("banana")`` and prints at import time
(:136-138).  Here the module attributes ``text`` / ``suffix_array`` are computed
on first access (importing must not need a GPU) and the print only happens with
HKCSA_DEMO_PRINT=1.
"""
import os

from hkcsa import views as _views


def build_suffix_array(text):
    """Suffix array of `text`: indices of all suffixes in code-point order, a proper prefix first.

    text: str (latin-1 range, one byte per code point: utils/data_loader.py:4), or bytes /
    uint8 array / uint8 CUDA tensor.  Returns list[int] (a lazy DeviceSequence with list
    semantics above hkcsa.views.MATERIALIZE_MAX entries).  "" -> [].
    """
    from hkcsa import engine
    if isinstance(text, str):
        text = engine.SymbolMap(text).encode(text)      # code points above 255: order-preserving re-coding to bytes
    d_text = engine.to_device_u8(text)
    return _views.int_sequence(engine.suffix_array(d_text))


def ksa(T):
    """The reference's DC3 entry point; same result as build_suffix_array (see module docstring)."""
    return build_suffix_array(T)


optimized_ksa = ksa


def __getattr__(name):
    # import-time demo of the reference (:136-138), evaluated lazily
    if name == "text":
        return "banana"
    if name == "suffix_array":
        sa = ksa("banana")
        if os.environ.get("HKCSA_DEMO_PRINT") == "1":
            print(f"Suffix Array using KSA: {sa}")
        return sa
    raise AttributeError(name)
