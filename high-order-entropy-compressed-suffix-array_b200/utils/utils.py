"""Drop-in for the reference's utils/utils.py: ``build_count`` (:16-24), ``build_occ`` (:26-32)
and the ``time_function`` decorator (:4-14)."""
import time

from hkcsa import views as _views


def time_function(func):
    """Decorator returning (result, seconds) -- utils/utils.py:4-14."""
    def wrapper(*args, **kwargs):
        start_time = time.time()
        result = func(*args, **kwargs)
        end_time = time.time()
        return result, end_time - start_time
    return wrapper


def build_count(text):
    """C[c] = number of symbols in `text` with a smaller code point, for every symbol present,
    keys in sorted order (utils/utils.py:16-24).  One byte-histogram kernel."""
    from hkcsa import engine
    smap = engine.SymbolMap(text) if isinstance(text, str) else None
    hist = engine.byte_hist(engine.to_device_u8(smap.encode(text) if smap is not None else text))
    count, total = {}, 0
    for b in range(256):
        if hist[b]:
            count[smap.symbol(b) if smap is not None else chr(b)] = total
            total += int(hist[b])
    return count


def build_occ(bwt):
    """occ[c][i] = occurrences of c in bwt[0:i], i in [0, n], for every symbol of the BWT
    (utils/utils.py:26-32).  Answered by rank queries on the device wavelet tree; a dict of
    lists when (n+1)*sigma is small, otherwise a lazy mapping with the same indexing."""
    from hkcsa import engine
    wt = bwt if isinstance(bwt, engine.DeviceWaveletTree) else engine.DeviceWaveletTree(engine.to_device_u8(bwt))
    return _views.occ_mapping(wt)
