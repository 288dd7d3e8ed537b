"""Drop-in for the reference's csa/high_order_entropy.py on B200: ``calculate_high_order_entropy(text, k)``
(csa/high_order_entropy.py:4-32) -- the k-th order empirical entropy H_k in bits per symbol.

Same conventions as the reference: 0 for empty text or k < 0 (:6-7); H_0 from symbol frequencies (:10-15);
0 when n <= k (:17-18); contexts are the k-grams followed by a symbol and the weights are
context_total / n -- n, not n - k (:30).

GPU formulation: the (k+1)-grams are packed into 64-bit keys (one byte per symbol, so k <= 7) and sorted with
libhkcsa's onesweep radix sort; with c = multiplicity of a (k+1)-gram and T = multiplicity of its k-gram context,
    H_k = (1/n) * ( sum_T T*log2(T) - sum_c c*log2(c) ).
Run lengths and the two fp64 sums are tensor plumbing on the sorted keys.  Floating point: the reference adds the
context terms in dictionary order, we add sorted run terms -- results agree to ~1e-12 relative (tests use 1e-9).
"""
def calculate_high_order_entropy(text, k):
    if not text or k < 0:
        return 0
    import numpy as np
    import torch
    from hkcsa import engine

    n = len(text)
    if k == 0:
        hist = engine.byte_hist(engine.to_device_u8(text)).astype(np.float64)
        p = hist[hist > 0] / n
        return float(-(p * np.log2(p)).sum())
    if n <= k:
        return 0
    if k > 7:
        raise NotImplementedError("k-gram keys are packed into 64 bits: k <= 7 on the device path")
    d = engine.to_device_u8(text)
    m = n - k                                              # windows text[i : i+k+1], i in [0, n-k)
    keys = torch.zeros(m, dtype=torch.int64, device=d.device)
    for j in range(k + 1):                                 # plumbing: pack k+1 bytes, most significant first
        keys |= d[j:j + m].to(torch.int64) << (8 * (k - j))
    vals = torch.empty(m, dtype=torch.int32, device=d.device)
    engine.sort_pairs_u64(keys, vals, 8 * (k + 1))         # hand-written radix sort (K1's primitive)

    def run_log_sum(x):
        _, counts = torch.unique_consecutive(x, return_counts=True)
        c = counts.to(torch.float64)
        return float((c * torch.log2(c)).sum().item())

    s_grams = run_log_sum(keys)
    s_ctx = run_log_sum(keys >> 8)
    hk = (s_ctx - s_grams) / n
    return float(hk)
