"""Drop-in for the reference's csa/high_order_entropy.py on B200: ``calculate_high_order_entropy(text, k)``
(csa/high_order_entropy.py:4-32) -- the k-th order empirical entropy H_k in bits per symbol.

Same conventions as the reference: 0 for empty text or k < 0 (:6-7); H_0 from symbol frequencies (:10-15);
0 when n <= k (:17-18); contexts are the k-grams followed by a symbol and the weights are
context_total / n -- n, not n - k (:30).

GPU formulation (csrc/entropy.cu): with c = multiplicity of a (k+1)-gram and T = multiplicity of its k-gram context,
    H_k = (1/n) * ( sum_T T*log2(T) - sum_c c*log2(c) ),
and the windows sharing a (k+1)-gram are a run of adjacent suffixes in the suffix array.  So the text's suffix array
(libhkcsa K1) is built once and one kernel pass per k flags the run heads, a scan numbers them and two fp64 sums are
reduced in a fixed order -- any k, no k-gram keys.  Floating point: the reference adds the context terms in
dictionary order, this adds run terms in suffix order: results agree to ~1e-12 relative (the tests use 1e-9).
"""


def calculate_high_order_entropy(text, k):
    if not text or k < 0:
        return 0
    from hkcsa import engine
    n = len(text)
    if k > 0 and n <= k:
        return 0
    if isinstance(text, str):
        text = engine.SymbolMap(text).encode(text)          # an order-preserving re-coding leaves H_k unchanged
    d = engine.to_device_u8(text)
    if k == 0:
        return engine.entropy_from_sa(d, None, 0)["hk"]
    return engine.entropy_from_sa(d, engine.suffix_array(d), k)["hk"]


def entropy_profile(text, orders=(0, 1, 2, 3, 4, 5)):
    """{k: H_k} for several orders from ONE suffix array."""
    from hkcsa import engine
    if not text:
        return {k: 0 for k in orders}
    if isinstance(text, str):
        text = engine.SymbolMap(text).encode(text)
    d = engine.to_device_u8(text)
    sa = engine.suffix_array(d) if any(k > 0 for k in orders) else None
    return {k: engine.entropy_from_sa(d, sa, k)["hk"] for k in orders}
