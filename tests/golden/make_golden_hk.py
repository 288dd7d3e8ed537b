#!/usr/bin/env python
"""Freeze H_k values of the REFERENCE ITSELF (csa/high_order_entropy.py:4-32, imported unmodified) for the texts
of golden_ref.npz and a few literals, incl. a text with code points above 255.

Run in the authoring container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_hk.py

Writes tests/golden/golden_hk.json: {case: {"text_latin1" | "text": ..., "hk": {k: value}}}; values are Python
floats printed with repr (round-trip exact).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HKCSA_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from csa.high_order_entropy import calculate_high_order_entropy  # noqa: E402

ORDERS = [-1, 0, 1, 2, 3, 5, 8, 12, 20]


def main():
    arr = np.load(os.path.join(HERE, "golden_ref.npz"))
    names = sorted({k.split("/")[0] for k in arr.files if k.endswith("/text")})
    out = {}
    for name in names:
        text = arr[f"{name}/text"].tobytes().decode("latin-1")
        out[name] = {"hk": {str(k): calculate_high_order_entropy(text, k) for k in ORDERS}}
    extra = {"unicode_greek": "αβγαβγδαβα" * 7 + "ω",
             "unicode_mixed": "naïve café — 北京 naïve café — 北京 coöperate",
             "k_equals_n": "abcabc",
             "mississippi_x50": "mississippi$" * 50}
    for name, text in extra.items():
        out["lit/" + name] = {"text": text, "hk": {str(k): calculate_high_order_entropy(text, k) for k in ORDERS + [6]}}
    with open(os.path.join(HERE, "golden_hk.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(f"{len(out)} cases x {len(ORDERS)} orders -> golden_hk.json")


if __name__ == "__main__":
    main()
