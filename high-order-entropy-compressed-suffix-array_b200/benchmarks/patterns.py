"""Pattern sampling with the semantics of the reference's tests/test_patterns.py:3-9: for every requested length one
substring of the text at a uniformly random start (the length is clipped to the text)."""
import random


def generate_random_patterns(text, pattern_lengths, rng=random):
    out = []
    for want in pattern_lengths:
        m = min(want, len(text))
        at = rng.randint(0, len(text) - m)
        out.append(text[at:at + m])
    return out
