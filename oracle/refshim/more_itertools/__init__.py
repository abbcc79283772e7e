"""Stand-in for more-itertools (import-time dependency of the reference's decoding.py only)."""
from itertools import chain, combinations


def powerset(iterable):
    s = list(iterable)
    return chain.from_iterable(combinations(s, r) for r in range(len(s) + 1))
