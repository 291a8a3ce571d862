"""Integer/sort utilities (oracle; test infrastructure only).

Restates src/sorting_tricks.jl.  Lists of tuples are plain Python lists or 2-D
integer numpy arrays; these helpers are setup-time only, never hot.
"""
import numpy as np


def sort_bitonic(t):
    """src/sorting_tricks.jl:1-29 -- sorting network for 1..4 integers."""
    return tuple(sorted(t))


def radix_sort(v, key=lambda x: x):
    """src/sorting_tricks.jl:44-76 -- LSD counting sort over the tuple fields.

    LSD counting sort is a *stable* lexicographic sort, which is exactly what
    Python's ``sorted`` is; stability fixes the owner order inside every
    interface cell (ascending element index)."""
    return sorted(v, key=key)


def binary_search(v, x, lo, hi):
    """src/sorting_tricks.jl:84-96 -- first index of x in v[lo:hi] (inclusive, 0-based)."""
    lo -= 1
    hi += 1
    while lo < hi - 1:
        m = (lo + hi) >> 1
        if v[m] < x:
            lo = m
        else:
            hi = m
    return hi


def remove_duplicates(v):
    """src/sorting_tricks.jl:109-125 -- unique of a sorted list."""
    out = []
    for x in v:
        if not out or out[-1] != x:
            out.append(x)
    return out


def remove_singletons(v, key=lambda x: x):
    """src/sorting_tricks.jl:130-154 -- drop values occurring exactly once (sorted input)."""
    out = []
    i = 0
    n = len(v)
    while i < n:
        j = i
        while j < n and key(v[j]) == key(v[i]):
            j += 1
        if j - i > 1:
            out.extend(v[i:j])
        i = j
    return out


def left_minus_right(lhs, rhs):
    """src/sorting_tricks.jl:160-189 -- sorted lhs minus sorted rhs."""
    out = []
    fast = 0
    idx = 0
    while fast < len(lhs) and idx < len(rhs):
        if lhs[fast] < rhs[idx]:
            out.append(lhs[fast])
            fast += 1
        elif lhs[fast] == rhs[idx]:
            fast += 1
            idx += 1
        else:
            idx += 1
    out.extend(lhs[fast:])
    return out


def complement(nodes, n):
    """src/sorting_tricks.jl:197-217 -- sorted (0..n-1) minus sorted ``nodes`` (0-based)."""
    mask = np.ones(n, dtype=bool)
    mask[np.asarray(nodes, dtype=np.int64)] = False
    return np.nonzero(mask)[0].astype(np.int64)


def remove_repeated_pairs(v, key=lambda x: x):
    """src/sorting_tricks.jl:222-248 -- drop every adjacent equal pair (sorted input)."""
    out = []
    n = len(v)
    fast = 0
    while fast < n - 1:
        if key(v[fast]) != key(v[fast + 1]):
            out.append(v[fast])
            fast += 1
        else:
            fast += 2
    if fast == n - 1:
        out.append(v[fast])
    return out
