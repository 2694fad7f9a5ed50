"""CPU models of two reformulations the round-2 ladder kernels rely on (csrc/scan_kernels.cu), checked against the literal rule
of the reference.  (The kernels themselves are compared with the oracle on the GPU: tests/test_scan_gpu.py.)"""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st


def seq_inc_leader(codes):
    """Seq.inc / argmax (utils.nim:192-197): the first class to reach the final maximum leads (strict >)."""
    cnt, best, leader = {}, 0, None
    for c in codes:
        cnt[c] = cnt.get(c, 0) + 1
        if cnt[c] > best:
            best, leader = cnt[c], c
    return best, leader


def sorted_sweep_leader(ranks, n_classes=700):
    """lane_count_sorted: keys rank << 5 | window, padded to 32 with classes of their own, sorted; one sweep keeps
    max((run length so far * 32 + 31 - window) << 16 | key)."""
    w = len(ranks)
    keys = sorted([(r << 5) | i for i, r in enumerate(ranks)] + [((n_classes + j) << 5) | 31 for j in range(w, 32)])
    prev, cur, best = None, 0, 0
    for e in keys:
        cur = cur + 1 if prev is not None and (e ^ prev) < 32 else 1
        best = max(best, ((cur * 32 + ((e & 31) ^ 31)) << 16) | e)
        prev = e
    if w == 0:
        return 0, None
    return best >> 21, (best & 0xFFFF) >> 5


@settings(max_examples=2000, deadline=None)
@given(ranks=st.lists(st.integers(0, 699), min_size=0, max_size=26), few=st.integers(1, 6))
def test_sorted_sweep_equals_seq_inc(ranks, few):
    for data in (ranks, [r % few for r in ranks]):       # many classes / a few classes with ties
        assert sorted_sweep_leader(data) == seq_inc_leader(data), data


def test_merge_exchange_network_and_bitonic_merge_sort_32_keys():
    # the comparator schedule of lane_count_sorted: Batcher's merge exchange on 16 keys (both packed halves at once), then
    # low key i against high key 15 - i and four half-cleaner stages
    comps = []
    p = 1
    while p < 16:
        k = p
        while k >= 1:
            j = k % p
            while j <= 15 - k:
                for i in range(0, min(k - 1, 15 - j - k) + 1):
                    if (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        comps.append((i + j, i + j + k))
                j += 2 * k
            k //= 2
        p *= 2
    assert len(comps) == 63
    for bits in range(1 << 16):                           # 0-1 principle
        a = [(bits >> i) & 1 for i in range(16)]
        for i, j in comps:
            if a[i] > a[j]:
                a[i], a[j] = a[j], a[i]
        assert all(a[i] <= a[i + 1] for i in range(15))
    rng = np.random.default_rng(1)
    for _ in range(3000):
        x = [int(v) for v in rng.integers(0, 40, size=32)]
        lo, hi = x[:16], x[16:]
        for arr in (lo, hi):
            for i, j in comps:
                if arr[i] > arr[j]:
                    arr[i], arr[j] = arr[j], arr[i]
        for i in range(16):
            a, b = lo[i], hi[15 - i]
            lo[i], hi[15 - i] = min(a, b), max(a, b)
        for arr in (lo, hi):
            s = 8
            while s >= 1:
                for i in range(16):
                    if (i & s) == 0 and arr[i] > arr[i + s]:
                        arr[i], arr[i + s] = arr[i + s], arr[i]
                s //= 2
        assert lo + hi == sorted(x)
