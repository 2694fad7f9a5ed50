"""The pre-filter's bound checked against the oracle on the CPU: whenever get_repeat (utils.nim:236-271)
returns a unit, the two largest 2-mer occurrence counts s1 >= s2 of the segment satisfy
s1 + (k - 2) * s2 >= (int(L * p / k) + 1) * (k - 1) for some k = 2..6 -- so a segment that fails the test for every k can be
given the empty result without running the ladder.  (The CUDA kernel's use of the bound is covered by tests/test_scan_gpu.py.)"""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import oracle as orc
from strling_b200 import synth


def top2_2mer_counts(s: str):
    c = {}
    for i in range(len(s) - 1):
        c[s[i:i + 2]] = c.get(s[i:i + 2], 0) + 1
    v = sorted(c.values(), reverse=True) + [0, 0]
    return v[0], v[1]


def survives(s: str, p: float) -> bool:
    L = len(s)
    s1, s2 = top2_2mer_counts(s)
    return any(s1 + (k - 2) * s2 >= (int(float(L) * p / float(k)) + 1) * (k - 1) for k in range(2, 7))


def test_bound_holds_on_the_config2_mix():
    reads, cls, lclip, rclip = synth.make_reads(60_000, seed=3, mix=(0.5, 0.1, 0.2, 0.2), noise=0.03)
    segs, _ = synth.segments_for(reads, lclip, rclip, 160)
    flat, off, lens = synth.segment_ascii(reads, segs, 160)
    P = np.asarray([0.8, 0.73, 0.6])
    units, counts = orc.get_repeat_batch(flat, off, lens, P[segs["pclass"]])
    hits = np.nonzero(counts > 0)[0]
    assert len(hits) > 5000
    n_filtered = 0
    for i in range(len(segs)):
        s = bytes(flat[int(off[i]): int(off[i]) + int(lens[i])]).decode()
        ok = survives(s, float(P[segs["pclass"][i]]))
        if counts[i] > 0:
            assert ok, (s, units[i], counts[i])
        n_filtered += not ok
    assert n_filtered > 0.5 * len(segs)   # and the filter is worth having: most segments end there


@settings(max_examples=400, deadline=None)
@given(unit=st.text(alphabet="ACGT", min_size=1, max_size=6), copies=st.integers(0, 80), pad=st.text(alphabet="ACGT", max_size=60),
       phase=st.integers(0, 5), p=st.sampled_from([0.6, 0.73, 0.8, 0.9, 0.5]), noise=st.lists(st.integers(0, 159), max_size=6))
def test_bound_holds_on_adversarial_repeats(unit, copies, pad, phase, p, noise):
    s = (pad[: len(pad) // 2] + (unit * (copies + 1))[phase % len(unit):][: copies * len(unit)] + pad[len(pad) // 2:])[:160]
    s = list(s)
    for j in noise:
        if j < len(s):
            s[j] = "ACGT"[(j * 7 + len(s)) % 4]
    s = "".join(s)
    unit_found, count = orc.get_repeat(s, p)
    if count > 0:
        assert survives(s, p), (s, p, unit_found, count)


# ---- round 2: the kernel counts only nine of the sixteen 2-mer cells and three column sums and derives the rest from the margins
# of the 4 x 4 table (csrc/scan_kernels.cu, prefilter_top2).  A numpy model of exactly those formulas: the derived cells are the
# true counts where the kernel claims exactness and over-estimate by at most one elsewhere, so the filter stays sound.
_CODE = {"C": 0, "A": 1, "T": 2, "G": 3}


def derived_cells(s: str):
    L = len(s)
    b = [_CODE[ch] for ch in s]
    true = np.zeros((4, 4), dtype=int)
    for i in range(L - 1):
        true[b[i], b[i + 1]] += 1
    c = np.zeros((4, 4), dtype=int)
    c[:3, :3] = true[:3, :3]                                   # the nine counted cells
    f = [sum(1 for i in range(L - 1) if b[i + 1] == x) for x in range(3)]   # counted column sums: successor is x
    f.append(max(L - 1, 0) - sum(f))
    rest = f[3]
    for x in range(3):
        c[3, x] = f[x] - c[:3, x].sum()                        # exact
    for a in range(3):
        row = c[a, :3].sum()
        nplus = f[a] + (1 if L >= 2 and b[0] == a else 0)      # >= occurrences of a at positions 0 .. L-2
        d = nplus - row                                        # = true count + [last base == a]
        c[a, 3] = d
        rest = rest - d
    c[3, 3] = rest + 1                                         # = true count + 1 - [last base is one of the three counted]
    return true, c


@settings(max_examples=600, deadline=None)
@given(s=st.text(alphabet="ACGT", min_size=0, max_size=160))
def test_derived_cells_bound_the_true_counts(s):
    true, c = derived_cells(s)
    assert (c >= true).all() and (c <= true + 1).all(), (s, true, c)
    assert (c[:3, :3] == true[:3, :3]).all() and (c[3, :3] == true[3, :3]).all()    # exact where the kernel says so


def test_filter_with_derived_cells_keeps_every_segment_the_oracle_gives_a_unit():
    reads, cls, lclip, rclip = synth.make_reads(30_000, seed=5, mix=(0.4, 0.1, 0.25, 0.25), noise=0.03)
    segs, _ = synth.segments_for(reads, lclip, rclip, 160)
    flat, off, lens = synth.segment_ascii(reads, segs, 160)
    P = np.asarray([0.8, 0.73, 0.6])
    units, counts = orc.get_repeat_batch(flat, off, lens, P[segs["pclass"]])
    extra = 0
    for i in range(len(segs)):
        s = bytes(flat[int(off[i]): int(off[i]) + int(lens[i])]).decode()
        p = float(P[segs["pclass"][i]])
        _, c = derived_cells(s)
        v = np.sort(c.reshape(-1))[::-1]
        keep = any(v[0] + (k - 2) * v[1] >= (int(float(len(s)) * p / float(k)) + 1) * (k - 1) for k in range(2, 7))
        if counts[i] > 0:
            assert keep, (s, units[i], counts[i])
        extra += keep and not survives(s, p)
    assert extra < 0.002 * len(segs)     # the over-estimates let almost nothing more through than the exact counts do
