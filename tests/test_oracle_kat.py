"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the hot path
(SURVEY.md section 4 / 8c).  Each test names the reference test it restates."""
import numpy as np
import pytest

from oracle import oracle as orc


def tr(position, split=orc.NONE, tid=1, repeat=b"ATG", **kw):
    return orc.make_tread(tid=tid, position=position, split=split, repeat=repeat, **kw)


def cat(*ts):
    return np.concatenate(ts)


# ---------------------------------------------------------------- tests/test_strling.nim
def test_monomer_repeat():  # test_strling.nim:46-66
    unit, rc = orc.get_repeat("A" * 150, 0.6)
    assert unit == b"A" and rc == 150


def test_triplet_repeat():  # test_strling.nim:68-89
    read = "TGC" * 50 + "T"
    assert len(read) == 151
    unit, rc = orc.get_repeat(read, 0.8)
    assert unit == b"CTG" and rc == 49


def test_unplaced_pair():  # test_strling.nim:91-107
    A = orc.make_tread(tid=0, position=222, repeat=b"AAAAAT", repeat_count=150, mapq=30)
    B = orc.make_tread(tid=0, position=222, repeat=b"AAAAAT", repeat_count=150, mapq=30)
    assert orc.unplaced_pair(A, B, 0.8, 20)
    A = orc.make_tread(tid=0, position=222, repeat=b"AAAAAT", repeat_count=150, mapq=16)
    B = orc.make_tread(tid=0, position=222, repeat=b"", repeat_count=0, mapq=16)
    assert orc.unplaced_pair(A, B, 0.8, 20)
    A = orc.make_tread(tid=0, position=222, repeat=b"", repeat_count=150, mapq=30)
    B = orc.make_tread(tid=0, position=222, repeat=b"", repeat_count=0, mapq=30)
    assert not orc.unplaced_pair(A, B, 0.8, 20)


# ---------------------------------------------------------------- tests/test_utils.nim
def test_reduce_repeat():  # test_utils.nim:36-64
    assert orc.reduce_repeat(b"CCC") == (b"C\0\0\0\0\0", 3)
    assert orc.reduce_repeat(b"AA") == (b"A\0\0\0\0\0", 2)
    assert orc.reduce_repeat(b"AAAAAA") == (b"A\0\0\0\0\0", 6)
    assert orc.reduce_repeat(b"CTC") == (b"CTC\0\0\0", 1)
    assert orc.reduce_repeat(b"CTCC") == (b"CTCC\0\0", 1)
    assert orc.reduce_repeat(b"CCCCCT") == (b"CCCCCT", 1)


def test_canonical_repeat():  # test_utils.nim:66-74
    assert orc.canonical_repeat(b"CCCTT") == b"AAGGG"


def test_median():  # utils.nim:139-146 (no reference vector; sanity of the restatement)
    f = np.zeros(4096, dtype=np.uint32)
    f[300] = 10
    f[400] = 10
    assert orc.median(f, 0.5) == 300
    assert orc.median(f, 0.98) == 400


# ---------------------------------------------------------------- tests/test_extract.nim
def test_adjust_by_clip():  # test_extract.nim:7-19
    A = orc.make_tread(tid=2, position=86914345, repeat=b"CCG", mapq=10, repeat_count=40, align_length=80)
    B = orc.make_tread(tid=16, position=17470852, split=orc.NONE_RIGHT, mapq=60, repeat_count=0, align_length=71)
    assert orc.adjust_by(A, B, 0.4, 20, 0, int(B["position"][0]))
    assert A["position"][0] == 17470852 + 71
    assert A["tid"][0] == 16 and A["split"][0] == orc.NONE and A["mapq"][0] == 60


# ---------------------------------------------------------------- appendix A.2 checked outputs + base-order evidence
@pytest.mark.parametrize(
    "read,p,unit,rc",
    [("CAG" * 50, 0.8, b"CAG", 50), ("ATTCT" * 30, 0.8, b"CTATT", 29), ("AC" * 75, 0.8, b"CA", 74)],
)
def test_scan_examples(read, p, unit, rc):
    assert orc.get_repeat(read, p) == (unit, rc)


def test_base_order_evidence():  # genome_strs.nim:204 : a get_repeat-produced unit `CACGAT`
    assert orc.get_repeat("ACGATC" * 16 + "ACGA", 0.8) == (b"CACGAT", 15)


def test_scan_edge_cases():
    assert orc.get_repeat("", 0.8) == (b"", 0)
    assert orc.get_repeat("A", 0.8) == (b"", 0)
    assert orc.get_repeat("N" * 21 + "A" * 129, 0.6) == (b"", 0)  # utils.nim:238
    assert orc.get_repeat("N" * 20 + "A" * 130, 0.6) == (b"A", 130)


# ---------------------------------------------------------------- tests/test_cluster.nim
def test_clustering():  # test_cluster.nim:38-55
    reads = cat(*[tr(p, repeat=b"AAAAAT") for p in (1, 1, 1, 200, 255)])
    cl = orc.cluster_bucket(reads, 125, 3)
    assert len(cl) == 1
    first, n, _, _ = cl[0]
    assert n == 4 and reads["position"][first] == 1 and reads["position"][first + n - 1] == 200


def test_bounds():  # test_cluster.nim:58-79
    reads = cat(tr(123), tr(123), *[tr(223, orc.LEFT)] * 4, *[tr(253, orc.RIGHT)] * 4, tr(283))
    b = orc.bounds_of(reads)
    assert (b["left"], b["right"], b["left_most"], b["right_most"]) == (223, 253, 123, 283)


def test_bounds_no_clips():  # test_cluster.nim:81-91
    b = orc.bounds_of(cat(tr(1), tr(2), tr(5)))
    assert (b["left"], b["right"]) == (2, 3)


def test_bounds_no_right():  # test_cluster.nim:93-105
    b = orc.bounds_of(cat(tr(1, orc.LEFT), tr(1, orc.LEFT), tr(2), tr(3), tr(5)))
    assert (b["left"], b["right"]) == (1, 2)


def test_bounds_no_left():  # test_cluster.nim:107-118
    b = orc.bounds_of(cat(tr(2), tr(2), tr(3, orc.RIGHT), tr(5)))
    assert (b["left"], b["right"]) == (3, 4)


def test_bounds_clip_filter():  # test_cluster.nim:120-138
    reads = cat(tr(100, orc.RIGHT), tr(123), tr(223), tr(223), tr(223), tr(253), tr(283))
    b = orc.bounds_of(reads, max_clip_dist=50)
    assert (b["left"], b["right"], b["left_most"], b["right_most"], b["center_mass"]) == (223, 224, 100, 283, 223)


def test_inverted_bounds():  # test_cluster.nim:188-196
    pos = (48086080, 48086101, 48086132, 48086164, 48086187, 48086281)
    b = orc.bounds_of(cat(*[tr(p, tid=20, repeat=b"TT") for p in pos]))
    assert b["left"] < b["right"]


def test_should_split_cluster():  # test_cluster.nim:200-228
    spec = [(370, orc.NONE), (391, orc.RIGHT), (391, orc.RIGHT), (391, orc.RIGHT), (403, orc.NONE), (503, orc.NONE),
            (850, orc.LEFT), (850, orc.LEFT), (850, orc.LEFT), (850, orc.LEFT), (880, orc.NONE)]
    reads = cat(*[tr(p, s, tid=0, repeat=b"") for p, s in spec])
    cl = orc.cluster_bucket(reads, 500, 1)
    assert len(cl) == 2
    (f1, n1, lm1, rm1), (f2, n2, lm2, rm2) = cl
    assert n1 == 6 and reads["position"][f1 + n1 - 1] == 503
    assert n2 == 5 and reads["position"][f2] == 850
    assert (rm1, lm2) == (620, 621)  # SURVEY appendix A.3 checked values


def test_inverted_bounds_again():  # test_cluster.nim:231-242
    spec = [(115977335, orc.NONE), (115977397, orc.NONE), (115977419, orc.NONE), (115977448, orc.LEFT),
            (115977585, orc.NONE), (115977598, orc.NONE)]
    b = orc.bounds_of(cat(*[tr(p, s, tid=11, repeat=b"") for p, s in spec]))
    assert b["left"] < b["right"]


def test_inverted_bounds_3():  # test_cluster.nim:244-252
    spec = [(92611809, orc.NONE), (92611833, orc.RIGHT), (92611833, orc.RIGHT), (92611921, orc.NONE), (92611939, orc.NONE)]
    b = orc.bounds_of(cat(*[tr(p, s, tid=10, repeat=b"") for p, s in spec]))
    assert b["left"] < b["right"]


def test_right_most_bug():  # test_cluster.nim:254-268
    spec = [(34847227, orc.LEFT), (34847227, orc.NONE), (34847883, orc.LEFT), (34847911, orc.NONE), (34847921, orc.LEFT),
            (34847921, orc.LEFT), (34847930, orc.NONE), (34848950, orc.LEFT), (34848950, orc.LEFT), (34848950, orc.LEFT)]
    b = orc.bounds_of(cat(*[tr(p, s, tid=5, repeat=b"") for p, s in spec]))
    assert b["left"] < b["right"]
    assert b["left_most"] <= b["left"] and b["right_most"] >= b["right"]


# ---------------------------------------------------------------- Nim stdlib emulation sanity (unpinned, self-consistency only)
def test_counttable_emulation():
    key, val, distinct = orc.counttable_largest([5, 7, 7, 9, 5, 7])
    assert (key, val, distinct) == (7, 3, 3)
    # growth path: 40 distinct keys force two enlargements; the unique maximum must survive re-insertion
    keys = list(range(1000, 1040)) + [1017, 1017]
    key, val, distinct = orc.counttable_largest(keys)
    assert (key, val, distinct) == (1017, 3, 40)


def test_cluster_all_merge_mode():
    # 6 reads of sample 0 clustered; merge mode requires >= min_support reads from one sample (merge.nim:18-25)
    a = cat(*[tr(p, tid=0, repeat=b"CAG", sample=0) for p in (100, 110, 120)],
            *[tr(p, tid=0, repeat=b"CAG", sample=1) for p in (105, 115)])
    b, _ = orc.cluster_all(a, 500, 3, merge_mode=True)
    assert len(b) == 1 and b[0]["n_total"] == 5
    b, _ = orc.cluster_all(a, 500, 4, merge_mode=True)
    assert len(b) == 0
    # unplaced bucket is reported per unit in call mode (call.nim:226-228)
    u = cat(*[tr(0, tid=-1, repeat=b"AAGGG") for _ in range(7)])
    b, unplaced = orc.cluster_all(cat(a, u), 500, 3, merge_mode=False)
    assert len(b) == 1 and unplaced == {b"AAGGG": 7}
