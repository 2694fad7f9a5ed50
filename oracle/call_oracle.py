"""CPU ORACLE (test infrastructure, not product code) for the rest of `strling call`:
spanning-read / spanning-pair evidence (collect.nim), the smoothed fragment distribution (spanning.nim), the
genotyper (genotyper.nim) and the call_main orchestration that writes -genotype.txt and the depth column of
-bounds.txt (call.nim:189-281).  Pure Python restatement for small cases; only tests/ may import it.

Parity status: the reference cannot be built here (see strling_oracle.h).  Pinned by the reference's own tests:
tests/test_genotyper.nim (spanning_read_est), tests/test_collect.nim (overlapping_read, spanning_fragment),
tests/test_utils.nim:10-12 (median_depth).  UNPINNED (no reference test, or behaviour of un-vendored code):
  * hts-nim `query(tid, start, stop)`: taken as "records with pos < stop and end > start in file order" (SAM spec / htslib);
  * Nim Table iteration order: `expected_spanners` is a float32 sum over `Table.values` (collect.nim:169-170) -- summed here in
    first-insertion order; the order of -genotype.txt lines (call.nim:268-278) -- emitted here by canonical repeat, then
    discovery order (file parity = sorted-line equality);
  * CountTable tie-breaks (most_frequent / largest, genotyper.nim:76-92): Nim 1.6 slot order (hashWangYi1, 64 slots).
"""
from __future__ import annotations

import math

import numpy as np

from . import extract_oracle as eo
from . import oracle as orc

SPANNING_FRAGMENT, SPANNING_READ, OVERLAPPING_READ = 0, 1, 2
_MASK64 = (1 << 64) - 1


# ------------------------------------------------------------------------------------------------ small numeric helpers
def cumulative(frag_dist) -> np.ndarray:  # spanning.nim:8-19 (float32 throughout)
    fd = np.asarray(frag_dist, dtype=np.uint32)
    res = np.zeros(4096, dtype=np.float32)
    for i in range(4096):
        acc = np.float32(0)
        for j in range(max(0, i - 11), min(i + 11, 4095) + 1):
            acc = np.float32(acc + np.float32(fd[j]))
        res[i] = acc
    res = np.add.accumulate(res, dtype=np.float32)   # math.cumsum: sequential float32 adds
    fmax = res[-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return (res / fmax).astype(np.float32)


def expected_spanning_probability(cd, start, stop, reverse, event_start, event_stop, min_spanning_bases=20) -> float:
    # spanning.nim:21-52
    if start < event_stop - min_spanning_bases:
        if reverse:
            return 0.0
        dist = event_start - start
        if dist < 0:
            return 0.0
        if dist + (event_stop - event_start) < min_spanning_bases:
            return 0.0
    else:
        if not reverse:
            return 0.0
        dist = stop - event_stop
        if dist < 0:
            return 0.0
        if dist + (event_stop - event_start) < min_spanning_bases:
            return 0.0
    dist += min_spanning_bases
    dist += event_stop - event_start
    if dist < 0 or dist > 4095:
        return 0.0
    return float(np.float32(1) - cd[dist])


def percentile(frag_dist, fragment_length: int) -> float:  # utils.nim:129-137
    total = int(np.asarray(frag_dist, dtype=np.uint64).sum()) & 0xFFFFFFFF
    s = 0
    for i in range(4096):
        s += int(frag_dist[i])
        if i >= fragment_length:
            break
    return float(s) / float(max(1, total))


def median_depth(D) -> int:  # utils.nim:148-158
    H = [0] * 1048
    for d in D:
        H[min(int(d), 1047)] += 1
    s = 0
    for i, h in enumerate(H):
        s += h
        if float(s) > float(len(D)) / 2.0:
            return i
    return 0


def _hash_wangyi1(x: int) -> int:  # Nim 1.6 hashes.nim
    def hi_xor_lo(a, b):
        p = (a & _MASK64) * (b & _MASK64)
        return ((p >> 64) ^ p) & _MASK64

    P0, P1, P58 = 0xA0761D6478BD642F, 0xE7037ED1A0B428DB, 0xEB44ACCAB455D165 ^ 8
    return hi_xor_lo(hi_xor_lo(P0, (x & _MASK64) ^ P1), P58)


def counttable_most_frequent(keys, cap0: int = 64):
    """Key with the largest count in a default-initialised Nim CountTable after `inc` of every key in order; ties go to the
    lowest slot (most_frequent's stable sort and `largest` agree on that).  Returns (key, count, n_distinct)."""
    cap = cap0
    slots_k = [0] * cap
    slots_v = [0] * cap
    counter = 0

    def raw_insert(sk, sv, c, k, v):
        h = _hash_wangyi1(k & _MASK64) & (c - 1)
        while sv[h] != 0:
            h = (h + 1) & (c - 1)
        sk[h], sv[h] = k, v

    for k in keys:
        h = _hash_wangyi1(k & _MASK64) & (cap - 1)
        found = False
        while slots_v[h] != 0:
            if slots_k[h] == k:
                slots_v[h] += 1
                found = True
                break
            h = (h + 1) & (cap - 1)
        if found:
            continue
        if cap * 2 < counter * 3 or cap - counter < 4:
            ncap = cap * 2
            nk, nv = [0] * ncap, [0] * ncap
            for i in range(cap):
                if slots_v[i] != 0:
                    raw_insert(nk, nv, ncap, slots_k[i], slots_v[i])
            slots_k, slots_v, cap = nk, nv, ncap
        raw_insert(slots_k, slots_v, cap, k, 1)
        counter += 1
    if counter == 0:
        return None, 0, 0
    mi = 0
    for h in range(1, cap):
        if slots_v[mi] < slots_v[h]:
            mi = h
    return slots_k[mi], slots_v[mi], counter


# ------------------------------------------------------------------------------------------------ collect.nim
_Q = set("MIS=X")
_R = set("MDN=X")


def find_read_position(a, position: int) -> int:  # collect.nim:50-72
    r_off, q_off = a.pos, 0
    for op, n in a.cigar:
        if r_off > position:
            return -1
        if op in _Q:
            q_off += n
        if op in _R:
            r_off += n
        if r_off < position:
            continue
        over = r_off - position
        if over > q_off:
            return -1
        if op not in _Q:
            return -1
        return q_off - over
    return -1


def _count_nonoverlap(s: str, sub: str) -> int:  # strutils.count(s, sub) with overlapping = false
    return s.count(sub) if sub else 0


def count_in_bounds(a, left: int, right: int, repeat: str) -> int:  # collect.nim:75-95
    if right < left:
        return 0
    dna = a.seq
    rl = find_read_position(a, left)
    rr = find_read_position(a, right)
    if rl >= 0 and rr < 0:
        rr = len(dna)
    if rl < 0 and rr < 0:
        return 0
    if rl < 0:
        rl = 0
    S = dna[rl:rr] if rr >= rl else ""
    res = _count_nonoverlap(S, repeat)
    if res < int(float(len(S)) * 0.7 / float(len(repeat))):
        res = 0
    return res


def _slop(left: int, right: int, repeat: str) -> int:
    slop = len(repeat) - 1
    if right - left < 5:
        slop += 5 - (right - left)
    return slop


def overlapping_read(a, tid: int, left: int, right: int, repeat: str):  # collect.nim:99-121 -> support dict or None
    stop = eo.aln_stop(a)
    if not (a.tid == tid and max(a.pos, left) <= min(stop, right)):   # cluster.nim:104-108
        return None
    s = dict(type=OVERLAPPING_READ, frag_len=0, frag_pct=0.0, rc=count_in_bounds(a, left, right, repeat) & 0xFF, ins=0, dele=0)
    slop = _slop(left, right, repeat)
    if a.pos < left - slop and stop > right + slop:
        s["type"] = SPANNING_READ
        for op, n in a.cigar:
            if op == "I":
                s["ins"] = (s["ins"] + (n & 0xFF)) & 0xFF
            if op == "D":
                s["dele"] = (s["dele"] + (n & 0xFF)) & 0xFF
    return s


def spanning_fragment(L, R, left: int, right: int, repeat: str, frag_dist):  # collect.nim:35-48 -> support dict or None
    assert L.pos <= R.pos
    slop = _slop(left, right, repeat)
    if L.pos < left - slop and eo.aln_stop(R) > right + slop:
        fl = max(1, abs(L.isize)) & 0xFFFFFFFF
        return dict(type=SPANNING_FRAGMENT, frag_len=fl, frag_pct=percentile(frag_dist, fl), rc=0, ins=0, dele=0)
    return None


def spanners(records, tid: int, left: int, right: int, repeat: str, window: int, frag_dist, cd=None, min_mapq: int = 20,
             max_size: int = 5000):
    """collect.nim:130-183 over a coordinate-sorted record list.  Returns (supports, median_depth, expected_spanners float32)."""
    if cd is None:
        cd = cumulative(frag_dist)
    wl, wr = left - window, right + window
    depths = [0] * (wr - wl)
    qbeg, qend = max(0, wl), wr
    support = []
    pairs = {}
    exp_by_q = {}
    for a in records:
        if a.tid != tid:
            continue
        stop = eo.aln_stop(a)
        if not (a.pos < qend and stop > qbeg):
            continue
        if a.flag & (0x100 | 0x800 | 0x400):
            continue
        if a.mapq < min_mapq:
            continue
        prob = expected_spanning_probability(cd, a.pos, stop, bool(a.flag & 0x10), left, right)
        if prob > 0:
            if a.qname in exp_by_q:
                exp_by_q[a.qname] = 0.5 * (exp_by_q[a.qname] + prob)
            else:
                exp_by_q[a.qname] = prob
        depths[max(0, a.pos - wl - 1)] += 1
        depths[min(len(depths) - 1, stop - wl - 1)] -= 1
        s = overlapping_read(a, tid, left, right, repeat)
        if s is not None:
            support.append(s)
        if a.tid != a.mate_tid:
            continue
        if abs(a.isize) > max_size:
            continue
        pairs.setdefault(a.qname, []).append(a)
        if len(pairs) > 20_000:
            return [], -1, np.float32(0)
    expected = np.float32(0)
    for v in exp_by_q.values():
        expected = np.float32(expected + np.float32(v))
    for q, pr in pairs.items():
        if len(pr) != 2:
            continue
        s = spanning_fragment(pr[0], pr[1], left, right, repeat, frag_dist)
        if s is not None:
            support.append(s)
    md = median_depth(np.cumsum(depths))
    return support, md, expected


# ------------------------------------------------------------------------------------------------ genotyper.nim
def spanning_read_est(supports):  # genotyper.nim:57-94 -> (allele1_bp, allele2_bp, allele1_ru, allele2_ru, supporting_reads)
    rcs, indels, n = [], [], 0
    for s in supports:
        if s["type"] == SPANNING_READ:
            rcs.append(s["rc"])
            indels.append(s["ins"] - s["dele"])
            n += 1

    def top2(keys):
        k1, _, distinct = counttable_most_frequent(keys)
        if distinct == 0:
            return math.nan, math.nan
        if distinct == 1:
            return float(k1), math.nan
        rest = [k for k in keys if k != k1]
        # second entry of the descending stable sort: the most frequent of the remaining keys, ties by slot of the FULL table
        full_order = _slot_order(keys)
        counts = {}
        for k in rest:
            counts[k] = counts.get(k, 0) + 1
        best = max(counts.values())
        k2 = [k for k in full_order if counts.get(k, 0) == best][0]
        return float(k1), float(k2)

    a1_ru, a2_ru = top2(rcs)
    a1_bp, a2_bp = top2(indels)
    return a1_bp, a2_bp, a1_ru, a2_ru, n


def _slot_order(keys, cap0: int = 64):
    """Distinct keys in slot order of the final CountTable."""
    cap = cap0
    sk, sv = [None] * cap, [0] * cap
    counter = 0
    for k in keys:
        h = _hash_wangyi1(k & _MASK64) & (cap - 1)
        found = False
        while sv[h] != 0:
            if sk[h] == k:
                sv[h] += 1
                found = True
                break
            h = (h + 1) & (cap - 1)
        if found:
            continue
        if cap * 2 < counter * 3 or cap - counter < 4:
            ncap = cap * 2
            nk, nv = [None] * ncap, [0] * ncap
            for i in range(cap):
                if sv[i] != 0:
                    hh = _hash_wangyi1(sk[i] & _MASK64) & (ncap - 1)
                    while nv[hh] != 0:
                        hh = (hh + 1) & (ncap - 1)
                    nk[hh], nv[hh] = sk[i], sv[i]
            sk, sv, cap = nk, nv, ncap
        h = _hash_wangyi1(k & _MASK64) & (cap - 1)
        while sv[h] != 0:
            h = (h + 1) & (cap - 1)
        sk[h], sv[h] = k, 1
        counter += 1
    return [sk[i] for i in range(cap) if sv[i] != 0]


def anchored_lm(sum_str_counts: int, depth: float) -> float:  # genotyper.nim:113-120
    if sum_str_counts == 0:
        return math.nan
    y = math.log2(float(sum_str_counts) / max(1.0, depth) + 1) * 0.7565329 + 4.3558142
    return math.pow(2, y)


def unplaced_est(unplaced_count: int, depth: float) -> float:  # genotyper.nim:131-136
    y = math.log2(float(unplaced_count) / depth + 1) * 0.7595562 + 8.9199168
    return math.pow(2, y)


def nim_float(x: float) -> str:
    """Nim 1.6 `$`(float): shortest round-trip digits ("%.17g"-style dragonbox output), with ".0" appended to integral values."""
    if math.isnan(x):
        return "nan"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    r = repr(float(x))
    return r


def fmt2(x: float) -> str:  # strformat "{x:.2f}"
    if math.isnan(x):
        return "nan"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    return "%.2f" % x


def genotype(chrom, left, right, n_left, n_right, repeat, tandems, supports, opts, depth: float):
    """genotyper.nim:142-196.  tandems: list of (repeat_count, split, qname).  Returns the Call as a dict."""
    c = dict(chrom=chrom, start=left, stop=right, repeat=repeat, allele1=0.0, allele2=0.0, anchored_reads=0, spanning_reads=0,
             spanning_pairs=0, expected_spanning_fragments=np.float32(0), oe_pct=np.float32(0), left_clips=n_left, right_clips=n_right,
             unplaced_reads=0, depth=depth, sum_str_counts=0, is_large=False)
    ru = len(repeat)
    if len(supports) == 0:
        c["allele1"] = math.nan
    else:
        a1_bp, _, _, _, n_span = spanning_read_est(supports)
        if not math.isnan(a1_bp):
            c["allele1"] = a1_bp / float(max(1, ru))
        c["spanning_reads"] = n_span
        c["spanning_pairs"] = sum(1 for s in supports if s["type"] == SPANNING_FRAGMENT)
    # is_large is evaluated while allele2 is still 0.0 (genotyper.nim:169)
    c["is_large"] = (n_left >= opts["min_clip"] and n_right >= opts["min_clip"] and ((n_left + n_right) & 0xFFFF) >= opts["min_clip_total"]
                     and len(tandems) >= opts["min_support"] and c["allele2"] > float(opts["median_fragment_length"]))
    s = sum(t[0] for t in tandems)
    c["sum_str_counts"] = s & 0xFFFFFFFF
    c["allele2"] = anchored_lm(s, depth) / float(max(1, ru))
    c["anchored_reads"] = len({t[2] for t in tandems if t[1] == orc.NONE})
    return c


def call_line(c) -> str:  # genotyper.nim:52-53
    return (f"{c['chrom']}\t{c['start']}\t{c['stop']}\t{c['repeat']}\t{fmt2(c['allele1'])}\t{fmt2(c['allele2'])}\t{c['anchored_reads']}\t"
            f"{c['spanning_reads']}\t{c['spanning_pairs']}\t{fmt2(float(c['expected_spanning_fragments']))}\t{fmt2(float(c['oe_pct']))}\t"
            f"{c['left_clips']}\t{c['right_clips']}\t{c['unplaced_reads']}\t{nim_float(c['depth'])}\t{c['sum_str_counts']}")


GT_HEADER = ("#chrom\tleft\tright\trepeatunit\tallele1_est\tallele2_est\tanchored_reads\tspanning_reads\tspanning_pairs\t"
             "expected_spanning_pairs\tspanning_pairs_pctl\tleft_clips\tright_clips\tunplaced_pairs\tdepth\tsum_str_counts")


def oe_ratio(c) -> np.float32:  # call.nim:31-34 (float32)
    obs = np.float32(c["spanning_pairs"])
    exp = np.float32(c["expected_spanning_fragments"])
    return np.float32(np.float32(np.float32(1) + obs - exp) / np.float32(exp + np.float32(1)))


def add_percentile(calls):  # call.nim:37-47
    oes = sorted(oe_ratio(c) for c in calls)
    arr = np.asarray(oes, dtype=np.float32)
    for c in calls:
        lb = int(np.searchsorted(arr, oe_ratio(c), side="left"))
        with np.errstate(divide="ignore", invalid="ignore"):
            c["oe_pct"] = np.float32(np.float32(lb) / np.float32(len(oes) - 1))


# ------------------------------------------------------------------------------------------------ call.nim:223-281 (discovery)
def sorted_tread_order(treads) -> np.ndarray:
    """Indices of the records in (tid, repeat, position) order, `.bin` order inside ties (call.nim:124-130)."""
    rep = np.frombuffer(treads["repeat"].tobytes(), dtype=np.uint8).reshape(-1, 6) if len(treads) else np.zeros((0, 6), dtype=np.uint8)
    keys = [treads["position"]] + [rep[:, j] for j in range(5, -1, -1)] + [treads["tid"]]
    return np.lexsort(keys)


def parse_bounds_lines(lines, targets):  # cluster.nim:143-169
    names = [t[0] for t in targets]
    out = []
    for l in lines:
        if l.startswith("#"):
            continue
        f = l.split("\t")
        assert len(f) == 11
        out.append(dict(tid=names.index(f[0]), left=int(f[1]), right=int(f[2]), repeat=f[3], name=f[4], left_most=int(f[5]),
                        right_most=int(f[6]), center_mass=int(f[7]), n_left=int(f[8]), n_right=int(f[9]), n_total=int(f[10])))
    return out


def parse_bed_lines(lines, targets, window):  # cluster.nim:111-141
    names = [t[0] for t in targets]
    out = []
    for l in lines:
        f = l.split()
        assert len(f) in (4, 5)
        tid = names.index(f[0])
        left, right = int(f[1]), int(f[2])
        out.append(dict(tid=tid, left=left, right=right, repeat=f[3], name=f[4] if len(f) == 5 else "", left_most=max(left - window, 0),
                        right_most=min(right + window, targets[tid][1]), center_mass=0, n_left=0, n_right=0, n_total=0))
    return out


def merge_loci_into_bounds(bounds, loci):  # call.nim:168-187
    loci = list(loci)
    for bound in bounds:
        for i, locus in enumerate(loci):
            if locus["tid"] == bound["tid"] and locus["repeat"] == bound["repeat"] and max(locus["left"], bound["left"]) <= min(locus["right"], bound["right"]):
                bound["name"], bound["left"], bound["right"] = locus["name"], locus["left"], locus["right"]
                loci[i] = loci[-1]      # seq.del(i)
                loci.pop()
                break
    return bounds + loci


def assign_reads_locus(locus, buckets, treads):  # callclusters.nim:14-50; buckets: {(tid, repeat bytes): [record indices by position]}
    key = (locus["tid"], locus["repeat"].encode())
    trs = buckets.get(key, [])
    lm = 0 if locus["left_most"] == 0 else locus["left_most"] - 1
    pos = [int(treads["position"][i]) for i in trs]
    import bisect

    li = bisect.bisect_left(pos, lm)
    ri = bisect.bisect_right(pos, locus["right_most"])
    res = []
    if trs:
        res = trs[li:ri]
        keep = trs[:li]
        if ri < len(trs) - 1:
            keep = keep + trs[ri + 1:]
        buckets[key] = keep
    locus["n_total"] = locus["n_left"] = locus["n_right"] = 0
    for i in res:
        locus["n_total"] += 1
        if int(treads["split"][i]) == orc.RIGHT:
            locus["n_right"] += 1
        elif int(treads["split"][i]) == orc.LEFT:
            locus["n_left"] += 1
    return res


def locus_bounds_line(b, targets) -> str:  # cluster.nim:262-266
    return (f"{targets[b['tid']][0]}\t{b['left']}\t{b['right']}\t{b['repeat']}\t{b['name']}\t{b['left_most']}\t{b['right_most']}\t{b['center_mass']}\t"
            f"{b['n_left']}\t{b['n_right']}\t{b['n_total']}")


def call(records, bin_data: bytes, min_support=5, min_clip=0, min_clip_total=0, min_mapq=40, bounds_lines=None, bed_lines=None):
    """call_main (call.nim:96-281).  Returns (genotype lines, bounds lines incl. depth, unplaced dict)."""
    u = eo.unpack_bin(bin_data)
    targets = eo.targets_from_header(u["header"])
    frag = eo.fragment_length_distribution(records)
    window = orc.median(frag, 0.99)
    med = orc.median(frag, 0.5)
    mcd = int(0.5 * float(med)) & 0xFFFF
    opts = dict(min_clip=min_clip, min_clip_total=min_clip_total, min_support=min_support, median_fragment_length=med)
    treads, qnames = u["treads"], u["qnames"]
    cd = cumulative(frag)
    by_tid = {}
    for a in records:
        if a.tid >= 0:
            by_tid.setdefault(a.tid, []).append(a)
    calls, bounds_out = [], []

    def genotype_one(tid, left, right, rep, n_left, n_right, idx, line):
        sup, md, expected = spanners(by_tid.get(tid, []), tid, left, right, rep, window, frag, cd, min_mapq)
        if len(sup) > 5000 or md == -1:
            return
        tandems = [(int(treads["repeat_count"][i]), int(treads["split"][i]), qnames[i]) for i in idx]
        c = genotype(targets[tid][0], left, right, n_left, n_right, rep, tandems, sup, opts, float(md))
        c["expected_spanning_fragments"] = expected
        c["canon"] = orc.canonical_repeat(rep)
        calls.append(c)
        bounds_out.append(line + "\t" + str(md))

    # -b / -l loci take their reads first (call.nim:158-218)
    keep = np.arange(len(treads))
    if bounds_lines or bed_lines:
        lb = merge_loci_into_bounds(parse_bounds_lines(bounds_lines or [], targets), parse_bed_lines(bed_lines or [], targets, window))
        order_all = sorted_tread_order(treads)
        buckets = {}
        for i in order_all:
            buckets.setdefault((int(treads["tid"][i]), bytes(treads["repeat"][i]).rstrip(b"\0")), []).append(int(i))
        for b in lb:
            idx = assign_reads_locus(b, buckets, treads)
            if b["right"] - b["left"] > 1000:
                continue
            genotype_one(b["tid"], b["left"], b["right"], b["repeat"], b["n_left"], b["n_right"], idx, locus_bounds_line(b, targets))
        keep = np.array(sorted(i for v in buckets.values() for i in v), dtype=np.int64)
    sub = treads[keep]
    b, unplaced = orc.cluster_all(sub, window, min_support, min_clip, min_clip_total, mcd, merge_mode=False)
    order = sorted_tread_order(sub)
    for x in b:
        tid, left, right = int(x["tid"]), int(x["left"]), int(x["right"])
        rep = bytes(x["repeat"]).rstrip(b"\0").decode()
        idx = [int(keep[j]) for j in order[int(x["first_read"]): int(x["first_read"]) + int(x["n_reads"])]]
        genotype_one(tid, left, right, rep, int(x["n_left"]), int(x["n_right"]), idx, eo.bounds_line(x, targets))
    add_percentile(calls)
    by_rep = {}
    for c in calls:
        by_rep.setdefault(c["canon"], []).append(c)
    lines = []
    for rep in sorted(by_rep):
        gts = by_rep[rep]
        large = [g for g in gts if g["is_large"]][:2]
        if len(large) == 1:
            n_un = unplaced.get(rep, 0)
            large[0]["unplaced_reads"] = n_un
            if n_un > 2:
                large[0]["allele2"] = unplaced_est(n_un, large[0]["depth"]) / float(len(large[0]["repeat"]))
        lines += [call_line(g) for g in gts]
    return lines, bounds_out, unplaced, (sub, b)
