"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/strgpu.h
declares, refuses to compute without a GPU (no CPU fallback), and its host packers agree with numpy."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import strling_b200 as sb
from strling_b200 import build as sb_build
from strling_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sb_build.build_lib()
    return sb.load_library()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "strgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(strgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libstrgpu.so does not export {n}"


def test_nim_shim_declares_every_entry_point_with_the_headers_arity():
    # strling_b200/nim/strgpu.nim cannot be compiled here (no nim): at least keep it in step with include/strgpu.h -- every
    # entry point declared, with as many parameters as the C prototype has
    hdr = open(os.path.join(ROOT, "include", "strgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    nim = open(os.path.join(ROOT, "strling_b200", "nim", "strgpu.nim")).read()
    protos = dict(re.findall(r"\b(strgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S))
    assert len(protos) >= 30
    for name, args in protos.items():
        n_c = 0 if args.strip() in ("", "void") else args.count(",") + 1
        m = re.search(r"proc " + name + r"\*\((.*?)\)(?::|\s*$)", nim, flags=re.S | re.M)
        assert m, f"strgpu.nim does not declare {name}"
        params = m.group(1)
        n_nim = 0
        for group in [g for g in params.split(";")] if ";" in params else [params]:
            for decl in re.split(r",(?![^\[]*\])", group):
                if decl.strip():
                    n_nim += 1
        assert n_nim == n_c, f"{name}: {n_c} parameters in strgpu.h, {n_nim} in strgpu.nim"


def test_struct_layouts():
    assert sb.SEGMENT_DTYPE.itemsize == 8 and sb.REPEAT_DTYPE.itemsize == 8
    assert sb.SEGMENT_DTYPE.fields["len"][1] == 4 and sb.REPEAT_DTYPE.fields["repeat_count"][1] == 6


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sb.StrGpuError) as e:
        sb.StrGpu(0)
    assert e.value.status == -3


def test_pack_ascii_matches_numpy(lib):
    reads, _, _, _ = synth.make_reads(300, seed=5, length=150, n_frac=0.2)
    seq2_np, nmask_np, stride = synth.pack_matrix(reads)
    seq2, nmask, segs, n_bases = sb.pack_reads([bytes(r) for r in reads])
    assert n_bases == 300 * stride
    assert np.array_equal(seq2[: n_bases // 4], seq2_np[: n_bases // 4])
    assert nmask is not None and nmask[1] is None and np.array_equal(nmask[0][: n_bases // 32], nmask_np[: n_bases // 32])
    assert np.array_equal(segs["base_off"], np.arange(300) * stride)


def test_pack_bam4_matches_ascii(lib):
    rng = np.random.default_rng(3)
    alphabet = b"=ACMGRSVTWYHKDBN"
    for length in (0, 1, 2, 3, 4, 5, 7, 8, 150, 151):
        nib = rng.integers(0, 16, size=length).astype(np.uint8)
        ascii_seq = bytes(alphabet[i] for i in nib)
        bam = np.zeros((length + 1) // 2 + 1, dtype=np.uint8)
        for i, v in enumerate(nib):
            bam[i // 2] |= v << (4 if i % 2 == 0 else 0)
        nb = 160
        a2 = np.zeros(lib.strgpu_seq2_bytes(nb), dtype=np.uint8)
        b2 = np.zeros_like(a2)
        am = np.zeros(lib.strgpu_nmask_bytes(nb) // 4, dtype=np.uint32)
        bm, ax, bx = np.zeros_like(am), np.zeros_like(am), np.zeros_like(am)
        ka = lib.strgpu_pack_ascii(ascii_seq, length, a2.ctypes.data, am.ctypes.data, ax.ctypes.data, 0)
        kb = lib.strgpu_pack_bam4(bam.ctypes.data, length, b2.ctypes.data, bm.ctypes.data, bx.ctypes.data, 0)
        assert ka == kb == sum(c not in b"ACGT" for c in ascii_seq)
        assert np.array_equal(a2, b2) and np.array_equal(am, bm) and np.array_equal(ax, bx)
        # nmask: every non-ACGT base; xmask: those that are not the literal 'N' (utils.nim:238 counts 'N' only)
        bits = lambda m: [(int(m[i >> 5]) >> (i & 31)) & 1 for i in range(length)]
        assert bits(am) == [int(c not in b"ACGT") for c in ascii_seq]
        assert bits(ax) == [int(c not in b"ACGTN") for c in ascii_seq]


def test_pack_reads_and_synth_agree_on_iupac_planes(lib):
    reads, _, _, _ = synth.make_reads(200, seed=6, length=150, n_frac=0.2, iupac_frac=0.2)
    seq2_np, masks_np, stride = synth.pack_matrix(reads)
    seq2, masks, segs, n_bases = sb.pack_reads([bytes(r) for r in reads])
    assert isinstance(masks_np, sb.Masks) and masks[1] is not None
    assert np.array_equal(masks[0][: n_bases // 32], masks_np[0][: n_bases // 32])
    assert np.array_equal(masks[1][: n_bases // 32], masks_np[1][: n_bases // 32])
