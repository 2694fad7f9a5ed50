"""BGZF inflate on the GPU (strgpu_inflate_bgzf, `strling extract --gpu-inflate`).  Run with -m gpu.

The decoder's arithmetic is host/inflate_fast.hpp compiled as a device function and is covered on the CPU
(tests/test_host_cpu.py::test_inflate_decoder_against_zlib -- both the direct form kernel 1 runs and the command-stream form
kernels 2 and 3 run --, ::test_extract_staging_and_replay_without_the_scan with STRLING_DEBUG_STAGED_INFLATE); what these tests add
is the CUDA side (csrc/decode_kernels.cu), for each of its three kernels (STRGPU_INFLATE_KERNEL).  Every case runs in a child
process with a timeout, so a device fault cannot take the test session's CUDA context with it."""
import os
import subprocess
import sys
import textwrap

import pytest

from oracle import extract_oracle as eo
from strling_b200 import bamio
from strling_b200 import build as sb_build

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DIRECT = textwrap.dedent("""
    import sys, zlib
    import numpy as np
    sys.path.insert(0, %r)
    import strling_b200 as sb
    from strling_b200.binding import BGZF_BLOCK_DTYPE

    rng = np.random.default_rng(7)
    raws, comps = [], []
    for i in range(400):
        n = int(rng.integers(0, 65281)) if i >= 40 else i
        kind = i %% 5
        if kind == 0: raw = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif kind == 1: raw = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n).tobytes()
        elif kind == 2: raw = bytes((j %% (1 + i %% 300)) & 255 for j in range(n))
        elif kind == 3: raw = bytes(70 if v > 8 else v for v in rng.integers(0, 256, n))
        else: raw = rng.integers(0, 2 + i %% 250, n, dtype=np.uint8).tobytes()
        level, strategy = [0, 1, 6, 9][i %% 4], [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE][(i // 4) %% 4]
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        comps.append(co.compress(raw) + co.flush())
        raws.append(raw)
    blocks = np.zeros(len(raws), dtype=BGZF_BLOCK_DTYPE)
    comp = bytearray()
    out_off = 100          # the bytes before the first block (a carried partial record in the reader) must stay untouched
    for i, (r, c) in enumerate(zip(raws, comps)):
        blocks[i] = (len(comp), len(c), len(r), out_off)
        comp += c + bytes(8)   # stands in for the 8-byte BGZF footer behind every payload
        out_off += len(r)
    with sb.StrGpu(0) as g:
        out = g.inflate_bgzf(np.frombuffer(bytes(comp), dtype=np.uint8), blocks, out_off + 50)
        assert not out[:100].any() and not out[out_off:].any()
        assert out[100:out_off].tobytes() == b"".join(raws)
        bad = bytearray(comp)
        k = int(blocks["in_off"][300]) + int(blocks["csize"][300]) // 2
        bad[k] ^= 0x10
        bad[k + 1] ^= 0x01
        try:
            res = g.inflate_bgzf(np.frombuffer(bytes(bad), dtype=np.uint8), blocks, out_off + 50)
            same = res[100:out_off].tobytes() == b"".join(raws)
            assert not same, "a corrupted stream decoded to the original bytes"
        except sb.StrGpuError as e:
            assert e.args[0] == -8, e.args
        assert not out[out_off:].any()
    print("ok")
""")


# kernel 4 was written after the round's GPU time was spent and has not run on hardware: it may fail without failing the suite
KERNEL4 = pytest.param("4", marks=pytest.mark.xfail(strict=False, reason="inflate kernel 4 has not run on hardware yet (env-selected, experimental)"))


@pytest.mark.parametrize("kernel", ["1", "2", "3", KERNEL4])
def test_gpu_inflate_matches_zlib(kernel):
    """400 blocks through strgpu_inflate_bgzf: stored / fixed / dynamic DEFLATE blocks of every zlib level and strategy, sizes 0..65280,
    incompressible data (payload larger than kernel 3's staging buffer) -- byte-identical to the input of zlib's encoder, nothing
    written outside [first out_off, last out_off + isize), and a damaged block is refused (STRGPU_ERR_DATA) or at least decodes
    to something else without touching memory outside its range."""
    r = subprocess.run([sys.executable, "-c", DIRECT % ROOT], capture_output=True, text=True, timeout=240, env=dict(os.environ, STRGPU_INFLATE_KERNEL=kernel))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]


def test_extract_with_gpu_inflate_writes_the_same_bin(tmp_path):
    """`strling extract --gpu-inflate` (all three kernels) == `strling extract` == the oracle's .bin, byte for byte."""
    cli = sb_build.build_cli()
    targets = [("chr1", 3_000_000), ("chr2", 2_000_000)]
    loci = [(0, 400_000, 400_150, "CAG"), (0, 900_000, 900_090, "AAAG"), (1, 300_000, 300_060, "ATTCT")]
    recs = bamio.simulate_alignments(17, 20_000, targets, loci, str_pair_frac=0.05, unmapped_pairs=200, n_frac=0.02)
    hdr = bamio.sam_header(targets)
    bam = str(tmp_path / "g.bam")
    bamio.write_bam(bam, hdr, targets, recs)
    outs = []
    for k, extra in enumerate(([], ["--gpu-inflate"], ["--gpu-inflate"], ["--gpu-inflate"])):
        out = str(tmp_path / f"x{k}.bin")
        r = subprocess.run([cli, "extract", "--batch-reads", "8192", *extra, bam, out], capture_output=True, text=True, timeout=240,
                           env=dict(os.environ, STRGPU_INFLATE_KERNEL=str(max(k, 1))))
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(open(out, "rb").read())
    exp, cache, _ = eo.extract(recs, targets, hdr)
    assert len(cache) > 500
    for got in outs:
        assert got == exp
