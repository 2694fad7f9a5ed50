"""BGZF inflate on the GPU (strgpu_inflate_bgzf, `strling extract --gpu-inflate`).  Run with -m gpu.

The decoder's arithmetic is host/inflate_fast.hpp compiled as a device function and is covered on the CPU
(tests/test_host_cpu.py::test_inflate_decoder_against_zlib, ::test_extract_staging_and_replay_without_the_scan with
STRLING_DEBUG_STAGED_INFLATE); what these tests add is the CUDA side (csrc/decode_kernels.cu).  That kernel was written after the
round's GPU budget was spent and HAS NOT RUN ON HARDWARE yet, so each test runs it in a child process with a timeout and is
marked xfail(strict=False): a failure here reads as "the opt-in GPU inflate does not work yet", never as a regression of the
default path (`strling extract` without --gpu-inflate inflates on the host threads and is covered by tests/test_cli_gpu.py)."""
import os
import subprocess
import sys
import textwrap

import pytest

from oracle import extract_oracle as eo
from strling_b200 import bamio
from strling_b200 import build as sb_build

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="csrc/decode_kernels.cu has not run on hardware yet (opt-in path)")]
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


def test_gpu_inflate_matches_zlib():
    r = subprocess.run([sys.executable, "-c", DIRECT % ROOT], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]


def test_extract_with_gpu_inflate_writes_the_same_bin(tmp_path):
    cli = sb_build.build_cli()
    targets = [("chr1", 3_000_000), ("chr2", 2_000_000)]
    loci = [(0, 400_000, 400_150, "CAG"), (0, 900_000, 900_090, "AAAG"), (1, 300_000, 300_060, "ATTCT")]
    recs = bamio.simulate_alignments(17, 20_000, targets, loci, str_pair_frac=0.05, unmapped_pairs=200, n_frac=0.02)
    hdr = bamio.sam_header(targets)
    bam = str(tmp_path / "g.bam")
    bamio.write_bam(bam, hdr, targets, recs)
    outs = []
    for extra in ([], ["--gpu-inflate"]):
        out = str(tmp_path / ("a.bin" if not extra else "b.bin"))
        r = subprocess.run([cli, "extract", "--batch-reads", "8192", *extra, bam, out], capture_output=True, text=True, timeout=240)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(open(out, "rb").read())
    exp, cache, _ = eo.extract(recs, targets, hdr)
    assert len(cache) > 500
    assert outs[0] == exp
    assert outs[1] == exp
