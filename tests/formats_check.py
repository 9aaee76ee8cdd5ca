"""Shared checkers of the on-disk format decoders (vce_bgzf.h: BGZF members, tabix index, SDF sequence data) against
oracle/formats_oracle.py (zlib itself for the inflate step; the reference's index reader restated; TabixIndexReaderTest's known
answers).  Used with the emulation in test_formats.py and with the CUDA library in test_gpu_parity.py."""
import base64
import json
import os
import zlib

import numpy as np

import formats_oracle as fo
import vcf_check

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fixtures():
    with open(os.path.join(ROOT, "tests", "golden", "formats_fixtures.json")) as f:
        return {k: base64.b64decode(v) for k, v in json.load(f).items()}


def is_bgzf(data):
    return len(data) >= 18 and data[:3] == b"\x1f\x8b\x08" and data[3] & 4 and data[12:14] == b"BC"


def check_reference_files(engine):
    """Every block-compressed file the reference's tests hold: whole text identical to zlib's; the plain-gzip ones refused."""
    n = 0
    for name, data in sorted(fixtures().items()):
        if is_bgzf(data):
            want = fo.bgzf_read(data)
            got, st = engine.bgzf_inflate(data)
            assert got == want, name
            assert st["text_bytes"] == len(want) and st["blocks"] == sum(1 for _ in fo.bgzf_members(data))
            assert engine.bgzf_text_size(data) == len(want)
            n += 1
        else:
            try:
                engine.bgzf_inflate(data)
            except ValueError as e:
                assert "block compressed" in str(e), (name, str(e))
            else:
                raise AssertionError("%s is plain gzip and must be refused" % name)
    return n


def check_tabix_known_answers(engine):
    """TabixIndexReaderTest.java:50-95."""
    fx = fixtures()
    t = engine.tabix_open(fx["snp_only.vcf.gz.tbi"])
    assert t.options == {"format": 2, "seq_col": 0, "beg_col": 1, "end_col": -1, "meta": ord("#"), "skip": 0}
    assert t.query("simulatedSequence19", 0, 5000) == (5480, 6025)
    assert t.query("simulatedSequence19", 500, 1000) == (5480, 6025)
    assert t.query("noSuchSequence", 0, 5000) is None
    t = engine.tabix_open(fx["readerWindow1.sam.gz.tbi"])
    assert t.options == {"format": 1, "seq_col": 2, "beg_col": 3, "end_col": -1, "meta": ord("@"), "skip": 0}
    assert t.query("simulatedSequence2", 32769, 33000) == ((44947 << 16) | 22268, (56124 << 16) | 50222)
    # testErrors: not block compressed / no complete header / no valid header
    for data, word in ((b"stuff", "block compressed"), (fo.bgzf_write(b"stuff")[0], "complete header"),
                       (fo.bgzf_write(b"stuffstdfafadfhjagjadshfgdjafhajdghadsjfhadsjfhjadsfhajkdsfhajdkshfajdkfahdjdsa")[0], "valid header")):
        try:
            engine.tabix_open(data)
        except ValueError as e:
            assert "TABIX" in str(e) and "not a valid" in str(e) and word in str(e), str(e)
        else:
            raise AssertionError("accepted a broken index")


def check_tabix_all_indexes(engine):
    """Every .tbi of the reference's tests: names, options and a grid of queries equal to the restated reader; and the text
    between the returned offsets equal to zlib's, where the data file is bundled too."""
    fx = fixtures()
    n = 0
    for name, data in sorted(fx.items()):
        if not name.endswith(".tbi"):
            continue
        want = fo.TabixIndexOracle(data)
        got = engine.tabix_open(data)
        assert got.names == want.names and got.options == want.options, name
        gz = fx.get(name[:-4])
        for seq in want.names + ["absent"]:
            for (b, e) in [(-1, -1), (0, 5000), (500, 1000), (0, 1), (16383, 16385), (32769, 33000), (100000, 200000), (0, 1 << 29),
                           (1 << 28, -1), (70000, 70001)]:
                w = want.query(seq, b, e)
                g = got.query(seq, b, e)
                assert g == w, (name, seq, b, e, g, w)
                n += 1
                if w is not None and gz is not None and is_bgzf(gz):
                    assert engine.bgzf_inflate(gz, w[0], w[1])[0] == fo.bgzf_read(gz, w[0], w[1]), (name, seq, b, e)
        got.close()
    return n


def random_text(rng, n, kind):
    if kind == "vcf":
        parts, pos = [], 1
        while sum(len(p) for p in parts) < n:
            pos += int(rng.integers(1, 900))
            parts.append("chr1\t%d\t.\t%s\t%s\t%d.%d\tPASS\t.\tGT:GQ\t%s:%d\n" % (
                pos, "ACGT"[int(rng.integers(4))], "ACGT"[int(rng.integers(4))], int(rng.integers(100)), int(rng.integers(10)),
                ["0/1", "1/1", "0|1", "1|0"][int(rng.integers(4))], int(rng.integers(99))))
        return "".join(parts).encode()[:n]
    if kind == "random":
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == "runs":       # long overlapping matches (distance < length) and 258-byte matches
        out = bytearray()
        while len(out) < n:
            unit = rng.integers(65, 70, int(rng.integers(1, 9)), dtype=np.uint8).tobytes()
            out += unit * int(rng.integers(1, 400))
        return bytes(out[:n])
    if kind == "skew":       # a skewed alphabet: long Huffman codes (beyond the 10-bit primary table)
        p = np.array([2.0 ** -i for i in range(1, 40)] + [2.0 ** -39])
        p /= p.sum()
        return rng.choice(len(p), n, p=p).astype(np.uint8).tobytes()
    if kind == "far":        # matches near the 32 KiB window limit
        base = rng.integers(0, 256, 30000, dtype=np.uint8).tobytes()
        return (base * (n // 30000 + 2))[:n]
    raise ValueError(kind)


def check_synthetic_streams(engine, seeds=range(6)):
    """DEFLATE variety the fixtures do not hold: stored members (incompressible data, level 0), fixed-Huffman members
    (Z_FIXED), dynamic members with long codes, overlapping and far matches, tiny and empty members, a missing EOF member,
    every compression level; whole files and random virtual-offset slices; each member's CRC32 checked on the way."""
    n = 0
    for seed in seeds:
        rng = np.random.default_rng(seed)
        for kind in ("vcf", "random", "runs", "skew", "far"):
            size = int(rng.choice([0, 1, 2, 100, 5000, 70000, 200000]))
            text = random_text(rng, size, kind)
            level = int(rng.choice([0, 1, 6, 9]))
            strategy = zlib.Z_FIXED if rng.random() < 0.25 else (zlib.Z_HUFFMAN_ONLY if rng.random() < 0.15 else zlib.Z_DEFAULT_STRATEGY)
            block = int(rng.choice([fo.MAX_BLOCK_TEXT, 65536 - 1000 if level else fo.MAX_BLOCK_TEXT, 4096, 17]))
            if size > 20000 and block == 17:
                block = 4096
            data, index = fo.bgzf_write(text, block, level, eof=bool(rng.random() < 0.8), strategy=strategy)
            got, st = engine.bgzf_inflate(data)
            assert got == text, (seed, kind, size, level, strategy, block)
            assert fo.bgzf_read(data) == text
            n += 1
            for _ in range(4):
                if not text:
                    break
                a = int(rng.integers(0, len(text)))
                b = int(rng.integers(a, len(text) + 1))
                va, vb = fo.voffset(index, a), fo.voffset(index, b)
                assert engine.bgzf_inflate(data, va, vb)[0] == text[a:b] == fo.bgzf_read(data, va, vb), (seed, kind, a, b)
                n += 1
    return n


def check_corrupt_members(engine, seeds=range(40)):
    """Damaged files end in ValueError (never in a crash or a hang), and whenever zlib still accepts the damaged member the
    decoder returns zlib's text (with check_crc off) -- the two must agree on accept / reject for a flipped payload bit."""
    rejected = accepted = 0
    for seed in seeds:
        rng = np.random.default_rng(1000 + seed)
        text = random_text(rng, int(rng.integers(200, 30000)), ["vcf", "runs", "skew", "random"][seed % 4])
        data = bytearray(fo.bgzf_write(text, level=int(rng.choice([1, 6, 9])))[0])
        _, off, ln, _, _, _ = next(fo.bgzf_members(bytes(data)))
        where = off + int(rng.integers(0, ln))
        data[where] ^= 1 << int(rng.integers(0, 8))
        data = bytes(data)
        try:
            want = fo.bgzf_read(data, check_crc=False)
        except (zlib.error, ValueError):
            want = None
        try:
            got = engine.bgzf_inflate(data, check_crc=False)[0]
        except ValueError:
            got = None
        if want is not None and got is not None:
            assert got == want, seed
        # zlib stops at the end-of-block code and ignores trailing payload bytes; so does the decoder.  Where zlib reports an
        # error the decoder must, too; where zlib accepts, the decoder may only disagree by rejecting an incomplete code set
        # zlib also rejects -- i.e. never.
        assert (got is None) == (want is None), (seed, got is None, want is None)
        # with the CRC check on, a changed text is always caught
        try:
            engine.bgzf_inflate(data, check_crc=True)
            assert want == text
        except ValueError:
            rejected += 1
        else:
            accepted += 1
    # truncated file, bad magic, oversized ISIZE
    good = fo.bgzf_write(b"hello world\n" * 100)[0]
    for bad in (good[:30], b"\x1f\x8b\x08\x00" + good[4:], good[:len(good) - 28 - 4] + b"\xff\xff\xff\x7f" + good[len(good) - 28:]):
        try:
            engine.bgzf_inflate(bad)
        except ValueError:
            rejected += 1
        else:
            raise AssertionError("accepted a broken file")
    return rejected, accepted


def synthetic_vcf_gz(rng, n_seq=3, records=400):
    """A sorted multi-sequence VCF, its BGZF form, a tabix index built from the line offsets, and the text."""
    from test_vcf_loader import fuzz_vcf
    names = ["chr%d" % (i + 1) for i in range(n_seq)]
    header = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2\n"
    lines = []
    for nm in names:
        body = fuzz_vcf(rng, n_records=records, template_len=4000)
        for ln in body.split("\n"):
            if not ln or ln.startswith("#"):
                continue
            f = ln.split("\t")
            if f[0] != "chr1":
                continue      # fuzz_vcf sprinkles records of another sequence; keep the file sorted by sequence
            f[0] = nm
            lines.append("\t".join(f))
    text = (header + "\n".join(lines) + "\n").encode()
    data, index = fo.bgzf_write(text, int(rng.choice([fo.MAX_BLOCK_TEXT, 3000, 777])), 6)
    recs, pos = [], len(header)
    for ln in lines:
        f = ln.split("\t")
        beg = int(f[1]) - 1
        end = beg + len(f[3].rstrip("\r"))
        nxt = pos + len(ln) + 1
        recs.append((names.index(f[0]), beg, max(end, beg + 1), fo.voffset(index, pos), fo.voffset(index, nxt)))
        pos = nxt
    return names, text, data, fo.tabix_build(recs, names)


def check_vcf_gz_chain(engine, seeds=range(4)):
    """.vcf.gz + .tbi -> per-sequence variants: index query, inflate and parse chained on the engine, against the reference
    loader restated in ref_host.py run on the plain text."""
    n = 0
    for seed in seeds:
        rng = np.random.default_rng(500 + seed)
        names, text, data, tbi = synthetic_vcf_gz(rng)
        idx = engine.tabix_open(tbi)
        want_idx = fo.TabixIndexOracle(tbi)
        assert idx.names == names
        for nm in names:
            q = idx.query(nm)
            assert q == want_idx.query(nm) and q is not None
            # the slice holds exactly that sequence's lines (plus nothing of the header)
            sl = fo.bgzf_read(data, q[0], q[1]).decode()
            assert sl and all(ln.startswith(nm + "\t") for ln in sl.split("\n") if ln), nm
            for sample, pass_only in ((0, True), (1, False)):
                want, skipped = vcf_check.expected_side(text.decode(), nm, 4000, sample, 2, pass_only, 1000)
                got, st = engine.vcf_parse_bgzf(data, q[0], q[1], nm, 4000, sample, 2, pass_only, 1000)
                # ids count the records READ for the sequence: the slice starts at the sequence's first record, as
                # VcfRecordTabixCallable's reader does, so they agree with the whole-text loader
                vcf_check.assert_same_side(got, want, "seed %d %s sample %d" % (seed, nm, sample))
                assert st["skipped"] == skipped and st["blocks"] >= 1
                n += 1
        # the whole file in one go (no index): every sequence's records are found by name
        got, _ = engine.vcf_parse_bgzf(data, 0, -1, names[-1], 4000, 0)
        want, _ = vcf_check.expected_side(text.decode(), names[-1], 4000, 0)
        vcf_check.assert_same_side(got, want, "whole file")
        idx.close()
    return n


def random_codes(rng, n):
    codes = rng.integers(1, 5, n).astype(np.uint8)
    for _ in range(int(rng.integers(0, 4))):   # runs of N
        a = int(rng.integers(0, max(1, n)))
        codes[a:a + int(rng.integers(1, 200))] = 0
    return codes


def check_sdf_decode(engine, seeds=range(8)):
    """Emulation: the SDF granule / mask bodies against FileBitwiseInputStream restated, and against packTemplate (the
    packing vce_template_register uploads)."""
    for seed in seeds:
        rng = np.random.default_rng(seed)
        lens = [int(x) for x in rng.integers(1, 5000, 4)] + [1, 15, 16, 17, 63, 64, 65, 511, 512, 513]
        seqs = [random_codes(rng, n) for n in lens]
        seqs.append(rng.integers(0, 8, 777).astype(np.uint8))   # codes 5..7 cannot come from the formatter; they must still round-trip
        allc = np.concatenate(seqs)
        file = fo.sdf_planes_write(allc)
        assert np.array_equal(fo.sdf_planes_read(file, 0, allc.size), allc)
        first = 0
        for codes in seqs:
            bases, packed, nmask = engine.sdf_decode(file, first, codes.size)
            assert np.array_equal(bases, fo.sdf_planes_read(file, first, codes.size)) and np.array_equal(bases, codes)
            p2, m2 = engine.pack_template(codes)
            assert np.array_equal(packed, p2) and np.array_equal(nmask, m2), (seed, first, codes.size)
            first += codes.size
