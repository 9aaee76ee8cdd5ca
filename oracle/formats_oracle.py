"""CPU oracle for the on-disk formats either side of the loader.  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this; the product never does.

* BGZF: the reference reads block-compressed files through htsjdk's BlockCompressedInputStream (sam-rtg-2.23.0.1.jar, third
  party, not under /root/reference) on top of java.util.zip.Inflater, i.e. zlib.  Python's `zlib` IS that library, so
  `bgzf_read` is the reference implementation of the inflate step ("kind: reference"); the member layout follows SAM spec
  4.1 / RFC 1952.  Call sites in the reference: src/com/rtg/tabix/TabixLineReader.java, BlockCompressedLineReader.java:58-133,
  BlockCompressedPositionReader.java.  Pinned on the block-compressed files the reference's tests read
  (tests/golden/formats_fixtures.json).
* tabix: `TabixIndexOracle` restates TabixIndexReader.java:63-146 and AbstractIndexReader.java:102-205,238-262; pinned on
  TabixIndexReaderTest.java:50-95 (known virtual offsets for snp_only.vcf.gz.tbi and readerWindow1.sam.gz.tbi).
  `tabix_build` (a small restatement of TabixIndexer's binning, for synthetic inputs only) is NOT pinned.
* SDF sequence data: `sdf_planes_write` restates BitwiseByteArray.set + dumpCompressedValues
  (src/com/rtg/util/bytecompression/BitwiseByteArray.java:304-340,426-439) / FileBitwiseOutputStream; `sdf_planes_read`
  restates FileBitwiseInputStream.get (src/com/rtg/reader/FileBitwiseInputStream.java:138-169).  The reference holds no
  SDF fixture on disk (its tests format one at run time), so this part is a writer/reader pair restated from the same
  source: parity unpinned beyond that.
"""
import struct
import zlib

import numpy as np

MAX_BLOCK_TEXT = 0xff00   # bgzip / htsjdk BlockCompressedStreamConstants.DEFAULT_UNCOMPRESSED_BLOCK_SIZE
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


# ----------------------------------------------------------------------------------------------------------- BGZF
def bgzf_members(data: bytes, start: int = 0):
    """Yields (header offset, payload offset, payload length, crc32, isize, member size) for the members from `start` on."""
    pos = start
    while pos < len(data):
        if len(data) - pos < 18 or data[pos:pos + 3] != b"\x1f\x8b\x08" or not data[pos + 3] & 4:
            raise ValueError("not a block compressed (BGZF) file: invalid GZIP header at offset %d" % pos)
        xlen = struct.unpack_from("<H", data, pos + 10)[0]
        bsize, p = None, pos + 12
        while p + 4 <= pos + 12 + xlen:
            si1, si2, slen = data[p], data[p + 1], struct.unpack_from("<H", data, p + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", data, p + 4)[0] + 1
            p += 4 + slen
        if bsize is None:
            raise ValueError("not a block compressed (BGZF) file: no BC field at offset %d" % pos)
        crc, isize = struct.unpack_from("<II", data, pos + bsize - 8)
        yield pos, pos + 12 + xlen, bsize - xlen - 20, crc, isize, bsize
        pos += bsize


def bgzf_read(data: bytes, voffset_begin: int = 0, voffset_end: int = -1, check_crc: bool = True) -> bytes:
    """The text between two virtual offsets, inflated member by member with zlib (raw DEFLATE, as Inflater(true) does)."""
    c_beg, u_beg = voffset_begin >> 16, voffset_begin & 0xffff
    c_end, u_end = (len(data), 0) if voffset_end < 0 else (voffset_end >> 16, voffset_end & 0xffff)
    out = []
    for hdr, off, n, crc, isize, _ in bgzf_members(data, c_beg):
        if hdr > c_end or (hdr == c_end and u_end == 0):
            break
        text = zlib.decompressobj(-15).decompress(data[off:off + n])
        if len(text) != isize:
            raise ValueError("BGZF block at offset %d: did not inflate the expected amount" % hdr)
        if check_crc and zlib.crc32(text) != crc:
            raise ValueError("BGZF block at offset %d: CRC32 mismatch" % hdr)
        if hdr == c_end:
            text = text[:u_end]
        out.append(text)
    whole = b"".join(out)
    return whole[u_beg:]


def bgzf_write(text: bytes, block_text: int = MAX_BLOCK_TEXT, level: int = 6, eof: bool = True, strategy=zlib.Z_DEFAULT_STRATEGY,
               cuts=None):
    """bgzip: `text` as BGZF members of at most `block_text` bytes of text each.  Returns (file bytes, list of (text offset,
    file offset) per member) so that callers can form virtual offsets.  `cuts`: explicit text offsets to start a new member at
    (besides the size limit)."""
    out, index, pos = [], [], 0
    size = 0
    cutset = sorted(set(cuts or []))
    ci = 0
    while pos < len(text):
        end = min(len(text), pos + block_text)
        while ci < len(cutset) and cutset[ci] <= pos:
            ci += 1
        if ci < len(cutset) and cutset[ci] < end:
            end = cutset[ci]
        chunk = text[pos:end]
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        payload = co.compress(chunk) + co.flush()
        if len(payload) + 26 > 65536:   # incompressible: stored
            co = zlib.compressobj(0, zlib.DEFLATED, -15)
            payload = co.compress(chunk) + co.flush()
        bsize = len(payload) + 26
        member = (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1) + payload +
                  struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        index.append((pos, size))
        out.append(member)
        size += len(member)
        pos = end
    if eof:
        out.append(EOF_BLOCK)
    return b"".join(out), index


def voffset(index, text_offset: int) -> int:
    """Virtual offset of a text offset, given bgzf_write's member index."""
    lo = 0
    for i, (t, _) in enumerate(index):
        if t <= text_offset:
            lo = i
    t, f = index[lo]
    return (f << 16) | (text_offset - t)


# ----------------------------------------------------------------------------------------------------------- tabix
class TabixIndexOracle:
    """TabixIndexReader.java:63-146 + AbstractIndexReader.java:102-205."""
    MAX_BIN = ((1 << 18) - 1) // 7 + 1
    NUM_BINS = MAX_BIN + 1
    MAX_COORD = 1 << 29

    def __init__(self, tbi: bytes):
        if tbi[:2] != b"\x1f\x8b" or len(tbi) < 18 or not tbi[3] & 4:
            raise ValueError("not a valid TABIX index. (index file is not block compressed)")
        buf = bgzf_read(tbi)
        if len(buf) < 36:
            raise ValueError("not a valid TABIX index. (index does not have a complete header)")
        if buf[:4] != b"TBI\x01":
            raise ValueError("not a valid TABIX index. (index does not have a valid header)")
        n_ref, fmt, seq_col, beg_col, end_col, meta, skip, name_len = struct.unpack_from("<8i", buf, 4)
        self.options = {"format": fmt, "seq_col": seq_col - 1, "beg_col": beg_col - 1, "end_col": end_col - 1, "meta": meta, "skip": skip}
        if seq_col - 1 < 0:
            raise ValueError("not a valid TABIX index. (invalid col_seq)")
        if beg_col - 1 < 0:
            raise ValueError("not a valid TABIX index. (invalid col_beg)")
        if len(buf) < 36 + name_len:
            raise ValueError("not a valid TABIX index. (sequence names truncated)")
        names = buf[36:36 + name_len].split(b"\0")[:-1]
        if len(names) > n_ref:
            raise ValueError("not a valid TABIX index. (more sequence names provided than expected)")
        if len(names) != n_ref:
            raise ValueError("not a valid TABIX index. (insufficient number of sequence names provided)")
        self.names = [n.decode() for n in names]
        self.lookup = {n: i for i, n in enumerate(self.names)}
        self.buf = buf
        self.bin_pos, self.lin_pos = [], []
        pos = 36 + name_len
        for _ in range(n_ref):
            self.bin_pos.append(pos)
            n_bins = struct.unpack_from("<i", buf, pos)[0]
            pos += 4
            for _ in range(n_bins):
                n_chunks = struct.unpack_from("<i", buf, pos + 4)[0]
                pos += 8 + n_chunks * 16
            self.lin_pos.append(pos)
            n_lin = struct.unpack_from("<i", buf, pos)[0]
            pos += 4 + n_lin * 8

    @staticmethod
    def reg2bins(beg, end):
        c_end = end - 1
        bins = [0]
        for base, shift in ((1, 26), (9, 23), (73, 20), (585, 17), (4681, 14)):
            bins.extend(range(base + (beg >> shift), base + (c_end >> shift) + 1))
        return bins

    def query(self, name, start=-1, end=-1):
        """getFilePointers for one interval; None when nothing is indexed for it."""
        seq = self.lookup.get(name)
        if seq is None:
            return None
        buf = self.buf
        pos = self.bin_pos[seq]
        n_bins = struct.unpack_from("<i", buf, pos)[0]
        pos += 4
        chunks = {}
        for _ in range(n_bins):
            bin_no, n_chunks = struct.unpack_from("<Ii", buf, pos)
            chunks[bin_no] = [struct.unpack_from("<QQ", buf, pos + 8 + 16 * k) for k in range(n_chunks)]
            pos += 8 + 16 * n_chunks
        n_lin = struct.unpack_from("<i", buf, pos)[0]
        linear = struct.unpack_from("<%dQ" % n_lin, buf, pos + 4) if n_lin else ()
        beg = 0 if start <= -1 else start
        fin = self.MAX_COORD if (end <= -1 or end == 0x7fffffff) else end
        if fin > self.MAX_COORD or beg > self.MAX_COORD:
            raise ValueError("coordinates greater than can be addressed by tabix/bam indexes")
        linear_beg = beg >> 14
        lo = hi = None
        for b in self.reg2bins(beg, fin):
            for cb, ce in chunks.get(b, ()):
                if b >= 4681 or (linear_beg < len(linear) and linear[linear_beg] < ce):
                    lo = cb if lo is None or cb < lo else lo
                    hi = ce if hi is None or hi < ce else hi
        return None if lo is None else (lo, hi)


def reg2bin(beg, end):
    """TabixIndexer / htslib: the smallest bin that holds [beg, end)."""
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def tabix_build(records, names):
    """A .tbi for VCF-like records [(sequence index, beg0, end0, voffset of the line, voffset after the line)] in file order
    (format 2, columns 1/2, meta '#').  One chunk per run of consecutive records of a bin; linear index per 16 kbp window."""
    body = struct.pack("<4sii", b"TBI\x01", len(names), 2) + struct.pack("<5i", 1, 2, 0, ord("#"), 0)
    nm = b"".join(n.encode() + b"\0" for n in names)
    body += struct.pack("<i", len(nm)) + nm
    for s in range(len(names)):
        bins, linear = {}, []
        last_bin = None
        for seq, beg, end, v0, v1 in records:
            if seq != s:
                continue
            b = reg2bin(beg, end)
            ch = bins.setdefault(b, [])
            if last_bin == b and ch and ch[-1][1] == v0:
                ch[-1][1] = v1
            else:
                ch.append([v0, v1])
            last_bin = b
            for w in range(beg >> 14, ((end - 1) >> 14) + 1):
                while len(linear) <= w:
                    linear.append(None)
                if linear[w] is None:
                    linear[w] = v0
        fill = 0
        for w in range(len(linear)):   # empty windows take the previous window's offset (htslib)
            if linear[w] is None:
                linear[w] = fill
            fill = linear[w]
        body += struct.pack("<i", len(bins))
        for b in sorted(bins):
            body += struct.pack("<Ii", b, len(bins[b]))
            for v0, v1 in bins[b]:
                body += struct.pack("<QQ", v0, v1)
        body += struct.pack("<i", len(linear)) + b"".join(struct.pack("<Q", v) for v in linear)
    return bgzf_write(body)[0]


# ----------------------------------------------------------------------------------------------------------- SDF
def sdf_planes_write(codes) -> bytes:
    """BitwiseByteArray(bits = 3).set + dumpCompressedValues: residue codes (0 = N, 1..4 = A C G T) as groups of three
    big-endian 64-bit bit planes per 64 residues, most significant bit of the code in the first long."""
    codes = np.asarray(codes, np.uint8)
    n = int(codes.size)
    groups = (n + 63) // 64 if n else 0
    padded = np.zeros(groups * 64, np.uint8)
    padded[:n] = codes
    out = np.zeros((groups, 3), np.uint64)
    for b in range(3):   # plane 0 = bit 2 of the code; residue i of a group is bit i of the long
        bits = (padded >> (2 - b)) & 1
        out[:, b] = np.packbits(bits, bitorder="little").view("<u8")
    return out.astype(">u8").tobytes()


def sdf_planes_read(file: bytes, first: int, length: int) -> np.ndarray:
    """FileBitwiseInputStream.get for `length` residues from residue `first` (the 3-bit DNA case, :151-169)."""
    longs = np.frombuffer(file, ">u8")
    idx = np.arange(first, first + length, dtype=np.int64)
    which = (idx >> 6) * 3
    bit = (idx & 63).astype(np.uint64)
    b0 = (longs[which] >> bit) & np.uint64(1)
    b1 = (longs[which + 1] >> bit) & np.uint64(1)
    b2 = (longs[which + 2] >> bit) & np.uint64(1)
    return ((b0 << np.uint64(2)) | (b1 << np.uint64(1)) | b2).astype(np.uint8)
