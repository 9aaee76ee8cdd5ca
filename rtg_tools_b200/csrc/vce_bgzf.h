// vce_bgzf.h -- SURVEY.md 8(f) rank 4: the on-disk formats either side of the VCF loader, decoded on the device.
//
//   BGZF   block-compressed VCF (`bgzip`): the reference reads it through htsjdk's BlockCompressedInputStream inside
//          sam-rtg-2.23.0.1.jar (third party, not under /root/reference; call sites src/com/rtg/tabix/TabixLineReader.java,
//          BlockCompressedLineReader.java:58-133, BlockCompressedPositionReader.java).  Restated from the published formats:
//          SAM spec 4.1 (a BGZF file is a series of gzip members of at most 64 KiB with a "BC" extra field carrying the
//          member size), RFC 1952 (member header / CRC32 / ISIZE trailer) and RFC 1951 (DEFLATE: stored, fixed and dynamic
//          Huffman blocks).  One BGZF block per warp: every lane runs the same bit reader and Huffman decode (table reads
//          are broadcasts, no divergence), literals are stored by lane 0, LZ77 matches are copied by all lanes.  Table entries are
//          pre-decoded (base value, number of extra bits, kind), so that a symbol costs one look-up and one bit-field extract.
//   tabix  AbstractIndexReader.getFilePointers (src/com/rtg/tabix/AbstractIndexReader.java:102-205, reg2bins :238-262) and the
//          TabixIndexReader constructor (TabixIndexReader.java:63-146) as plain host code over the inflated index: the virtual
//          offsets [begin, end) that hold a region's records.  The index itself is BGZF and goes through the same kernel.
//   SDF    sequence data of a formatted reference: BitwiseByteArray.dumpCompressedValues
//          (src/com/rtg/util/bytecompression/BitwiseByteArray.java:426-439) writes, per 64 residues, three big-endian
//          64-bit bit planes (most significant bit of the residue code first; residue i of the group is bit i);
//          FileBitwiseInputStream.get (src/com/rtg/reader/FileBitwiseInputStream.java:151-169) reads them back.  The
//          kernel turns the planes straight into the engine's resident template (2 bits per base + N mask per 16 bases)
//          and the byte-per-base array the caller keeps.
//
// The pass bodies are shared by the CUDA instantiation (vce_vcf.cu) and the test-only host emulation (one lane).
#ifndef VCE_BGZF_H_
#define VCE_BGZF_H_

#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "vce_device.h"

namespace vce {

// ------------------------------------------------------------------------------------------------ BGZF members
struct BgzfBlock {
  int64_t src;       // offset of the DEFLATE payload in the file
  int64_t dst;       // offset of the block's text in the inflated output
  int32_t srcLen;    // payload bytes
  int32_t dstLen;    // ISIZE (RFC 1952): bytes of text
  uint32_t crc;      // CRC32 of the text
  int32_t hdrBack;   // src - hdrBack = offset of the member's header
};

enum : int32_t {
  kInfOk = 0,
  kInfBadType = 1,      // BTYPE 3
  kInfBadStored = 2,    // LEN != ~NLEN
  kInfBadLengths = 3,   // over-subscribed code, repeat without a previous length, too many lengths
  kInfBadCode = 4,      // bit pattern that is no code / length symbol 286, 287 / distance symbol 30, 31
  kInfTooFar = 5,       // distance reaches before the start of the block
  kInfOverflow = 6,     // more text than ISIZE
  kInfOverrun = 7,      // payload exhausted before the final block ended
  kInfShort = 8,        // less text than ISIZE
  kInfCrc = 9           // CRC32 mismatch (only when checked)
};

inline const char* bgzfStatusText(int s) {
  static const char* const t[] = {"ok", "invalid block type", "invalid stored block lengths", "invalid code lengths set",
                                  "invalid literal/length or distance code", "invalid distance too far back",
                                  "more data than the block's ISIZE", "compressed data ends early",
                                  "did not inflate the expected amount", "CRC32 mismatch"};
  return s >= 0 && s <= 9 ? t[s] : "?";
}

// Host: walks the members of data[from, ...) up to (excluding) the member that starts at or after `toExcl`.
// Returns "" or what is wrong with the file (BlockCompressedInputStream's "Invalid GZIP header" family).
inline std::string bgzfWalk(const uint8_t* data, int64_t len, int64_t from, int64_t toExcl, std::vector<BgzfBlock>& out) {
  auto le16 = [&](int64_t p) { return (uint32_t) data[p] | ((uint32_t) data[p + 1] << 8); };
  auto le32 = [&](int64_t p) { return le16(p) | (le16(p + 2) << 16); };
  int64_t pos = from, dst = out.empty() ? 0 : out.back().dst + out.back().dstLen;
  while (pos < toExcl && pos < len) {
    if (len - pos < 18) return "truncated BGZF block header at offset " + std::to_string(pos);
    if (data[pos] != 0x1f || data[pos + 1] != 0x8b || data[pos + 2] != 8 || !(data[pos + 3] & 4))
      return "not a block compressed (BGZF) file: invalid GZIP header at offset " + std::to_string(pos);
    const int xlen = (int) le16(pos + 10);
    if (len - pos < 12 + xlen) return "truncated BGZF extra field at offset " + std::to_string(pos);
    int bsize = -1;
    for (int64_t p = pos + 12; p + 4 <= pos + 12 + xlen;) {
      const int slen = (int) le16(p + 2);
      if (data[p] == 'B' && data[p + 1] == 'C' && slen == 2 && p + 6 <= pos + 12 + xlen) bsize = (int) le16(p + 4) + 1;
      p += 4 + slen;
    }
    if (bsize < 0) return "not a block compressed (BGZF) file: no BC field at offset " + std::to_string(pos);
    const int payload = bsize - xlen - 20;
    if (payload < 0 || pos + bsize > len) return "truncated BGZF block at offset " + std::to_string(pos);
    BgzfBlock b;
    b.src = pos + 12 + xlen;
    b.srcLen = payload;
    b.crc = le32(pos + bsize - 8);
    const uint32_t isize = le32(pos + bsize - 4);
    if (isize > 65536u) return "BGZF block of more than 64 KiB at offset " + std::to_string(pos);
    b.dstLen = (int32_t) isize;
    b.dst = dst;
    b.hdrBack = 12 + xlen;
    dst += b.dstLen;
    out.push_back(b);
    pos += bsize;
  }
  return "";
}

// ------------------------------------------------------------------------------------------------ DEFLATE, one block per warp
struct HostWarp {
  static constexpr int W = 1;
  static int lane() { return 0; }
  static void sync() {}
};
#if defined(__CUDACC__)
struct CudaWarp {
  static constexpr int W = 32;
  __device__ static int lane() { return (int) (threadIdx.x & 31u); }
  __device__ static void sync() { __syncwarp(); }
};
#endif

constexpr int kLitBits = 10;   // primary table of the literal/length code: codes of up to 10 bits in one lookup
constexpr int kDistBits = 8;
constexpr int kClBits = 7;     // the code-length code has at most 7 bits: always one lookup

struct HuffMeta {            // canonical code description for the bit-by-bit walk (codes longer than the primary table)
  uint16_t count[16];        // codes per length
  uint16_t first[16];        // first canonical code of each length
  uint16_t index[16];        // position of that code's symbol in `sym`
};

// Pre-decoded table entry of the literal/length and distance codes: what the inner loop needs, without arithmetic on the symbol.
//   bits 0-3  code length (0 = the code is longer than the primary table: canonical walk)
//   bits 4-7  number of extra bits
//   bits 8-23 value: the literal, or the base of the length / distance (RFC 1951 3.2.5)
//   bits 24-25 kind
enum : uint32_t { kEntLiteral = 0, kEntBase = 1, kEntEnd = 2, kEntInvalid = 3 };
VCE_HD uint32_t litEntry(int s, int codeLen) {          // literal/length symbol s
  uint32_t kind, val = 0, x = 0;
  if (s < 256) { kind = kEntLiteral; val = (uint32_t) s; }
  else if (s == 256) kind = kEntEnd;
  else if (s > 285) kind = kEntInvalid;
  else {
    kind = kEntBase;
    const int li = s - 257;
    if (li < 8) val = 3 + (uint32_t) li;
    else if (li == 28) val = 258;
    else { x = (uint32_t) (li - 4) >> 2; val = 3 + ((4 + ((uint32_t) li & 3u)) << x); }
  }
  return (uint32_t) codeLen | (x << 4) | (val << 8) | (kind << 24);
}
VCE_HD uint32_t distEntry(int s, int codeLen) {         // distance symbol s
  uint32_t kind = kEntBase, val = 0, x = 0;
  if (s > 29) kind = kEntInvalid;
  else if (s < 4) val = 1 + (uint32_t) s;
  else { x = ((uint32_t) s >> 1) - 1; val = 1 + ((2 + ((uint32_t) s & 1u)) << x); }
  return (uint32_t) codeLen | (x << 4) | (val << 8) | (kind << 24);
}

struct InflateWs {           // per-warp workspace (shared memory on the device): 10.4 KB
  uint32_t litTab[1 << kLitBits];    // pre-decoded entries, 0 = not in the table
  uint32_t distTab[1 << kDistBits];
  uint16_t clTab[1 << kClBits];      // code-length code: (symbol << 4) | length
  uint16_t litSym[288];      // symbols ordered by (length, symbol)
  uint16_t distSym[32];
  HuffMeta lit, dist;
  uint8_t lens[320];         // code lengths: literal/length then distance
  int32_t scratch[4];
  // the last kRing bytes of text: a match whose distance reaches no further back is copied out of shared memory (text this
  // warp wrote a moment ago comes back from L2 otherwise: that wait was 30 % of the kernel's stall samples)
  uint8_t ring[4096];
};
constexpr int kRing = 4096;

// LSB-first bit stream over aligned 32-bit words (little endian).  Two words are held in registers, `nxt` one word ahead of
// what is being decoded (its load was issued a whole symbol earlier); a look at the next 32 bits is one funnel shift.
struct BitReader {
  const uint32_t* wp;        // word after `nxt`
  const uint32_t* wStop;     // loading at or beyond this word means the payload was overrun by more than the slack
  uint32_t cur, nxt;
  int off;                   // bits of `cur` already consumed, 0..31
  bool over;
  VCE_HD uint32_t load() {
    uint32_t w = 0;
    if (wp < wStop) w = *wp; else over = true;
    ++wp;
    return w;
  }
  VCE_HD void initAt(const uint8_t* p, const uint8_t* end) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t) 3);
    wStop = reinterpret_cast<const uint32_t*>((reinterpret_cast<uintptr_t>(end) + 3) & ~(uintptr_t) 3) + 2;
    over = false;
    cur = load();
    nxt = load();
    off = (int) (a & 3) * 8;
  }
  VCE_HD uint32_t peek() const {   // the next 32 bits
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(cur, nxt, (uint32_t) off);
#else
    return (uint32_t) ((((uint64_t) nxt << 32) | cur) >> off);
#endif
  }
  VCE_HD void drop(int n) {        // n <= 32
    off += n;
    if (off >= 32) { cur = nxt; nxt = load(); off -= 32; }
  }
  VCE_HD uint32_t take(int n) { const uint32_t v = peek() & ((1u << n) - 1u); drop(n); return v; }
  // address of the first byte that no consumed bit belongs to
  VCE_HD const uint8_t* bytePos() const { return reinterpret_cast<const uint8_t*>(wp - 2) + ((off + 7) >> 3); }
};

VCE_HD uint32_t bitReverse(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

// Builds the decoding structures of one canonical Huffman code from its code lengths (RFC 1951 3.2.2).
// KIND 0: code-length code (16-bit entries (symbol << 4) | length), 1: literal/length code, 2: distance code (pre-decoded
// 32-bit entries).  Returns false when the lengths over-subscribe the code space.
template <class WP, int KIND, class E>
VCE_HDN bool huffBuild(const uint8_t* lens, int n, E* tab, int tabBits, uint16_t* sym, HuffMeta* meta, int32_t* flag) {
  const int lane = WP::lane();
  WP::sync();   // `lens` written by lane 0
  if (lane == 0) {
    for (int l = 0; l < 16; ++l) meta->count[l] = 0;
    for (int s = 0; s < n; ++s) ++meta->count[lens[s]];
    int left = 1, ok = 1;
    uint32_t code = 0, idx = 0;
    for (int l = 1; l < 16; ++l) {
      left = (left << 1) - (int) meta->count[l];
      if (left < 0) ok = 0;
      code = (code + (l > 1 ? meta->count[l - 1] : 0)) << 1;
      meta->first[l] = (uint16_t) code;
      meta->index[l] = (uint16_t) idx;
      idx += meta->count[l];
    }
    *flag = ok;
    if (ok) {
      uint16_t next[16];
      for (int l = 1; l < 16; ++l) next[l] = meta->index[l];
      for (int s = 0; s < n; ++s) if (lens[s]) sym[next[lens[s]]++] = (uint16_t) s;
    }
  }
  for (int i = lane; i < (1 << tabBits); i += WP::W) tab[i] = 0;
  WP::sync();
  if (!*flag) return false;
  const int nCodes = (int) meta->index[15] + (int) meta->count[15];
  for (int i = lane; i < nCodes; i += WP::W) {
    const int s = sym[i];
    const int l = lens[s];
    if (l > tabBits) continue;
    const uint32_t code = (uint32_t) meta->first[l] + (uint32_t) (i - (int) meta->index[l]);
    const E e = KIND == 0 ? (E) ((s << 4) | l) : KIND == 1 ? (E) litEntry(s, l) : (E) distEntry(s, l);
    for (uint32_t k = bitReverse(code, l); k < (1u << tabBits); k += 1u << l) tab[k] = e;
  }
  WP::sync();
  return true;
}

// The canonical walk bit by bit, for codes longer than the primary table (out of line: rare, and 15 unrolled steps at every
// call site would crowd the instruction cache).  -1 = no such code.
VCE_HDN int huffWalk(uint32_t v, const uint16_t* sym, const HuffMeta* meta, int* used) {
  uint32_t code = 0;
  for (int l = 1; l < 16; ++l) {
    code = (code << 1) | (v & 1u);
    v >>= 1;
    const uint32_t c = meta->count[l], f = meta->first[l];
    if (code >= f && code - f < c) { *used = l; return sym[meta->index[l] + (code - f)]; }
  }
  *used = 1;
  return -1;
}
// One symbol of the code-length code from the bits `v` (the reader's next 32 bits); the code's length goes to *used.
VCE_HD int huffDecode(uint32_t v, const uint16_t* tab, int tabBits, const uint16_t* sym, const HuffMeta* meta, int* used) {
  const uint32_t e = tab[v & ((1u << tabBits) - 1u)];
  if (e) { *used = (int) (e & 15u); return (int) (e >> 4); }
  return huffWalk(v, sym, meta, used);
}
// Pre-decoded entry for a code longer than the primary table (or for no code at all: kEntInvalid).
template <int KIND>
VCE_HDN uint32_t huffLong(uint32_t v, const uint16_t* sym, const HuffMeta* meta) {
  int used;
  const int s = huffWalk(v, sym, meta, &used);
  if (s < 0) return 1u | (kEntInvalid << 24);
  return KIND == 1 ? litEntry(s, used) : distEntry(s, used);
}
VCE_HD uint32_t bitField(uint32_t v, uint32_t pos, uint32_t n) {   // n bits of v from bit pos (n may be 0)
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(pos), "r"(n));
  return r;
#else
  return (v >> pos) & ((1u << n) - 1u);
#endif
}

// Inflates one raw DEFLATE stream src[0, srcLen) into dst[0, dstLen); every lane of the warp calls it with the same
// arguments and gets the same status.  No preset dictionary: a distance never reaches before dst[0].
template <class WP>
VCE_HDN int inflateBlock(const uint8_t* src, int srcLen, uint8_t* dst, int dstLen, InflateWs* ws) {
  const int lane = WP::lane();
  BitReader br;
  br.initAt(src, src + srcLen);
  int op = 0;
  int status = kInfOk;
  // A short match's byte is loaded when the match is decoded and stored when the NEXT match is (or at the end): the load's
  // latency -- shared memory, or L2 for a source further back than the ring -- passes while the next symbol is decoded.
  int pendPos = -1;
  uint8_t pendC = 0;
  auto flush = [&]() {
    if (pendPos >= 0) { dst[pendPos] = pendC; ws->ring[pendPos & (kRing - 1)] = pendC; pendPos = -1; }
  };
  for (bool last = false; !last && status == kInfOk;) {
    const uint32_t hdr = br.take(3);
    last = (hdr & 1u) != 0;
    const uint32_t type = hdr >> 1;
    if (type == 3) { status = kInfBadType; break; }
    if (type == 0) {
      br.drop((8 - (br.off & 7)) & 7);
      const uint32_t v = br.peek();
      const uint32_t n = v & 0xffffu, nn = v >> 16;
      br.drop(32);
      if ((n ^ nn) != 0xffffu) { status = kInfBadStored; break; }
      const uint8_t* p = br.bytePos();
      if (p + n > src + srcLen) { status = kInfOverrun; break; }
      if (op + (int) n > dstLen) { status = kInfOverflow; break; }
      for (int i = lane; i < (int) n; i += WP::W) { const uint8_t c = p[i]; dst[op + i] = c; ws->ring[(op + i) & (kRing - 1)] = c; }
      op += (int) n;
      br.initAt(p + n, src + srcLen);
      continue;
    }
    int nLit, nDist;
    if (type == 1) {          // fixed code (RFC 1951 3.2.6)
      WP::sync();
      for (int s = lane; s < 320; s += WP::W) ws->lens[s] = (uint8_t) (s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : s < 288 ? 8 : 5);
      nLit = 288;
      nDist = 32;
    } else {                  // dynamic code (3.2.7)
      const uint32_t h = br.take(14);
      nLit = (int) (h & 31u) + 257;
      nDist = (int) ((h >> 5) & 31u) + 1;
      const int nCl = (int) (h >> 10) + 4;
      if (nLit > 286 || nDist > 30) { status = kInfBadLengths; break; }
      WP::sync();
      if (lane == 0) for (int i = 0; i < 19; ++i) ws->lens[i] = 0;
      for (int i = 0; i < nCl; ++i) {
        const uint32_t v = br.take(3);
        // order of the code-length code lengths (16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15)
        const int pos = i < 3 ? 16 + i : i == 3 ? 0 : (i & 1) ? 7 - ((i - 5) >> 1) : 6 + (i >> 1);
        if (lane == 0) ws->lens[pos] = (uint8_t) v;
      }
      if (!huffBuild<WP, 0>(ws->lens, 19, ws->clTab, kClBits, ws->distSym, &ws->dist, &ws->scratch[0])) { status = kInfBadLengths; break; }
      // the lengths are decoded by every lane alike; lane 0 writes them over the code-length lengths (the table above was
      // derived from those, it does not refer to them)
      int got = 0, prev = 0;
      uint8_t* out = ws->lens;
      WP::sync();
      while (got < nLit + nDist && status == kInfOk) {
        const uint32_t v = br.peek();
        int used;
        const int s = huffDecode(v, ws->clTab, kClBits, ws->distSym, &ws->dist, &used);
        if (s < 0) { status = kInfBadCode; break; }
        const uint32_t x = v >> used;
        int rep = 1, val = s;
        if (s == 16) { if (got == 0) { status = kInfBadLengths; break; } val = prev; rep = 3 + (int) (x & 3u); used += 2; }
        else if (s == 17) { val = 0; rep = 3 + (int) (x & 7u); used += 3; }
        else if (s == 18) { val = 0; rep = 11 + (int) (x & 127u); used += 7; }
        br.drop(used);
        if (got + rep > nLit + nDist) { status = kInfBadLengths; break; }
        if (lane == 0) for (int r = 0; r < rep; ++r) out[got + r] = (uint8_t) val;
        got += rep;
        prev = val;
      }
      if (status != kInfOk) break;
      WP::sync();
      if (ws->lens[256] == 0) { status = kInfBadLengths; break; }   // no end-of-block code
    }
    // the distance lengths sit right behind the literal/length lengths
    if (!huffBuild<WP, 1>(ws->lens, nLit, ws->litTab, kLitBits, ws->litSym, &ws->lit, &ws->scratch[0])) { status = kInfBadLengths; break; }
    if (!huffBuild<WP, 2>(ws->lens + nLit, nDist, ws->distTab, kDistBits, ws->distSym, &ws->dist, &ws->scratch[1])) { status = kInfBadLengths; break; }
    for (;;) {
      uint32_t v = br.peek();
      uint32_t e = ws->litTab[v & ((1u << kLitBits) - 1u)];
      if (!e) e = huffLong<1>(v, ws->litSym, &ws->lit);
      uint32_t n = e & 15u;
      const uint32_t kind = e >> 24;
      if (kind == kEntLiteral) {
        if (op >= dstLen) { status = kInfOverflow; break; }
        br.drop((int) n);
        if (lane == 0) { dst[op] = (uint8_t) (e >> 8); ws->ring[op & (kRing - 1)] = (uint8_t) (e >> 8); }
        ++op;
        continue;
      }
      if (kind != kEntBase) {
        if (kind == kEntEnd) br.drop((int) n); else status = kInfBadCode;
        break;
      }
      uint32_t x = (e >> 4) & 15u;
      const int len = (int) (((e >> 8) & 0xffffu) + bitField(v, n, x));
      br.drop((int) (n + x));        // at most 15 + 5 bits
      v = br.peek();
      uint32_t d = ws->distTab[v & ((1u << kDistBits) - 1u)];
      if (!d) d = huffLong<2>(v, ws->distSym, &ws->dist);
      n = d & 15u;
      x = (d >> 4) & 15u;
      const int dist = (int) (((d >> 8) & 0xffffu) + bitField(v, n, x));
      br.drop((int) (n + x));        // at most 15 + 13 bits
      // one test for everything that can be wrong with a match; which one it was is sorted out on the way out
      if (((d >> 24) != kEntBase) | (dist > op) | (op + len > dstLen) | br.over) {
        status = (d >> 24) != kEntBase ? kInfBadCode : dist > op ? kInfTooFar : op + len > dstLen ? kInfOverflow : kInfOverrun;
        break;
      }
      flush();
      WP::sync();   // lane 0's literals and the previous copies are visible to every lane
      if (len <= WP::W && dist >= len) {            // the common case: one load per lane now, its two stores later
        if (lane < len) {
          pendC = dist <= kRing ? ws->ring[(op - dist + lane) & (kRing - 1)] : dst[op - dist + lane];
          pendPos = op + lane;
        }
      } else if (dist <= kRing) {                   // the source is in the ring (every byte of text goes through it)
        const int src0 = op - dist;
        {
          // longer than a warp, or overlapping itself (bytes [op - dist, op) repeat with period dist).  The ring holds kRing
          // bytes and dist + len may exceed that: a round's reads are finished before its writes start.
          for (int i0 = 0; i0 < len; i0 += WP::W) {
            const int i = i0 + lane;
            uint8_t c = 0;
            if (i < len) c = ws->ring[(src0 + (dist >= len ? i : i % dist)) & (kRing - 1)];
            WP::sync();
            if (i < len) { dst[op + i] = c; ws->ring[(op + i) & (kRing - 1)] = c; }
            WP::sync();
          }
        }
      } else {
        const uint8_t* from = dst + op - dist;      // further back than the ring: from the text in global memory (dist > len)
        for (int i = lane; i < len; i += WP::W) { const uint8_t c = from[i]; dst[op + i] = c; ws->ring[(op + i) & (kRing - 1)] = c; }
      }
      op += len;
    }
    if (br.over && status == kInfOk) status = kInfOverrun;
  }
  flush();
  if (status == kInfOk) {
    if (br.bytePos() > src + srcLen) status = kInfOverrun;
    else if (op != dstLen) status = kInfShort;
  }
  WP::sync();
  return status;
}

// CRC32 (RFC 1952 8.1.1, polynomial 0xedb88320), one thread per block, byte by byte through a 256-entry table.
VCE_HD uint32_t crc32TableEntry(uint32_t i) {
  uint32_t c = i;
  for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
  return c;
}
VCE_HD uint32_t crc32Bytes(const uint8_t* p, int n, const uint32_t* table) {
  uint32_t c = 0xffffffffu;
  for (int i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xffu] ^ (c >> 8);
  return c ^ 0xffffffffu;
}

// Host: bytes of text between two virtual offsets, from the members' ISIZE fields alone.
inline std::string bgzfTextSize(const uint8_t* data, int64_t len, int64_t vBeg, int64_t vEnd, int64_t* textLen) {
  const int64_t cBeg = vBeg >> 16, uBeg = vBeg & 0xffff;
  const int64_t cEnd = vEnd >= 0 ? vEnd >> 16 : len, uEnd = vEnd >= 0 ? vEnd & 0xffff : 0;
  if (vBeg < 0 || cBeg > len || cEnd > len) return "BGZF: virtual offsets outside the file";
  std::vector<BgzfBlock> blocks;
  const std::string w = bgzfWalk(data, len, cBeg, uEnd > 0 ? cEnd + 1 : cEnd, blocks);
  if (!w.empty()) return w;
  int64_t end = blocks.empty() ? 0 : blocks.back().dst + blocks.back().dstLen;
  if (uEnd > 0 && !blocks.empty()) end = blocks.back().dst + uEnd;
  *textLen = end > uBeg ? end - uBeg : 0;
  return "";
}

// ------------------------------------------------------------------------------------------------ driver
struct BgzfStats {
  int64_t blocks = 0, compressed = 0, text = 0;
  double h2dMs = 0, kernelMs = 0, d2hMs = 0;
  int launches = 0;
};

// The text between two virtual offsets (compressed offset << 16 | offset within the block's text), as TabixLineReader
// walks it.  vEnd < 0: to the end of the file.  On success *dText / *textLen describe the slice in device memory
// (owned by the ops' slab); `hostOut` (optional) receives a copy.
template <class Ops>
int bgzfDrive(Ops& ops, const uint8_t* data, int64_t len, int64_t vBeg, int64_t vEnd, bool checkCrc, const uint8_t** dText,
              int64_t* textLen, std::vector<uint8_t>* hostOut, BgzfStats& st, std::string& err) {
  *dText = nullptr;
  *textLen = 0;
  if (!data || len < 0 || vBeg < 0) { err = "BGZF: bad arguments"; return 1; }
  const int64_t cBeg = vBeg >> 16, uBeg = vBeg & 0xffff;
  int64_t cEnd = len, uEnd = 0;
  if (vEnd >= 0) { cEnd = vEnd >> 16; uEnd = vEnd & 0xffff; }
  if (cBeg > len || cEnd > len || (vEnd >= 0 && vEnd < vBeg)) { err = "BGZF: virtual offsets outside the file"; return 1; }
  std::vector<BgzfBlock> blocks;
  // the member that holds the end offset is needed when the end lies inside its text
  std::string w = bgzfWalk(data, len, cBeg, uEnd > 0 ? cEnd + 1 : cEnd, blocks);
  if (!w.empty()) { err = w; return 1; }
  int64_t total = blocks.empty() ? 0 : blocks.back().dst + blocks.back().dstLen;
  int64_t sliceEnd = total;
  if (uEnd > 0) {
    if (blocks.empty() || blocks.back().src - blocks.back().hdrBack != cEnd || uEnd > blocks.back().dstLen) { err = "BGZF: end offset beyond its block"; return 1; }
    sliceEnd = blocks.back().dst + uEnd;
  }
  if (uBeg > 0 && (blocks.empty() || uBeg > blocks[0].dstLen)) { err = "BGZF: start offset beyond its block"; return 1; }
  if (sliceEnd < uBeg) sliceEnd = uBeg;
  if (total >= 0x7ffffff0LL) { err = "BGZF: 2 GB of text or more: hand it over in pieces (one sequence per call)"; return 1; }
  const int nb = (int) blocks.size();
  const int64_t span = nb ? blocks.back().src + blocks.back().srcLen + 8 - cBeg : 0;   // compressed bytes needed
  for (BgzfBlock& b : blocks) b.src -= cBeg;
  st.blocks = nb;
  st.compressed = span;
  st.text = sliceEnd - uBeg;
  ops.mark(0);
  uint8_t* dData = ops.template alloc<uint8_t>((size_t) span + 64);
  BgzfBlock* dBlocks = ops.template alloc<BgzfBlock>((size_t) nb + 1);
  int32_t* dStatus = ops.template alloc<int32_t>((size_t) nb + 1);
  uint8_t* dOut = ops.template alloc<uint8_t>((size_t) total + 64);
  if (!dData || !dBlocks || !dStatus || !dOut) { err = "out of device memory for the BGZF blocks"; return 2; }
  ops.upload(dData, data + cBeg, (size_t) span);
  ops.upload(dBlocks, blocks.data(), (size_t) nb * sizeof(BgzfBlock));
  ops.mark(1);
  ops.inflate(nb, dBlocks, dData, dOut, dStatus, checkCrc);
  std::vector<int32_t> status((size_t) nb);
  ops.download(status.data(), dStatus, (size_t) nb * sizeof(int32_t));
  for (int b = 0; b < nb; ++b) {
    if (status[(size_t) b] != kInfOk) {
      err = "BGZF block at offset " + std::to_string(cBeg + blocks[(size_t) b].src - blocks[(size_t) b].hdrBack) + ": " + bgzfStatusText(status[(size_t) b]);
      return 3;
    }
  }
  *dText = dOut + uBeg;
  *textLen = sliceEnd - uBeg;
  if (hostOut) {
    hostOut->resize((size_t) *textLen);
    ops.download(hostOut->data(), *dText, (size_t) *textLen);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ tabix index (host)
// The inflated index (TabixIndexReader.java:63-146).  Positions are byte offsets into `raw`.
struct TabixIndex {
  std::vector<uint8_t> raw;
  int32_t format = 0, seqCol = 0, begCol = 0, endCol = 0, meta = 0, skip = 0;
  std::vector<std::string> names;
  std::vector<int64_t> binPos, linPos;
};

inline std::string tabixParse(TabixIndex& t) {
  const std::vector<uint8_t>& b = t.raw;
  auto le32 = [&](size_t p) { return (int32_t) ((uint32_t) b[p] | ((uint32_t) b[p + 1] << 8) | ((uint32_t) b[p + 2] << 16) | ((uint32_t) b[p + 3] << 24)); };
  if (b.size() < 36) return "not a valid TABIX index. (index does not have a complete header)";
  if (b[0] != 'T' || b[1] != 'B' || b[2] != 'I' || b[3] != 1) return "not a valid TABIX index. (index does not have a valid header)";
  const int32_t nRef = le32(4);
  t.format = le32(8);
  t.seqCol = le32(12) - 1;
  t.begCol = le32(16) - 1;
  t.endCol = le32(20) - 1;
  t.meta = le32(24);
  t.skip = le32(28);
  if (t.seqCol < 0) return "not a valid TABIX index. (invalid col_seq)";
  if (t.begCol < 0) return "not a valid TABIX index. (invalid col_beg)";
  const int32_t nameLen = le32(32);
  if (nameLen < 0 || nRef < 0 || b.size() < (size_t) 36 + (size_t) nameLen) return "not a valid TABIX index. (sequence names truncated)";
  size_t start = 36;
  for (size_t i = 36; i < (size_t) 36 + (size_t) nameLen; ++i) {
    if (b[i] == 0) {
      if ((int32_t) t.names.size() >= nRef) return "not a valid TABIX index. (more sequence names provided than expected)";
      t.names.emplace_back(reinterpret_cast<const char*>(b.data() + start), i - start);
      start = i + 1;
    }
  }
  if ((int32_t) t.names.size() != nRef) return "not a valid TABIX index. (insufficient number of sequence names provided)";
  size_t pos = (size_t) 36 + (size_t) nameLen;
  for (int32_t s = 0; s < nRef; ++s) {
    t.binPos.push_back((int64_t) pos);
    if (pos + 4 > b.size()) return "not a valid TABIX index. (truncated)";
    const int32_t nBins = le32(pos);
    pos += 4;
    for (int32_t i = 0; i < nBins; ++i) {
      if (pos + 8 > b.size()) return "not a valid TABIX index. (truncated)";
      const int32_t nChunks = le32(pos + 4);
      if (nChunks < 0) return "not a valid index. (numChunks = " + std::to_string(nChunks) + ")";
      pos += 8 + (size_t) nChunks * 16;
    }
    t.linPos.push_back((int64_t) pos);
    if (pos + 4 > b.size()) return "not a valid TABIX index. (truncated)";
    const int32_t nLin = le32(pos);
    if (nLin < 0) return "not a valid index. (linearIndexSize = " + std::to_string(nLin) + ")";
    pos += 4 + (size_t) nLin * 8;
    if (pos > b.size()) return "not a valid TABIX index. (truncated)";
  }
  return "";
}

// AbstractIndexReader.getFilePointers(seqName, intervals...) for ONE interval [beg, end) (0-based, end exclusive; end < 0 or
// INT_MAX = to the end of the sequence): the smallest chunk begin / largest chunk end over the region's bins, linear-index
// filtered (:170-199).  Returns 1 and the offsets, 0 when the index holds nothing for the region, -1 on a bad request.
inline int tabixQuery(const TabixIndex& t, const std::string& name, int32_t beg0, int32_t end0, int64_t* vBeg, int64_t* vEnd, std::string& err) {
  int seq = -1;
  for (size_t i = 0; i < t.names.size(); ++i) if (t.names[i] == name) seq = (int) i;   // HashMap.put: the last one wins
  if (seq < 0) return 0;
  const std::vector<uint8_t>& b = t.raw;
  auto le32 = [&](size_t p) { return (int32_t) ((uint32_t) b[p] | ((uint32_t) b[p + 1] << 8) | ((uint32_t) b[p + 2] << 16) | ((uint32_t) b[p + 3] << 24)); };
  auto le64 = [&](size_t p) { return (uint64_t) (uint32_t) le32(p) | ((uint64_t) (uint32_t) le32(p + 4) << 32); };
  constexpr int32_t kMaxCoord = 1 << 29;
  const int32_t beg = (beg0 <= -1) ? 0 : beg0;
  const int32_t end = (end0 <= -1 || end0 == 0x7fffffff) ? kMaxCoord : end0;
  if (end > kMaxCoord || beg > kMaxCoord) { err = "The requested region contains coordinates greater than can be addressed by tabix/bam indexes"; return -1; }
  const size_t linAt = (size_t) t.linPos[(size_t) seq];
  const int32_t nLin = le32(linAt);
  const int32_t linearBeg = beg >> 14;
  // bins of the region (reg2bins :238-262)
  std::vector<int32_t> bins;
  const int32_t cEnd = end - 1;
  bins.push_back(0);
  for (int32_t k = 1 + (beg >> 26); k <= 1 + (cEnd >> 26); ++k) bins.push_back(k);
  for (int32_t k = 9 + (beg >> 23); k <= 9 + (cEnd >> 23); ++k) bins.push_back(k);
  for (int32_t k = 73 + (beg >> 20); k <= 73 + (cEnd >> 20); ++k) bins.push_back(k);
  for (int32_t k = 585 + (beg >> 17); k <= 585 + (cEnd >> 17); ++k) bins.push_back(k);
  for (int32_t k = 4681 + (beg >> 14); k <= 4681 + (cEnd >> 14); ++k) bins.push_back(k);
  // bin id -> position of its chunk list (a later duplicate of a bin id replaces the earlier one, :154)
  constexpr int32_t kNumBins = ((1 << 18) - 1) / 7 + 2;
  std::vector<int64_t> where((size_t) kNumBins, -1);
  size_t pos = (size_t) t.binPos[(size_t) seq];
  const int32_t nBins = le32(pos);
  pos += 4;
  for (int32_t i = 0; i < nBins; ++i) {
    const uint32_t id = (uint32_t) le32(pos);
    const int32_t nChunks = le32(pos + 4);
    if (id >= (uint32_t) kNumBins) { err = "not a valid index. (bin " + std::to_string(id) + ")"; return -1; }   // the reference: ArrayIndexOutOfBounds
    where[id] = (int64_t) pos;
    pos += 8 + (size_t) nChunks * 16;
  }
  bool have = false;
  uint64_t lo = 0, hi = 0;
  const uint64_t linMin = linearBeg < nLin ? le64(linAt + 4 + (size_t) linearBeg * 8) : 0;
  for (int32_t bin : bins) {
    if (bin >= kNumBins || where[(size_t) bin] < 0) continue;
    const size_t at = (size_t) where[(size_t) bin];
    const int32_t nChunks = le32(at + 4);
    for (int32_t k = 0; k < nChunks; ++k) {
      const uint64_t cb = le64(at + 8 + (size_t) k * 16), ce = le64(at + 16 + (size_t) k * 16);
      if (bin >= 4681 || (linearBeg < nLin && linMin < ce)) {
        if (!have || cb < lo) lo = cb;
        if (!have || hi < ce) hi = ce;
        have = true;
      }
    }
  }
  if (!have) return 0;
  *vBeg = (int64_t) lo;
  *vEnd = (int64_t) hi;
  return 1;
}

// ------------------------------------------------------------------------------------------------ SDF sequence data
// One 16-base granule of the resident template per call: residues [first + g * 16, ...) of the bit-plane file become
// packed[g] (2 bits per base, base code - 1), the granule's N bit and 16 bytes of the byte-per-base array.
// Codes: 0 = N, 1..4 = A C G T (DNA.ordinal()); codes 5..7 cannot be written by the formatter and are kept as they are
// in the byte array (they count as "other" in the mask, like N).
VCE_HD uint64_t sdfPlane(const uint8_t* file, int64_t longIdx) {   // big-endian 64-bit word (DataOutputStream.writeLong)
#if defined(__CUDA_ARCH__)
  // (the planes sit at multiples of 8 bytes of a 256-byte aligned buffer)
  const uint2 w = *reinterpret_cast<const uint2*>(file + longIdx * 8);
  return ((uint64_t) __byte_perm(w.x, 0, 0x0123) << 32) | (uint64_t) __byte_perm(w.y, 0, 0x0123);
#else
  const uint8_t* p = file + longIdx * 8;
  uint64_t v = 0;
  for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
  return v;
#endif
}
VCE_HD uint32_t sdfSpread16(uint32_t x) {   // bit i -> bit 2 i
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
// The granule's 16 residues as three 16-bit masks (bit q = residue q), straight out of the planes: the packed word and the N flag
// are bit logic on the masks -- with x = 2 * plane1 + plane2, (code - 1) & 3 = (x - 1) & 3 whatever plane0 says, i.e. low bit =
// ~plane2, high bit = ~(plane1 ^ plane2) -- and the byte-per-base array takes the masks four residues at a time.
VCE_HD void sdfGranule(const uint8_t* file, int64_t first, int32_t len, int32_t g, uint32_t* packed, uint8_t* otherFlag, uint8_t* bases) {
  const int32_t p0 = g * 16;
  const int valid = len - p0 < 16 ? len - p0 : 16;
  const int64_t r = first + p0;
  const int64_t group = r >> 6;
  const int bit = (int) (r & 63);
  uint64_t a = sdfPlane(file, group * 3) >> bit, b = sdfPlane(file, group * 3 + 1) >> bit, c = sdfPlane(file, group * 3 + 2) >> bit;
  if (bit + valid > 64) {   // the granule runs on into the next group of 64 residues
    a |= sdfPlane(file, group * 3 + 3) << (64 - bit);
    b |= sdfPlane(file, group * 3 + 4) << (64 - bit);
    c |= sdfPlane(file, group * 3 + 5) << (64 - bit);
  }
  const uint32_t m = valid == 16 ? 0xffffu : (1u << valid) - 1u;
  const uint32_t A = (uint32_t) a & m, B = (uint32_t) b & m, C = (uint32_t) c & m;
  packed[g] = sdfSpread16(~C & m) | (sdfSpread16(~(B ^ C) & m) << 1);
  otherFlag[g] = (((~A & ~B & ~C) | (A & (B | C))) & m) ? 1 : 0;
  if (!bases) return;
  uint32_t w[4];
  for (int k = 0; k < 4; ++k) {   // four residues per word: bit j of the nibble -> byte j
    const uint32_t na = (A >> (4 * k)) & 15u, nb = (B >> (4 * k)) & 15u, nc = (C >> (4 * k)) & 15u;
    w[k] = (((na * 0x00204081u) & 0x01010101u) << 2) | (((nb * 0x00204081u) & 0x01010101u) << 1) | ((nc * 0x00204081u) & 0x01010101u);
  }
#if defined(__CUDA_ARCH__)
  if (valid == 16) { *reinterpret_cast<uint4*>(bases + p0) = make_uint4(w[0], w[1], w[2], w[3]); return; }
#endif
  for (int q = 0; q < valid; ++q) bases[p0 + q] = (uint8_t) (w[q >> 2] >> (8 * (q & 3)));
}
// mask word m gathers the flags of granules [32 m, 32 m + 32)
VCE_HD void sdfMaskWord(const uint8_t* otherFlag, int32_t granules, int32_t m, uint32_t* nmask) {
  uint32_t mask = 0;
  for (int gi = 0; gi < 32; ++gi) {
    const int32_t g = m * 32 + gi;
    if (g < granules && otherFlag[g]) mask |= 1u << gi;
  }
  nmask[m] = mask;
}

}  // namespace vce
#endif
