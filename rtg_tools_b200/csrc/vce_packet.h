// vce_packet.h -- building the self-contained region packets of the thread-per-region kernel (layout: vce_micro.h).
//
// ONE source for two places: the host digest (staged jobs, settle retries, unsorted inputs) and the device builder
// kernel k_build (vce_cuda.cu), which reads wholesale copies of the caller's arrays so that the host no longer chases
// variant -> allele -> bases per region.  A PacketView makes both look alike: every pointer is indexed with the
// caller's own indices (input variant index, allele slot, nt offset) or with the pass-wide sorted variant index; the
// device side passes pointers into the uploaded slices that were moved back by the slice's first index.
#ifndef VCE_PACKET_H_
#define VCE_PACKET_H_

#include <stdint.h>

#include "vce_device.h"
#include "vce_micro.h"

namespace vce {

struct RegRange { int32_t cFirst, cCount, bFirst, bCount; };  // pass-wide sorted variant indices, calls then baseline

struct PacketSideView {
  const int32_t* vStart;     // [pass-wide variant index] locus bounds computed by the digest
  const int32_t* vEnd;
  int32_t sideEnd;           // one past the sequence's last pass-wide variant index
  int32_t first;             // pass-wide index of the sequence's first variant
  const int32_t* idx;        // (start,end)-sorted position -> input index, nullptr = identity
  int32_t gtPloidy;          // entries per variant in gt[]
  const int8_t* gt;          // [input index * gtPloidy]
  const uint8_t* phased;     // [input index] or nullptr
  const uint8_t* statusIn;   // [input index] or nullptr
  const int32_t* alleleOff;  // [input index] -> slot
  const int32_t* aStart;     // [slot]
  const int32_t* aEnd;
  const uint8_t* aNull;
  const int32_t* aNtOff;     // [slot] -> nt offset
  const uint8_t* nt;         // [nt offset]
};

struct PacketView {
  PacketSideView s[2];       // 0 = calls, 1 = baseline
  int32_t tl;                // template length
  int32_t H;                 // haplotypes of the pass
  int32_t expectPloidy[2];   // GT ploidy the pass was set up for
  int32_t margin;            // reference bases kept beyond the last variant end
};

// Window of a region: relative position 0 is one base before the first variant start.  Returns false when the region
// cannot be a packet (first / last region of a sequence are never offered).
VCE_HD bool microWindow(const PacketView& pv, const RegRange& g, int32_t* winStart, int32_t* winLen, int32_t nextStart[2]) {
  const int first[2] = {g.cFirst, g.bFirst};
  const int count[2] = {g.cCount, g.bCount};
  int firstStart = 0x7fffffff;
  long long maxEnd = -(1LL << 62);
  for (int side = 0; side < 2; ++side) {
    const PacketSideView& s = pv.s[side];
    const int e = first[side] + count[side];
    nextStart[side] = e < s.sideEnd ? s.vStart[e] : kNoNext;
    for (int v = first[side]; v < e; ++v) {
      const int st = s.vStart[v];
      const long long en = s.vEnd[v];
      firstStart = st < firstStart ? st : firstStart;
      maxEnd = en > maxEnd ? en : maxEnd;
    }
  }
  if (nextStart[0] == kNoNext && nextStart[1] == kNoNext) return false;  // the last region finishes at the template end
  if (firstStart == 0x7fffffff || firstStart < 1 || maxEnd < firstStart) return false;
  const int ws = firstStart - 1;
  long long we = maxEnd + pv.margin;
  if (we > pv.tl) we = pv.tl;
  const long long wl = we - ws;
  if (wl < 1 || wl > kMicroMaxWin) return false;
  *winStart = ws;
  *winLen = (int32_t) wl;
  return true;
}

// Where the reference bases of a window come from: the caller's template bytes (host digest, slab copies), or the
// device-resident copy of a registered template (vce_template_register): 2 bits per base (base - 1), 16 bases per word,
// plus one bit per 16-base granule that holds a base other than A/C/G/T (such windows are not packets).
struct WinBytes {
  const uint8_t* p;   // reference base of relative position 0
  VCE_HD int at(int q) const { return p[q]; }
};
struct WinPacked {
  const uint32_t* t2;
  const uint32_t* nmask;
  int32_t start;      // template position of relative position 0
  VCE_HD int at(int q) const {
    const uint32_t pos = (uint32_t) (start + q);
    if ((nmask[pos >> 9] >> ((pos >> 4) & 31u)) & 1u) return 0;
    return (int) ((t2[pos >> 4] >> ((pos & 15u) * 2u)) & 3u) + 1;
  }
};

// Writes the packet of region `g` to `out` (T::PKT_MAX bytes available); returns its length in bytes, 0 when the region
// does not fit the thread-per-region kernel.  `win` delivers the reference bases from relative position 0.
template <class T, class Win>
VCE_HD int microBuildPacketW(const PacketView& pv, const RegRange& g, int32_t winStart, int32_t winLen, const int32_t nextStart[2],
                             const Win& win, uint8_t* out);

template <class T>
VCE_HD int microBuildPacket(const PacketView& pv, const RegRange& g, int32_t winStart, int32_t winLen, const int32_t nextStart[2],
                            const uint8_t* win, uint8_t* out) {
  return microBuildPacketW<T, WinBytes>(pv, g, winStart, winLen, nextStart, WinBytes{win}, out);
}

template <class T, class Win>
VCE_HD int microBuildPacketW(const PacketView& pv, const RegRange& g, int32_t winStart, int32_t winLen, const int32_t nextStart[2],
                             const Win& win, uint8_t* out) {
  if (g.cCount > T::KV || g.bCount > T::KV) return 0;
  const int first[2] = {g.cFirst, g.bFirst};
  const int count[2] = {g.cCount, g.bCount};
  uint8_t raw[256 + 8];   // allele bases then window bases, one per byte; packed to 2 bits at the end
  int nBases = 0;
  uint8_t alleles[4 * T::KV * MA_BYTES];
  int nAl = 0;
  uint8_t* vout = out + MP_VARS;
  for (int side = 0; side < 2; ++side) {
    const PacketSideView& s = pv.s[side];
    const int P = s.gtPloidy;
    if (P != pv.expectPloidy[side]) return 0;
    if (pv.H == 2 ? P != 2 : (P < 1 || P > 2)) return 0;
    for (int v = first[side]; v < first[side] + count[side]; ++v) {
      const int li = v - s.first;
      const int ii = s.idx ? s.idx[li] : li;
      const int off0 = s.alleleOff[ii];
      const int nSl = s.alleleOff[ii + 1] - off0;
      const int g0 = s.gt[(size_t) ii * P];
      const int gt[2] = {g0, P > 1 ? (int) s.gt[(size_t) ii * P + 1] : g0};
      int al[2] = {kMicroNoAllele, kMicroNoAllele};
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && gt[1] == gt[0]) { al[1] = al[0]; break; }
        const int id = gt[h];
        if (id < -1 || id + 1 >= nSl) continue;
        const int slot = off0 + id + 1;
        if (s.aNull[slot]) continue;
        const long long st = (long long) s.aStart[slot] - winStart, en = (long long) s.aEnd[slot] - winStart;
        const int nb = s.aNtOff[slot];
        const int len = s.aNtOff[slot + 1] - nb;
        if (st < 0 || en < st || en > kMicroMaxWin || len > kMicroMaxWin || nBases + len + winLen > 255) return 0;
        const uint8_t* nt = s.nt + nb;
        for (int q = 0; q < len; ++q) raw[nBases + q] = nt[q];
        uint8_t* a = alleles + nAl * MA_BYTES;
        a[MA_START] = (uint8_t) st; a[MA_END] = (uint8_t) en; a[MA_NT] = (uint8_t) nBases; a[MA_LEN] = (uint8_t) len;
        nBases += len;
        al[h] = nAl++;
      }
      const long long vs = (long long) s.vStart[v] - winStart, ve = (long long) s.vEnd[v] - winStart;
      if (vs < 1 || ve < vs || ve > kMicroMaxWin) return 0;
      uint8_t f = (uint8_t) ((s.phased && s.phased[ii]) ? 1 : 0);
      f |= (uint8_t) (P << 1);
      if (s.statusIn) f |= s.statusIn[ii] & VCE_STATUS_OUTSIDE_EVAL;
      vout[MV_START] = (uint8_t) vs; vout[MV_END] = (uint8_t) ve; vout[MV_FLAGS] = f;
      vout[MV_GT0] = (uint8_t) (int8_t) gt[0]; vout[MV_GT1] = (uint8_t) (int8_t) gt[1];
      vout[MV_A0] = (uint8_t) al[0]; vout[MV_A1] = (uint8_t) al[1]; vout[7] = 0;
      vout += MV_BYTES;
    }
  }
  const int winBase = nBases;
  for (int q = 0; q < winLen; ++q) raw[nBases + q] = (uint8_t) win.at(q);
  nBases += winLen;
  raw[nBases] = raw[nBases + 1] = raw[nBases + 2] = 1;  // padding of the last 2-bit byte
  for (int q = 0; q < nAl * MA_BYTES; ++q) vout[q] = alleles[q];
  uint8_t* bout = vout + nAl * MA_BYTES;
  const int nb2 = (nBases + 3) >> 2;
  // four bases (1..4, one per byte) -> one byte of 2-bit codes, 32 bits at a time: y = codes 0..3 per byte,
  // t = y | y >> 6 pairs them up, t | t >> 12 brings the upper pair down.  N (0) and the explicit-unknown token (5)
  // leave bits outside 0x03 per byte: such regions do not become packets.
  uint32_t bad = 0;
  for (int q = 0; q < nb2; ++q) {
    const uint32_t x = (uint32_t) raw[4 * q] | ((uint32_t) raw[4 * q + 1] << 8) | ((uint32_t) raw[4 * q + 2] << 16) | ((uint32_t) raw[4 * q + 3] << 24);
    const uint32_t y = x - 0x01010101u;
    bad |= y;
    const uint32_t t = y | (y >> 6);
    bout[q] = (uint8_t) (t | (t >> 12));
  }
  bad &= 0xfcfcfcfcu;
  if (bad != 0) return 0;
  int bytes = (int) (bout + nb2 - out);
  while (bytes & 3) out[bytes++] = 0;
  if (bytes > T::PKT_MAX) return 0;
  out[MP_WINSTART] = (uint8_t) (winStart & 0xff);
  out[MP_WINSTART + 1] = (uint8_t) ((winStart >> 8) & 0xff);
  out[MP_WINSTART + 2] = (uint8_t) ((winStart >> 16) & 0xff);
  out[MP_WINSTART + 3] = (uint8_t) ((winStart >> 24) & 0xff);
  out[MP_COUNTS] = (uint8_t) (g.cCount | (g.bCount << 4));
  out[MP_WINLEN] = (uint8_t) winLen;
  for (int side = 0; side < 2; ++side) {
    int rel = kMicroNoNext;
    if (nextStart[side] != kNoNext) {
      const long long d = (long long) nextStart[side] - winStart;
      rel = (int) (d < kMicroFar ? d : kMicroFar);
    }
    out[side == 0 ? MP_CNEXT : MP_BNEXT] = (uint8_t) rel;
  }
  const long long tr = (long long) pv.tl - winStart;
  out[MP_TMPLLEN] = (uint8_t) (tr < kMicroFar + 2 ? tr : kMicroFar + 2);
  out[MP_NALLELES] = (uint8_t) nAl;
  out[MP_WINBASE] = (uint8_t) winBase;
  out[MP_NWORDS] = (uint8_t) (bytes / 4);
  return bytes;
}

// Shape class of a packet (15 bits): regions of one class take the same branches in the search, so packets are handed to
// the kernel grouped by class -- the lanes of a warp (32 consecutive packets) then run in lock step.
VCE_HD uint32_t microPacketClass(const uint8_t* pk) {
  const int cC = pk[MP_COUNTS] & 0xf, bC = pk[MP_COUNTS] >> 4;
  const uint8_t* al = pk + MP_VARS + (cC + bC) * MV_BYTES;
  uint32_t key = (uint32_t) (cC | (bC << 4));
  const uint8_t* v = pk + MP_VARS;
  uint32_t sig0 = 0, sig1 = 0;
  for (int i = 0; i < cC + bC && i < 16; ++i, v += MV_BYTES) {
    const int a0 = v[MV_A0], a1 = v[MV_A1];
    uint32_t l0 = 3u, l1 = 3u;
    if (a0 != kMicroNoAllele) { l0 = al[a0 * MA_BYTES + MA_LEN]; l0 = l0 < 2u ? l0 : 2u; }
    if (a1 != kMicroNoAllele) { l1 = al[a1 * MA_BYTES + MA_LEN]; l1 = l1 < 2u ? l1 : 2u; }
    uint32_t span = (uint32_t) (v[MV_END] - v[MV_START]);
    span = span < 3u ? span : 3u;
    const uint32_t het = v[MV_GT0] != v[MV_GT1] ? 1u : 0u;
    const uint32_t ph = v[MV_FLAGS] & 1u;
    const uint32_t sig = l0 | (l1 << 2) | (span << 4) | (het << 6) | (ph << 7);
    if (i == 0) sig0 = sig;
    if (i == 1) sig1 = sig;
    key = key * 257u + sig;
  }
  if (cC == 1 && bC == 1) {
    const uint8_t* vc = pk + MP_VARS;
    const uint8_t* vb = vc + MV_BYTES;
    const uint32_t same = (vc[MV_START] == vb[MV_START] && vc[MV_END] == vb[MV_END] && sig0 == sig1) ? 1u : 0u;
    key = key * 2u + same;
  }
  return key & 0x7fffu;
}

}  // namespace vce
#endif
