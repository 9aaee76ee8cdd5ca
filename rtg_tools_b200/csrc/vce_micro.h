// vce_micro.h -- the bounded best-first haplotype search for ONE small region, run by ONE thread.
//
// Behavioural contract (bit-exact, same as vce_search.h): PathFinder.bestPath and everything it calls
//   E/PathFinder.java:129-348, E/Path.java:190-357, E/HalfPath.java:140-326, E/HaplotypePlayback.java:91-365,
//   E/Allele.java:95-119, E/MaxSumBoth.java:43-82, E/MaxCallsMinBaseline.java:42-61, Path.calculateWeights
//   E/Path.java:385-453, and Orientor.orientations for the unphased / phased / squash orientors
//   (E/Orientor.java:74-187)      (E/ = src/com/rtg/vcf/eval/ of the reference).
//
// Design.  A whole-genome comparison consists of millions of tiny independent regions (one or two variants per
// side) whose searches take 5..30 iterations with a frontier below ten paths.  One warp per region wastes 31 lanes
// on them, so the bulk of the work runs "one region per thread":
//   * the region arrives as one self-contained PACKET (header, variants with genotypes, the alleles they reference,
//     2-bit packed allele and reference-window bases) -- one contiguous read from HBM into the thread's
//     shared-memory slice; every position inside is RELATIVE to the window start so that a search node fits in
//     a few dozen bytes;
//   * genotype/phase expansion (K1, Orientor.orientations) happens in place while the packet is unpacked;
//   * search nodes are fixed-size int8 records in shared memory, the frontier is a small sorted index array,
//     the free list a 32-bit mask in a register;
//   * the tie-break / weight epilogue (K3) writes one fixed-size result record.
// Anything that does not fit the fixed capacities (frontier, pending alleles, sync points, window misses, the
// reference's own too-complex budget, finishing at the template end) is reported as kMicroFallback and re-run by
// the general kernels (vce_search.h), so the capacities never change a result.
//
// The same source runs on the device (one instance per thread) and, for the host-logic tests only, on the CPU.
#ifndef VCE_MICRO_H_
#define VCE_MICRO_H_

#include "vce_device.h"
#include "../../include/vce.h"

namespace vce {

// ---- packet layout (bytes) -----------------------------------------------------------------------------
//  0..3  int32  winStart   absolute template position of relative position 0 (== position of the fresh in-sync path)
//  4     u8     cCount | bCount << 4
//  5     u8     winLen     reference bases in the window (relative positions 0 .. winLen-1)
//  6     u8     cNextRel   start of the next call variant after the region (relative, clamped to kMicroFar), 0x7f = none
//  7     u8     bNextRel
//  8     u8     tmplLenRel template length (relative, clamped to kMicroFar + 2)
//  9     u8     nAlleles
// 10     u8     winBase    index of window base 0 in the 2-bit area (== number of allele bases)
// 11     u8     nWords     packet length in 4-byte words
// 12..   variants, calls then baseline, 8 bytes each:
//          startRel, endRel, flags (bit0 phased, bit4 outside-eval, bits 1-2 GT ploidy), gt0, gt1, a0, a1, pad
//          a0 / a1 = local allele index of allele id gt0 / gt1, 0xff when that allele is null or missing
//        alleles, 4 bytes each: startRel, endRel, ntBase (index into the 2-bit area), len
//        2-bit area: allele bases then window bases, base i in bits (2i, 2i+1) of byte i/4, value = base - 1
enum { MP_WINSTART = 0, MP_COUNTS = 4, MP_WINLEN = 5, MP_CNEXT = 6, MP_BNEXT = 7, MP_TMPLLEN = 8, MP_NALLELES = 9,
       MP_WINBASE = 10, MP_NWORDS = 11, MP_VARS = 12 };
enum { MV_START = 0, MV_END = 1, MV_FLAGS = 2, MV_GT0 = 3, MV_GT1 = 4, MV_A0 = 5, MV_A1 = 6, MV_BYTES = 8 };
enum { MA_START = 0, MA_END = 1, MA_NT = 2, MA_LEN = 3, MA_BYTES = 4 };

constexpr int kMicroFar = 110;       // every relative position >= this means "beyond the region"
constexpr int kMicroNoNext = 0x7f;
constexpr int kMicroMaxWin = 100;    // largest window / allele length a packet may hold
constexpr int kMicroNoAllele = 0xff;

// result status
enum : uint8_t { kMicroPending = 0, kMicroClean = 1, kMicroFallback = 2 };
// result record (bytes): status, flags (bit0 a tie-break looked at the prefix's emptiness, bit1 at its allele id), nSync, lastBaseAlleleId (-128 = none), iterations (u16),
// maxFrontier, pad, sync[SMAX] (relative), decision[2*KV] (calls then baseline: 0 untouched, 1 excluded, 2+k orientation k),
// weightNum[KV], weightDen[KV] (calls)
enum { MR_STATUS = 0, MR_FLAGS = 1, MR_NSYNC = 2, MR_LASTID = 3, MR_ITERS = 4, MR_MAXF = 6, MR_SYNC = 8 };
constexpr int kMicroNoLastId = -128;

struct MicroConfig {
  int H;               // haplotypes (1 or 2)
  int kind[2];         // orientor per side: [0] calls, [1] baseline (VCE_ORIENT_UNPHASED / PHASED / PHASED_INVERT / SQUASH_GT)
  int maxPaths, maxIterations, pruneNoOps, preference;
};

VCE_HD bool microOrientorSupported(int kind, int H) {
  switch (kind) {
    case VCE_ORIENT_UNPHASED:
    case VCE_ORIENT_PHASED:
    case VCE_ORIENT_PHASED_INVERT: return H == 2;
    case VCE_ORIENT_SQUASH_GT: return H == 1;
    default: return false;
  }
}

template <int KV_, int Q_, int SMAX_, int CAP_>
struct MicroTraits {
  static constexpr int KV = KV_, Q = Q_, SMAX = SMAX_, CAP = CAP_;
  static constexpr int MAXROWS = 2;                       // orientations per variant for the supported orientors
  static constexpr int HW = 5 + Q;                        // tpos, pia, lastEnd, next, qinfo, queue
  static constexpr int DW = (KV * 2 + 7) / 8;             // 2-bit decisions
  static constexpr int SW = 2 * HW + 5 + DW;              // 2 haplotypes + varIdx, varEnd, inclEnd, inclCnt, lastId + decisions
  static constexpr int WRAW = 2 * SW + 3 + SMAX;          // + nSync, cSince, bSince, sync[]
  static constexpr int W = (WRAW + 3) & ~3;               // bytes per path record
  static constexpr int NSLOTS = CAP + 1 + 1 + MAXROWS;    // frontier + head + children
  static constexpr int ROWS = 2 * KV * MAXROWS;           // orientation rows of the region
  static constexpr int ROWB = 3;                          // allele hap0, allele hap1, alleleId()
  static constexpr int PKT_MAX = 12 + 2 * KV * MV_BYTES + 4 * KV * MA_BYTES + 64;  // packet bytes a slice can hold
  static constexpr int RES_BYTES = (MR_SYNC + SMAX + 2 * KV + 2 * KV + 3) & ~3;
  // per-thread workspace (bytes): path slots, frontier order, orientation rows, row index per variant, packet copy
  static constexpr int OFF_ORDER = NSLOTS * W;
  static constexpr int OFF_ROWS = OFF_ORDER + ((CAP + 1 + 3) & ~3);
  static constexpr int OFF_VROW = OFF_ROWS + ((ROWS * ROWB + 3) & ~3);   // per variant: row0, nRows
  static constexpr int OFF_PKT = OFF_VROW + ((2 * KV * 2 + 3) & ~3);
  static constexpr int WS_RAW = OFF_PKT + ((PKT_MAX + 3) & ~3);
  // odd number of words per slice: lanes that read the same field of their own slice hit 32 different banks
  static constexpr int WS_BYTES = ((WS_RAW / 4) | 1) * 4;
  static_assert(NSLOTS <= 32, "free list is one 32-bit mask");
};

#ifndef VCE_MICROA_CAP
#define VCE_MICROA_CAP 10
#endif
typedef MicroTraits<1, 0, 3, VCE_MICROA_CAP> MicroA;   // one variant per side
typedef MicroTraits<4, 3, 6, 28> MicroB;   // up to four variants per side

template <class T>
struct MicroSearch {
  typedef int8_t E;
  enum { HP_TPOS = 0, HP_PIA = 1, HP_LASTEND = 2, HP_NEXT = 3, HP_QINFO = 4, HP_Q0 = 5 };
  enum { SD_VARIDX = 0, SD_VAREND = 1, SD_INCLEND = 2, SD_INCLCNT = 3, SD_LASTID = 4, SD_DEC0 = 5 };
  enum { PT_NSYNC = 0, PT_CSINCE = 1, PT_BSINCE = 2, PT_SYNC0 = 3 };
  enum { FL_OVERFLOW = 1, FL_WINMISS = 2, FL_ERRMOVE = 4, FL_ERRORDER = 8 };
  static constexpr int kIterCap = 2000;  // total iterations after which the region is handed to the general kernels

  uint8_t* ws;          // workspace slice (shared memory on the device)
  MicroConfig cfg;
  // region scalars
  int cCount, bCount, winLen, cNext, bNext, tmplLen, winBase, varsOff, allelesOff, ntOff;
  int n, iters, totalIters, usedPrefix, maxFrontier, flags;
  uint32_t freeMask;

  VCE_HD static int hapOff(int side, int h) { return side * T::SW + h * T::HW; }
  VCE_HD static int sideOff(int side) { return side * T::SW + 2 * T::HW; }
  VCE_HD static int pathOff() { return 2 * T::SW; }
  VCE_HD E* P(int slot) const { return reinterpret_cast<E*>(ws) + slot * T::W; }
  VCE_HD int8_t* order() const { return reinterpret_cast<int8_t*>(ws + T::OFF_ORDER); }
  VCE_HD uint8_t* rows() const { return ws + T::OFF_ROWS; }
  VCE_HD uint8_t* vrow() const { return ws + T::OFF_VROW; }
  VCE_HD const uint8_t* pk() const { return ws + T::OFF_PKT; }

  // ---- packet accessors
  VCE_HD int sideCount(int side) const { return side == 0 ? cCount : bCount; }
  VCE_HD int sideNext(int side) const { return side == 0 ? cNext : bNext; }
  VCE_HD const uint8_t* var(int side, int idx) const { return pk() + varsOff + ((side == 0 ? 0 : cCount) + idx) * MV_BYTES; }
  VCE_HD int varIndex(int side, int idx) const { return (side == 0 ? 0 : cCount) + idx; }
  VCE_HD const uint8_t* allele(int a) const { return pk() + allelesOff + a * MA_BYTES; }
  VCE_HD int base2(int i) const { return ((pk()[ntOff + (i >> 2)] >> ((i & 3) * 2)) & 3) + 1; }
  VCE_HD int alLen(int a) const { return allele(a)[MA_LEN]; }
  VCE_HD int alBase(int a, int i) const { return base2(allele(a)[MA_NT] + i); }

  VCE_HD int tmplBase(int pos) {  // HaplotypePlayback.nt :120-125, template branch
    if (pos >= tmplLen || pos < 0) return 0;
    if (pos >= winLen) { flags |= FL_WINMISS; return 0; }
    return base2(winBase + pos);
  }

  // Allele.compareTo :95-119 (reversed base comparison); a, b are local allele indices of the same side
  VCE_HD int alleleCompare(int a, int b) const {
    if (a == b) return 0;
    const uint8_t* x = allele(a);
    const uint8_t* y = allele(b);
    if (x[MA_START] != y[MA_START]) return x[MA_START] < y[MA_START] ? -1 : 1;
    if (x[MA_END] != y[MA_END]) return x[MA_END] < y[MA_END] ? -1 : 1;
    if (x[MA_LEN] != y[MA_LEN]) return x[MA_LEN] < y[MA_LEN] ? -1 : 1;
    const int len = x[MA_LEN];
    for (int i = 0; i < len; ++i) {
      const int p = base2(x[MA_NT] + i), q = base2(y[MA_NT] + i);
      if (p != q) return q < p ? -1 : 1;
    }
    return 0;
  }

  // ---- HaplotypePlayback
  VCE_HD void hapAdd(E* hp, int a) {  // :91-106
    if (a == kMicroNoAllele) return;
    const uint8_t* al = allele(a);
    if (al[MA_START] == al[MA_END] && al[MA_LEN] == 0) return;
    hp[HP_LASTEND] = (E) al[MA_END];
    if (hp[HP_NEXT] < 0) {
      hp[HP_NEXT] = (E) a;
    } else {
      const int ql = hp[HP_QINFO] & 0x3f;
      if (ql >= T::Q) { flags |= FL_OVERFLOW; return; }
      hp[HP_Q0 + (T::Q > 0 ? ql : 0)] = (E) a;
      hp[HP_QINFO] = (E) ((hp[HP_QINFO] & ~0x3f) | (ql + 1));
    }
  }
  VCE_HD bool hapIsNew(const E* hp, int a) const {  // :112
    return a == kMicroNoAllele || allele(a)[MA_START] >= hp[HP_LASTEND];
  }
  VCE_HD int hapNt(const E* hp) {  // :120-125
    if (hp[HP_PIA] < 0) return tmplBase(hp[HP_TPOS]);
    return alBase(hp[HP_NEXT], hp[HP_PIA]);
  }
  VCE_HD void hapNext(E* hp) {  // :170-214
    int tpos = hp[HP_TPOS], pia = hp[HP_PIA], nx = hp[HP_NEXT];
    if (pia < 0) {
      ++tpos;
      if (nx >= 0 && allele(nx)[MA_START] == tpos) pia = 0;
    } else {
      ++pia;
    }
    if (nx >= 0) {
      while (true) {
        if (pia != alLen(nx)) break;
        tpos = allele(nx)[MA_END];
        pia = -1;
        const int qi = hp[HP_QINFO];
        const int ql = qi & 0x3f;
        if (T::Q > 0 && ql > 0) {
          nx = hp[HP_Q0];
          for (int i = 1; i < ql; ++i) hp[HP_Q0 + i - 1] = hp[HP_Q0 + i];
          hp[HP_QINFO] = (E) ((qi & ~0x3f) | (ql - 1));
        } else {
          nx = -1;
          break;
        }
        if (tpos < allele(nx)[MA_START]) break;
        pia = 0;
        if (tpos != allele(nx)[MA_START]) { flags |= FL_ERRORDER; break; }
      }
    }
    hp[HP_TPOS] = (E) tpos; hp[HP_PIA] = (E) pia; hp[HP_NEXT] = (E) nx;
  }
  VCE_HD void hapStep(E* hp) {  // :131-141
    if (hp[HP_TPOS] < tmplLen - 1) hapNext(hp);
    else hp[HP_QINFO] |= 0x40;
  }
  VCE_HD bool hapFinished(const E* hp) const { return (hp[HP_QINFO] & 0x40) != 0; }
  VCE_HD bool hapWants(const E* hp) const {  // :151-164
    const int nx = hp[HP_NEXT];
    if (nx < 0) return true;
    const int pia = hp[HP_PIA];
    if (pia >= 0 && pia < alLen(nx) - 1) return false;
    if (T::Q > 0) {
      const int ql = hp[HP_QINFO] & 0x3f;
      for (int i = 0; i < ql; ++i) if (alLen(hp[HP_Q0 + i]) > 0) return false;
    }
    return true;
  }
  VCE_HD void hapMoveForward(E* hp, int position) {  // :258-267
    if (hp[HP_PIA] >= 0) { flags |= FL_ERRMOVE; return; }
    hp[HP_TPOS] = (E) (position - 1);
    hapNext(hp);
  }
  VCE_HD int hapCompare(const E* a, const E* b) const {  // :326-365
    const int d = a[HP_TPOS] - b[HP_TPOS];
    if (d != 0) return d;
    const int na = a[HP_NEXT], nb = b[HP_NEXT];
    if (na < 0) return nb < 0 ? 0 : -1;
    if (nb < 0) return 1;
    const int c = alleleCompare(na, nb);
    if (c != 0) return c;
    const int v = a[HP_PIA] - b[HP_PIA];
    if (v != 0) return v;
    if (T::Q > 0) {
      const int qa = a[HP_QINFO] & 0x3f, qb = b[HP_QINFO] & 0x3f;
      for (int i = 0; i < qa; ++i) {
        if (i >= qb) return 1;
        const int f = alleleCompare(a[HP_Q0 + i], b[HP_Q0 + i]);
        if (f != 0) return f;
      }
      if (qb > qa) return -1;
    }
    return 0;
  }

  // ---- HalfPath / Path
  VCE_HD int sidePos(const E* p, int side) const {  // HalfPath.getPosition :256-262
    int m = p[hapOff(side, 0) + HP_TPOS];
    if (cfg.H > 1) { const int t = p[hapOff(side, 1) + HP_TPOS]; m = t > m ? t : m; }
    return m;
  }
  VCE_HD int sideTrailing(const E* p, int side) const {  // HalfPath.trailingHaplotype :304-318
    if (cfg.H < 2) return -1;
    const int a = p[hapOff(side, 0) + HP_TPOS], b = p[hapOff(side, 1) + HP_TPOS];
    if (a == b) return -1;
    return b < a ? 1 : 0;
  }
  VCE_HD bool sideOnTemplate(const E* p, int side) const {
    for (int h = 0; h < cfg.H; ++h) if (p[hapOff(side, h) + HP_PIA] >= 0) return false;
    return true;
  }
  VCE_HD bool sideFinished(const E* p, int side) const {
    for (int h = 0; h < cfg.H; ++h) if (!hapFinished(p + hapOff(side, h))) return false;
    return true;
  }
  VCE_HD bool sideWants(const E* p, int side) const {  // HalfPath.wantsFutureVariantBases :176-178
    for (int h = 0; h < cfg.H; ++h) if (hapWants(p + hapOff(side, h))) return true;
    return false;
  }
  VCE_HD bool inSync(const E* p) const {  // Path.inSync :197-220
    if (sideTrailing(p, 0) != -1) return false;
    if (sideTrailing(p, 1) != -1) return false;
    const int cp = sidePos(p, 0);
    if (cp != sidePos(p, 1)) return false;
    if (cp < p[sideOff(0) + SD_INCLEND]) return false;
    if (cp < p[sideOff(1) + SD_INCLEND]) return false;
    if (!sideOnTemplate(p, 0)) return false;
    if (!sideOnTemplate(p, 1)) return false;
    return true;
  }
  VCE_HD bool hasNoOp(const E* p) const {  // Path.hasNoOp :227-231
    const int c = p[pathOff() + PT_CSINCE], b = p[pathOff() + PT_BSINCE];
    return (b == 0 && c > 0) || (c == 0 && b > 0);
  }
  VCE_HD bool pathMatches(const E* p) {  // Path.matches :346
    for (int h = 0; h < cfg.H; ++h) {
      const E* a = p + hapOff(0, h);
      const E* b = p + hapOff(1, h);
      if (hapFinished(a) || hapFinished(b)) continue;
      if (a[HP_PIA] < 0 && b[HP_PIA] < 0 && a[HP_TPOS] == b[HP_TPOS]) continue;  // same template base on both sides
      if (hapNt(a) != hapNt(b)) return false;
    }
    return true;
  }
  VCE_HD int pathCompare(const E* a, const E* b) const {  // Path.compareTo :351-357
    for (int side = 0; side < 2; ++side) {
      for (int h = 0; h < cfg.H; ++h) {
        const int c = hapCompare(a + hapOff(side, h), b + hapOff(side, h));
        if (c != 0) return c;
      }
    }
    return 0;
  }
  VCE_HD int lastSync(const E* p) const {  // null sync list = position 0 (absolute): below every position of the region
    const int ns = p[pathOff() + PT_NSYNC];
    return ns == 0 ? -1000 : p[pathOff() + PT_SYNC0 + ns - 1];
  }

  // PathPreference.better(first, second): true when `first` wins.  The baseline included list before this region is
  // assumed non-empty with an unknown last allele id (0); usedPrefix records that the outcome depended on it.
  VCE_HD bool firstIsBetter(const E* f, const E* s) {
    const int fc = f[sideOff(0) + SD_INCLCNT], sc = s[sideOff(0) + SD_INCLCNT];
    const int fb = f[sideOff(1) + SD_INCLCNT], sb = s[sideOff(1) + SD_INCLCNT];
    if (cfg.preference == 0) {  // MaxSumBoth.java:43-82
      const int fs = fc + fb, ss = sc + sb;
      if (fs == ss) {
        if (fb == 0 || sb == 0) usedPrefix |= 1;   // whether ANY baseline variant was included before the region matters
        int fd = f[pathOff() + PT_BSINCE] - f[pathOff() + PT_CSINCE];
        int sd = s[pathOff() + PT_BSINCE] - s[pathOff() + PT_CSINCE];
        fd = fd < 0 ? -fd : fd;
        sd = sd < 0 ? -sd : sd;
        if (fd != sd) return fd < sd;
        const int syncDelta = lastSync(f) - lastSync(s);
        if (syncDelta != 0) return syncDelta > 0;
        if ((fb == 0) != (sb == 0)) usedPrefix |= 2;   // ... and here its allele id does
        const int fid = fb > 0 ? f[sideOff(1) + SD_LASTID] : 0;
        const int sid = sb > 0 ? s[sideOff(1) + SD_LASTID] : 0;
        return fid < sid;
      }
      return fs > ss;
    }
    if (fc == sc) {  // MaxCallsMinBaseline.java:42-61
      if (fb == sb) {
        if (fb > 0) return f[sideOff(1) + SD_LASTID] < s[sideOff(1) + SD_LASTID];
      }
      return fb < sb;
    }
    return fc > sc;
  }

  VCE_HD void setDecision(E* p, int side, int idx, int code) const {
    E* d = p + sideOff(side) + SD_DEC0 + (idx >> 2);
    const int sh = (idx & 3) * 2;
    *d = (E) ((((unsigned) (uint8_t) *d) & ~(3u << sh)) | ((unsigned) code << sh));
  }
  VCE_HD int getDecision(const E* p, int side, int idx) const {
    return (((unsigned) (uint8_t) p[sideOff(side) + SD_DEC0 + (idx >> 2)]) >> ((idx & 3) * 2)) & 3;
  }

  VCE_HD void copyPath(int dst, int src) const {
    uint32_t* d = reinterpret_cast<uint32_t*>(P(dst));
    const uint32_t* s = reinterpret_cast<const uint32_t*>(P(src));
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < T::W / 4; ++i) d[i] = s[i];
  }
  VCE_HD int allocSlot() {
#if defined(__CUDA_ARCH__)
    const int s = __ffs((int) freeMask) - 1;
#else
    const int s = __builtin_ctz(freeMask);
#endif
    freeMask &= ~(1u << s);
    return s;
  }
  VCE_HD void freeSlot(int s) { freeMask |= 1u << s; }

  // The frontier is kept in DESCENDING order: order()[n-1] is the least path, so that pollFirst is --n and new
  // paths (children of the least path) are inserted near the end.
  // PathFinder.addIfBetter :321-348.  Returns false on frontier overflow.
  VCE_HD bool addIfBetter(int c) {
    int8_t* ord = order();
    const E* pc = P(c);
    int i = n - 1;  // scan from the least path upwards
    for (; i >= 0; --i) {
      const int cmp = pathCompare(P(ord[i]), pc);
      if (cmp == 0) {
        const int other = ord[i];
        if (firstIsBetter(pc, P(other))) { ord[i] = (int8_t) c; freeSlot(other); }
        else freeSlot(c);
        return true;
      }
      if (cmp > 0) break;  // existing path is greater: the new one goes right below it
    }
    if (n >= T::CAP) return false;
    // insert at physical index i + 1
    for (int j = n; j > i + 1; --j) ord[j] = ord[j - 1];
    ord[i + 1] = (int8_t) c;
    ++n;
    if (n > maxFrontier) maxFrontier = n;
    return true;
  }

  // PathFinder.nextVariant :296-309: local index, -1 none, -2 the variant that is due lies beyond this region
  VCE_HD int nextVariant(const E* p, int side) const {
    const int nextIdx = p[sideOff(side) + SD_VARIDX] + 1;
    if (nextIdx >= sideCount(side)) {
      const int ns = sideNext(side);
      if (ns == kMicroNoNext) return -1;
      return ns <= sidePos(p, side) + 1 ? -2 : -1;
    }
    const int st = var(side, nextIdx)[MV_START];
    if (st <= sidePos(p, side) + 1) return nextIdx;
    if (sideWants(p, side) && st <= p[sideOff(side) + SD_VAREND]) return nextIdx;
    return -1;
  }
  VCE_HD int futureVariantPos(const E* p, int side) const {  // :285-292
    const int nextIdx = p[sideOff(side) + SD_VARIDX] + 1;
    if (nextIdx >= sideCount(side)) {
      const int ns = sideNext(side);
      return ns == kMicroNoNext ? tmplLen - 1 : ns;
    }
    return var(side, nextIdx)[MV_START];
  }
  VCE_HD void skipToNextVariant(E* p) {  // :271-282
    const int aNext = futureVariantPos(p, 0);
    const int bNext = futureVariantPos(p, 1);
    int m = aNext < bNext ? aNext : bNext;
    m = m < tmplLen - 1 ? m : tmplLen - 1;
    const int nextPos = m - 1;
    if (nextPos > sidePos(p, 0)) {
      for (int side = 0; side < 2; ++side)
        for (int h = 0; h < cfg.H; ++h) hapMoveForward(p + hapOff(side, h), nextPos);
    }
  }
  VCE_HD void pathStep(E* p) {  // Path.step :330-341
    const int trailing = sideTrailing(p, 0);
    if (trailing == -1) {
      for (int side = 0; side < 2; ++side)
        for (int h = 0; h < cfg.H; ++h) hapStep(p + hapOff(side, h));
    } else {
      hapStep(p + hapOff(0, trailing));
      hapStep(p + hapOff(1, trailing));
    }
  }
  VCE_HD bool sideIsNew(const E* p, int side, int idx, const uint8_t* row) const {  // HalfPath.isNew :140-151
    if ((int) var(side, idx)[MV_START] >= p[sideOff(side) + SD_INCLEND]) return true;
    for (int h = 0; h < cfg.H; ++h) if (!hapIsNew(p + hapOff(side, h), row[h])) return false;
    return true;
  }

  // ---- K1 in place: Orientor.orientations for the variant (UnphasedOrientor :74-99, PhasedOrientor :131-147,
  // SQUASH_GT :166-187); rows are (allele hap0, allele hap1, alleleId()).
  VCE_HD int orient(int side, const uint8_t* v, uint8_t* out) const {
    const int kind = cfg.kind[side];
    const int g0 = (int8_t) v[MV_GT0], g1 = (int8_t) v[MV_GT1];
    const int a0 = v[MV_A0], a1 = v[MV_A1];
    const bool phased = (v[MV_FLAGS] & 1) != 0;
    if (kind == VCE_ORIENT_SQUASH_GT) {
      const int ploidy = (v[MV_FLAGS] >> 1) & 3;
      int lo = g0, hi = ploidy > 1 ? g1 : g0, alo = a0, ahi = ploidy > 1 ? a1 : a0;
      if (hi < lo) { int t = lo; lo = hi; hi = t; t = alo; alo = ahi; ahi = t; }
      int nr = 0, prev = 0;
      if (lo > prev && alo != kMicroNoAllele) { out[0] = (uint8_t) alo; out[1] = kMicroNoAllele; out[2] = (uint8_t) lo; ++nr; prev = lo; }
      if (hi > prev && ahi != kMicroNoAllele) {
        out[nr * 3] = (uint8_t) ahi; out[nr * 3 + 1] = kMicroNoAllele; out[nr * 3 + 2] = (uint8_t) hi; ++nr;
      }
      return nr;
    }
    if ((kind == VCE_ORIENT_PHASED || kind == VCE_ORIENT_PHASED_INVERT) && phased) {
      if (kind == VCE_ORIENT_PHASED_INVERT) { out[0] = (uint8_t) a1; out[1] = (uint8_t) a0; out[2] = (uint8_t) g1; }
      else { out[0] = (uint8_t) a0; out[1] = (uint8_t) a1; out[2] = (uint8_t) g0; }
      return 1;
    }
    out[0] = (uint8_t) a0; out[1] = (uint8_t) a1; out[2] = (uint8_t) g0;
    if (g0 == g1) return 1;
    out[3] = (uint8_t) a1; out[4] = (uint8_t) a0; out[5] = (uint8_t) g1;
    return 2;
  }

  // ---- region set-up: the packet has already been copied to ws + OFF_PKT
  VCE_HD void begin() {
    const uint8_t* k = pk();
    cCount = k[MP_COUNTS] & 0xf;
    bCount = k[MP_COUNTS] >> 4;
    winLen = k[MP_WINLEN];
    cNext = k[MP_CNEXT];
    bNext = k[MP_BNEXT];
    tmplLen = k[MP_TMPLLEN];
    winBase = k[MP_WINBASE];
    varsOff = MP_VARS;
    allelesOff = varsOff + (cCount + bCount) * MV_BYTES;
    ntOff = allelesOff + k[MP_NALLELES] * MA_BYTES;
    int r = 0;
    for (int side = 0; side < 2; ++side) {
      for (int i = 0; i < sideCount(side); ++i) {
        const int vi = varIndex(side, i);
        const int nr = orient(side, var(side, i), rows() + r * T::ROWB);
        vrow()[vi * 2] = (uint8_t) r;
        vrow()[vi * 2 + 1] = (uint8_t) nr;
        r += nr;
      }
    }
    n = 1; iters = 0; totalIters = 0; usedPrefix = 0; maxFrontier = 1; flags = 0;
    freeMask = T::NSLOTS >= 32 ? 0xffffffffu : ((1u << T::NSLOTS) - 1u);
    const int s0 = allocSlot();
    E* p = P(s0);
    for (int i = 0; i < T::W; ++i) p[i] = 0;
    for (int side = 0; side < 2; ++side) {
      for (int h = 0; h < 2; ++h) {
        E* hp = p + hapOff(side, h);
        hp[HP_TPOS] = 0;       // the fresh in-sync path stands at relative position 0
        hp[HP_PIA] = -1;
        hp[HP_LASTEND] = -1;
        hp[HP_NEXT] = -1;
      }
      p[sideOff(side) + SD_VARIDX] = -1;
      p[sideOff(side) + SD_VAREND] = -1;   // absolute 0 lies at or below relative -1 (regions at a sequence start are not micro)
      p[sideOff(side) + SD_INCLEND] = -1;
    }
    order()[0] = (int8_t) s0;
  }

  // One iteration of PathFinder.bestPath :143-210.  Returns kMicroPending, kMicroClean (resultSlot valid) or kMicroFallback.
  VCE_HD int iterate(int& resultSlot) {
    if (n <= 0) return kMicroFallback;  // every path died although more variants follow: the merged run decides
    ++iters;
    ++totalIters;
    if (totalIters > kIterCap) return kMicroFallback;
    const int head = order()[--n];  // pollFirst
    if (n == 0) iters = 0;          // :151-161
    else if (n > cfg.maxPaths || iters > cfg.maxIterations) return kMicroFallback;  // too complex: exact handling elsewhere
    E* hp = P(head);

    int side = 0;
    int vidx = nextVariant(hp, 0);  // :170 calls first
    if (vidx == -1) { side = 1; vidx = nextVariant(hp, 1); }  // :173 then baseline
    if (vidx == -2) {
      // The next variant belongs to the following region: exact only if this is the single surviving path, in sync,
      // right before that variant, with nothing of this region left.
      const int boundary = (cNext < bNext ? cNext : bNext) - 1;
      if (n == 0 && inSync(hp) && sidePos(hp, 0) == boundary && hp[sideOff(0) + SD_VARIDX] + 1 >= cCount
          && hp[sideOff(1) + SD_VARIDX] + 1 >= bCount) {
        resultSlot = head;
        return kMicroClean;
      }
      return kMicroFallback;
    }
    if (vidx >= 0) {  // enqueueVariant :231-247 -> Path.addVariant :288-323
      if (inSync(hp)) {
        E* pb = hp + pathOff();
        const int syncPos = sidePos(hp, 0);
        const int ns = pb[PT_NSYNC];
        if (!(ns > 0 && pb[PT_SYNC0 + ns - 1] == syncPos)) {
          if (ns >= T::SMAX) return kMicroFallback;
          pb[PT_SYNC0 + ns] = (E) syncPos;
          pb[PT_NSYNC] = (E) (ns + 1);
        }
        pb[PT_CSINCE] = 0;
        pb[PT_BSINCE] = 0;
      }
      const uint8_t* v = var(side, vidx);
      const int vend = v[MV_END];
      const int vi = varIndex(side, vidx);
      const int row0 = vrow()[vi * 2], nOr = vrow()[vi * 2 + 1];
      int child[1 + T::MAXROWS];
      int nChild = 0;
      {  // HalfPath.exclude :153-158
        const int cs = allocSlot();
        copyPath(cs, head);
        E* sb = P(cs) + sideOff(side);
        sb[SD_VAREND] = (E) (sb[SD_VAREND] > vend ? sb[SD_VAREND] : vend);
        sb[SD_VARIDX] = (E) vidx;
        setDecision(P(cs), side, vidx, 1);
        child[nChild++] = cs;
      }
      for (int k = 0; k < nOr; ++k) {
        const uint8_t* row = rows() + (row0 + k) * T::ROWB;
        if (!sideIsNew(hp, side, vidx, row)) continue;
        const int cs = allocSlot();  // HalfPath.include :160-170, Path.include :267-275
        copyPath(cs, head);
        E* cp = P(cs);
        E* sb = cp + sideOff(side);
        sb[SD_VAREND] = (E) (sb[SD_VAREND] > vend ? sb[SD_VAREND] : vend);
        sb[SD_VARIDX] = (E) vidx;
        sb[SD_INCLEND] = (E) (sb[SD_INCLEND] > vend ? sb[SD_INCLEND] : vend);
        sb[SD_INCLCNT] += 1;
        sb[SD_LASTID] = (E) (int8_t) row[2];
        setDecision(cp, side, vidx, 2 + k);
        for (int h = 0; h < cfg.H; ++h) hapAdd(cp + hapOff(side, h), row[h]);
        cp[pathOff() + (side == 0 ? PT_CSINCE : PT_BSINCE)] += 1;
        child[nChild++] = cs;
      }
      if (flags) return kMicroFallback;
      freeSlot(head);
      for (int k = 0; k < nChild; ++k) if (!addIfBetter(child[k])) return kMicroFallback;
      return kMicroPending;
    }

    pathStep(hp);  // :177
    bool alive = true;
    if (inSync(hp)) {  // :182-190
      if (cfg.pruneNoOps && hasNoOp(hp)) alive = false;
      else skipToNextVariant(hp);
    }
    if (alive && pathMatches(hp)) {  // :192
      if (flags) return kMicroFallback;
      if (sideFinished(hp, 0) && sideFinished(hp, 1)) return kMicroFallback;  // finishing is the last region's business
      if (!addIfBetter(head)) return kMicroFallback;
      return kMicroPending;
    }
    if (flags) return kMicroFallback;
    freeSlot(head);
    return kMicroPending;
  }

  // ---- K3 epilogue: decisions, sync points, weights (Path.calculateWeights E/Path.java:385-453 restricted to the region)
  VCE_HD void writeResult(int status, int slot, uint8_t* out) const {
    for (int i = 0; i < T::RES_BYTES; ++i) out[i] = 0;
    out[MR_STATUS] = (uint8_t) status;
    out[MR_FLAGS] = (uint8_t) (usedPrefix & 3);
    const int it = totalIters > 0xffff ? 0xffff : totalIters;
    out[MR_ITERS] = (uint8_t) (it & 0xff);
    out[MR_ITERS + 1] = (uint8_t) (it >> 8);
    out[MR_MAXF] = (uint8_t) maxFrontier;
    out[MR_LASTID] = (uint8_t) (int8_t) kMicroNoLastId;
    if (status != kMicroClean) return;
    const E* p = P(slot);
    const int ns = p[pathOff() + PT_NSYNC];
    const E* sy = p + pathOff() + PT_SYNC0;
    out[MR_NSYNC] = (uint8_t) ns;
    for (int i = 0; i < ns; ++i) out[MR_SYNC + i] = (uint8_t) sy[i];
    uint8_t* dec = out + MR_SYNC + T::SMAX;
    for (int i = 0; i < cCount; ++i) dec[i] = (uint8_t) getDecision(p, 0, i);
    for (int i = 0; i < bCount; ++i) dec[T::KV + i] = (uint8_t) getDecision(p, 1, i);
    uint8_t* wn = dec + 2 * T::KV;
    uint8_t* wd = wn + T::KV;
    for (int i = 0; i < cCount; ++i) {
      int num = 0, den = 1;
      if (dec[i] >= 2) {
        const int st = var(0, i)[MV_START];
        int k = 0;
        while (k < ns && sy[k] < st) ++k;
        const int lo = k == 0 ? -1000 : sy[k - 1];  // interval = starts in (lo, hi]
        const int hi = k < ns ? sy[k] : 1000;
        int calledTP = 0, baseTP = 0, baseInside = 0;
        for (int j = 0; j < cCount; ++j) {
          if (dec[j] >= 2) { const int sj = var(0, j)[MV_START]; if (sj > lo && sj <= hi) ++calledTP; }
        }
        for (int j = 0; j < bCount; ++j) {
          if (dec[T::KV + j] >= 2) {
            const int sj = var(1, j)[MV_START];
            if (sj > lo && sj <= hi) { ++baseTP; if (!(var(1, j)[MV_FLAGS] & 0x10)) ++baseInside; }
          }
        }
        if (baseTP != 0) { num = baseInside; den = calledTP; }
      }
      wn[i] = (uint8_t) num;
      wd[i] = (uint8_t) den;
    }
    if (p[sideOff(1) + SD_INCLCNT] > 0) out[MR_LASTID] = (uint8_t) p[sideOff(1) + SD_LASTID];
  }

  // Whole region on one thread (host emulation; the device kernel interleaves iterations of different regions).
  VCE_HD void runRegion(uint8_t* out) {
    begin();
    int slot = -1;
    int st;
    do { st = iterate(slot); } while (st == kMicroPending);
    writeResult(st, slot, out);
  }
};

}  // namespace vce
#endif
