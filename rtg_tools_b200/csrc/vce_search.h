// vce_search.h -- the bounded best-first haplotype search for ONE region, run by one warp.
//
// Behavioural contract (bit-exact): PathFinder.bestPath and everything it calls
//   E/PathFinder.java:129-348, E/Path.java:190-357, E/HalfPath.java:140-326,
//   E/HaplotypePlayback.java:91-365, E/Allele.java:95-119, E/MaxSumBoth.java:43-82,
//   E/MaxCallsMinBaseline.java:42-61, Path.calculateWeights E/Path.java:385-453
// (E/ = src/com/rtg/vcf/eval/ of the reference).  This is a re-design, not a port:
//   * a search node ("path") is a fixed-size int32 record, the persistent included/excluded lists of
//     the reference become an inline decision nibble per variant, the sync list an inline array;
//   * the ordered frontier (java.util.TreeSet) is a sorted index array; insert position and the
//     "equal path already present" test are found by all lanes comparing in parallel and a ballot;
//   * child paths are created by a cooperative warp copy, one lane then specialises each child;
//   * the region starts from a synthetic in-sync path at the gap before its first variant and must end
//     with a single in-sync path at the gap after its last one, otherwise it reports SPILL and the
//     host merges it with its neighbour.
//
// The code is written against a "warp policy" W with L lanes so that the very same source runs
// on the device (L = 32, __ballot_sync/__syncwarp) and, for host-logic tests only, on the CPU with
// L = 1.  The product library only ever instantiates the device policy.
#ifndef VCE_SEARCH_H_
#define VCE_SEARCH_H_

#include "vce_device.h"

namespace vce {

// ---- path record layout (int32 words) ------------------------------------------------------------
// hap   : [0]=templatePos [1]=posInAllele(-1 = on template) [2]=lastAlleleEnd [3]=nextAllele slot(-1 = none)
//         [4]=qlen | finished<<16   [5..5+Q) = pending allele slots (FIFO)
// side  : H haps, then [0]=variantIndex(local) [1]=variantEnd [2]=includedVariantEnd [3]=includedCount
//         [4]=alleleId() of last included   [5..5+DW) = decision nibbles/bytes
// path  : calls side, baseline side, then [0]=nSync [1]=cSinceSync [2]=bSinceSync [3]=spare [4..4+SMAX) = sync positions
// the four words every comparison reads come first: one 16-byte load per haplotype (records are 16-byte aligned)
enum { HP_TPOS = 0, HP_NEXT = 1, HP_PIA = 2, HP_QINFO = 3, HP_LASTEND = 4, HP_Q0 = 5 };
struct alignas(16) Quad { int32_t x, y, z, w; };
enum { SD_VARIDX = 0, SD_VAREND = 1, SD_INCLEND = 2, SD_INCLCNT = 3, SD_LASTID = 4, SD_DEC0 = 5 };
enum { PT_NSYNC = 0, PT_CSINCE = 1, PT_BSINCE = 2, PT_SPARE = 3, PT_SYNC0 = 4 };

struct PathLayout {
  int H, Q, DW, SMAX, decBits;
  int HW, SW, W;
  VCE_HD void init(int h, int q, int nVarMaxSide, int smax, int bits) {
    H = h; Q = q; decBits = bits;
    DW = (nVarMaxSide * bits + 31) / 32;
    if (DW < 1) DW = 1;
    SMAX = smax;
    HW = (HP_Q0 + Q + 3) & ~3;
    SW = (H * HW + SD_DEC0 + DW + 3) & ~3;
    W = (2 * SW + PT_SYNC0 + SMAX + 3) & ~3;
  }
  VCE_HD int hap(int side, int h) const { return side * SW + h * HW; }
  VCE_HD int sideBase(int side) const { return side * SW + H * HW; }
  VCE_HD int pathBase() const { return 2 * SW; }
};

struct SearchConfig {
  int maxPaths, maxIterations, pruneNoOps, preference;  // PathFinder.Config
};

// The search is a chain of dependent, data-dependent steps: unrolled loops only grow the instruction stream, and the
// kernels are bound by instruction fetch (every resident warp walks a different part of it), so loops stay rolled.
#if defined(__CUDA_ARCH__)
#define VCE_NOUNROLL _Pragma("unroll 1")
#else
#define VCE_NOUNROLL
#endif

// ---- warp policies ----------------------------------------------------------------------------------
#if defined(__CUDACC__)
struct DeviceWarp {
  static constexpr int L = 32;
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  __device__ __forceinline__ int bcast(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
};
#endif
struct HostLane {  // one lane; host-logic tests only
  static constexpr int L = 1;
  int lane() const { return 0; }
  void sync() const {}
  unsigned ballot(bool p) const { return p ? 1u : 0u; }
  int bcast(int v, int) const { return v; }
};

VCE_HD int vpopc(unsigned x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
VCE_HD int vffs(unsigned x) {  // index of lowest set bit, x != 0
#if defined(__CUDA_ARCH__)
  return __ffs(x) - 1;
#else
  return __builtin_ctz(x);
#endif
}

// ---- the search -------------------------------------------------------------------------------------
// Warp-uniform, read-only state of one region's search.  On the device it lives in SHARED memory (one copy per warp,
// broadcast reads): kept in the thread-local frame it would be spilled to local memory, and with the L1 carve-out
// given to shared memory every access of the serial part would then go to L2.
struct RegionConst {
  RegionDesc R;
  SearchConfig cfg;
  PathLayout L;
  SideArrays S[2];       // 0 = calls, 1 = baseline
};

template <class WP, bool CHUNKED>
struct RegionSearch {
  WP w;
  RegionConst* K;        // see RegionConst
  RegionDesc& R;
  SearchConfig& cfg;
  SideArrays* S;         // [2]
  PathLayout& L;
  const uint8_t* win;    // reference window bases for this region (0..4 per byte)
  VCE_HD explicit RegionSearch(RegionConst* k) : K(k), R(k->R), cfg(k->cfg), S(k->S), L(k->L) {}
  int32_t* slots;        // (nSlots + 2) path records: [nSlots] = last-sync path, [nSlots+1] = best finished path
  int32_t* order;        // sorted frontier as a deque: logical element i lives at order[oh + i]; [0] is the least path
  int32_t* freeList;     // stack of free slot ids
  int32_t* ctl;          // 16 + 2*maxChildren words of lane-0 -> warp mailbox
  int cap;               // frontier capacity of this kernel (overflow beyond)
  int ordCap;            // physical length of order[] (>= cap + 1)
  // Large frontiers (large-capacity kernel): the sorted frontier is a two-level structure inside order[] -- a directory
  // of blocks of up to kBlk slot ids each -- so that an insert moves at most one block (java.util.TreeSet is O(log n);
  // a flat sorted array would move half the frontier per insert).  See the frontier* / chunk* members.
  static constexpr bool chunked = CHUNKED;
  static constexpr int kBlk = 64;          // slot ids per block
  static constexpr int kBlkStride = kBlk + 2;  // [0] = start, [1] = count, then the ids
  int nbMax = 0, nb = 0, dh = 0, nbFree = 0;  // blocks available / in use, directory head offset, free blocks
  VCE_HD static int chunkWords(int capacity) {  // words of order[] the chunked frontier needs for `capacity` paths
    const int nbm = capacity / (kBlk / 2) + 8;
    return 2 * nbm + nbm + nbm * kBlkStride;
  }
  VCE_HD int32_t* cdir() const { return order; }                       // [2*nbMax] directory (deque, head offset dh)
  VCE_HD int32_t* cfree() const { return order + 2 * nbMax; }          // [nbMax] free block stack
  VCE_HD int32_t* cblk(int b) const { return order + 3 * nbMax + b * kBlkStride; }
  int nSlots;
  int maxChildren;
  int oh;                // deque head offset (warp-uniform)

  // warp-uniform scalars
  int n, nFree, iters, curMaxPos, lastSyncPos, bestValid, usedPrefix, nWarn, maxFrontier;
  int totalIters;
  WarnRec* warnOut; int warnCap; int regionId;

  enum { CTL_ACT = 0, CTL_SIDE = 1, CTL_VIDX = 2, CTL_FLAGS = 3, CTL_AUX = 4, CTL_CHILD0 = 16 };
  VCE_HD int32_t& childValid(int k) const { return ctl[CTL_CHILD0 + k]; }
  VCE_HD int32_t& childSlot(int k) const { return ctl[CTL_CHILD0 + maxChildren + k]; }
  VCE_HD int32_t& ORD(int i) const { return order[oh + i]; }
  enum { FL_OVERFLOW = 1, FL_WINMISS = 2, FL_ERRMOVE = 4, FL_ERRORDER = 8 };
  enum { ACT_ENQUEUE = 1, ACT_STEP_DROP = 2, ACT_STEP_REINSERT = 3, ACT_STEP_FINISH = 4, ACT_BOUNDARY_CLEAN = 5,
         ACT_BOUNDARY_SPILL = 6 };

  VCE_HD int32_t* P(int slot) const { return slots + (size_t) slot * L.W; }
  VCE_HD int slotLS() const { return nSlots; }
  VCE_HD int slotBest() const { return nSlots + 1; }
  VCE_HD void flag(int f) { flagS(ctl, f); }
  VCE_HD static void flagS(int32_t* ctlp, int f) {
#if defined(__CUDA_ARCH__)
    atomicOr(ctlp + CTL_FLAGS, f);
#else
    ctlp[CTL_FLAGS] |= f;
#endif
  }

  // ---------------------------------------------------------------- template / allele access
  VCE_HD int tmplBase(int pos) {  // HaplotypePlayback.nt :120-125 template branch
    if (pos >= R.tmplLen || pos < 0) return 0;
    const int o = pos - R.winStart;
    if (o < 0 || o >= R.winLen) { flag(FL_WINMISS); return 0; }
    return win[o];
  }
  VCE_HD int alLen(int side, int slot) const { return S[side].aNtOff[slot + 1] - S[side].aNtOff[slot]; }

  // Allele.compareTo :95-119 (reversed base comparison)
  VCE_HD int alleleCompare(int side, int a, int b) const { return a == b ? 0 : alleleCompareS(S + side, a, b); }
  VCE_HDN static int alleleCompareS(const SideArrays* sp, int a, int b) {
    if (a == b) return 0;
    const SideArrays& s = *sp;
    int x = s.aStart[a], y = s.aStart[b];
    if (x != y) return x < y ? -1 : 1;
    x = s.aEnd[a]; y = s.aEnd[b];
    if (x != y) return x < y ? -1 : 1;
    const int oa = s.aNtOff[a], ob = s.aNtOff[b];
    x = s.aNtOff[a + 1] - oa; y = s.aNtOff[b + 1] - ob;
    if (x != y) return x < y ? -1 : 1;
    VCE_NOUNROLL
    for (int i = 0; i < x; ++i) {
      const int p = s.nt[oa + i], q = s.nt[ob + i];
      if (p != q) return q < p ? -1 : 1;
    }
    return 0;
  }

  // ---------------------------------------------------------------- HaplotypePlayback
  VCE_HD void hapAdd(int32_t* hp, int side, int slot) {  // :91-106
    if (slot < 0) return;
    const SideArrays& s = S[side];
    const int st = s.aStart[slot], en = s.aEnd[slot];
    if (st == en && alLen(side, slot) == 0) return;
    hp[HP_LASTEND] = en;
    if (hp[HP_NEXT] < 0) {
      hp[HP_NEXT] = slot;
    } else {
      const int ql = hp[HP_QINFO] & 0xffff;
      if (ql >= L.Q) { flag(FL_OVERFLOW); return; }
      hp[HP_Q0 + ql] = slot;
      hp[HP_QINFO] = (hp[HP_QINFO] & ~0xffff) | (ql + 1);
    }
  }
  VCE_HD bool hapIsNew(const int32_t* hp, int side, int slot) const {  // :112
    return slot < 0 || S[side].aStart[slot] >= hp[HP_LASTEND];
  }
  VCE_HD int hapNt(const int32_t* hp, int side) {  // :120-125
    if (hp[HP_PIA] < 0) return tmplBase(hp[HP_TPOS]);
    return S[side].nt[S[side].aNtOff[hp[HP_NEXT]] + hp[HP_PIA]];
  }
  VCE_HD void hapNext(int32_t* hp, int side) { hapNextS(hp, S + side, ctl); }  // :170-214
  VCE_HDN static void hapNextS(int32_t* hp, const SideArrays* sp, int32_t* ctlp) {
    int tpos = hp[HP_TPOS], pia = hp[HP_PIA], nx = hp[HP_NEXT];
    const SideArrays& s = *sp;
    if (pia < 0) {
      ++tpos;
      if (nx >= 0 && s.aStart[nx] == tpos) pia = 0;
    } else {
      ++pia;
    }
    if (nx >= 0) {
      while (true) {
        if (pia != s.aNtOff[nx + 1] - s.aNtOff[nx]) break;
        tpos = s.aEnd[nx];
        pia = -1;
        const int qi = hp[HP_QINFO];
        const int ql = qi & 0xffff;
        if (ql > 0) {
          nx = hp[HP_Q0];
          VCE_NOUNROLL
          for (int i = 1; i < ql; ++i) hp[HP_Q0 + i - 1] = hp[HP_Q0 + i];
          hp[HP_QINFO] = (qi & ~0xffff) | (ql - 1);
        } else {
          nx = -1;
          break;
        }
        if (tpos < s.aStart[nx]) break;
        pia = 0;
        if (tpos != s.aStart[nx]) { flagS(ctlp, FL_ERRORDER); break; }
      }
    }
    hp[HP_TPOS] = tpos; hp[HP_PIA] = pia; hp[HP_NEXT] = nx;
  }
  VCE_HD void hapStep(int32_t* hp, int side) {  // :131-141
    if (hp[HP_TPOS] < R.tmplLen - 1) hapNext(hp, side);
    else hp[HP_QINFO] |= 0x10000;
  }
  VCE_HD bool hapFinished(const int32_t* hp) const { return (hp[HP_QINFO] & 0x10000) != 0; }
  VCE_HD bool hapWants(const int32_t* hp, int side) const {  // :151-164
    const int nx = hp[HP_NEXT];
    if (nx < 0) return true;
    const int pia = hp[HP_PIA];
    if (pia >= 0 && pia < alLen(side, nx) - 1) return false;
    const int ql = hp[HP_QINFO] & 0xffff;
    VCE_NOUNROLL
    for (int i = 0; i < ql; ++i) if (alLen(side, hp[HP_Q0 + i]) > 0) return false;
    return true;
  }
  VCE_HD void hapMoveForward(int32_t* hp, int side, int position) {  // :258-267
    if (hp[HP_PIA] >= 0) { flag(FL_ERRMOVE); return; }
    hp[HP_TPOS] = position - 1;
    hapNext(hp, side);
  }
  VCE_HD int hapCompare(const int32_t* a, const int32_t* b, int side) const { return hapCompareS(a, b, S + side); }
  VCE_HD static int hapCompareS(const int32_t* a, const int32_t* b, const SideArrays* sp) {  // :326-365
    const Quad A = *reinterpret_cast<const Quad*>(a), B = *reinterpret_cast<const Quad*>(b);  // TPOS, NEXT, PIA, QINFO
    const int d = A.x - B.x;
    if (d != 0) return d;
    const int na = A.y, nb = B.y;
    if (na < 0) return nb < 0 ? 0 : -1;
    if (nb < 0) return 1;
    const int c = (na == nb ? 0 : alleleCompareS(sp, na, nb));
    if (c != 0) return c;
    const int v = A.z - B.z;
    if (v != 0) return v;
    const int qa = A.w & 0xffff, qb = B.w & 0xffff;
    VCE_NOUNROLL
    for (int i = 0; i < qa; ++i) {
      if (i >= qb) return 1;
      const int f = alleleCompareS(sp, a[HP_Q0 + i], b[HP_Q0 + i]);
      if (f != 0) return f;
    }
    if (qb > qa) return -1;
    return 0;
  }

  // ---------------------------------------------------------------- HalfPath / Path (read side)
  VCE_HD int sidePos(const int32_t* p, int side) const {  // HalfPath.getPosition :256-262
    int m = p[L.hap(side, 0) + HP_TPOS];
    VCE_NOUNROLL
    for (int h = 1; h < L.H; ++h) { const int t = p[L.hap(side, h) + HP_TPOS]; m = t > m ? t : m; }
    return m;
  }
  VCE_HD int sideTrailing(const int32_t* p, int side) const {  // HalfPath.trailingHaplotype :304-318
    int minh = 0;
    int minp = p[L.hap(side, 0) + HP_TPOS];
    int maxp = minp;
    VCE_NOUNROLL
    for (int h = 1; h < L.H; ++h) {
      const int t = p[L.hap(side, h) + HP_TPOS];
      if (t < minp) { minp = t; minh = h; }
      else if (t > minp) { maxp = t; }
    }
    return minp == maxp ? -1 : minh;
  }
  VCE_HD bool sideOnTemplate(const int32_t* p, int side) const {
    VCE_NOUNROLL
    for (int h = 0; h < L.H; ++h) if (p[L.hap(side, h) + HP_PIA] >= 0) return false;
    return true;
  }
  VCE_HD bool sideFinished(const int32_t* p, int side) const {
    VCE_NOUNROLL
    for (int h = 0; h < L.H; ++h) if (!hapFinished(p + L.hap(side, h))) return false;
    return true;
  }
  VCE_HD bool sideWants(const int32_t* p, int side) const {  // HalfPath.wantsFutureVariantBases :176-178
    VCE_NOUNROLL
    for (int h = 0; h < L.H; ++h) if (hapWants(p + L.hap(side, h), side)) return true;
    return false;
  }
  VCE_HD bool inSync(const int32_t* p) const {  // Path.inSync :197-220
    if (sideTrailing(p, 0) != -1) return false;
    if (sideTrailing(p, 1) != -1) return false;
    const int cp = sidePos(p, 0);
    if (cp != sidePos(p, 1)) return false;
    if (cp < p[L.sideBase(0) + SD_INCLEND]) return false;
    if (cp < p[L.sideBase(1) + SD_INCLEND]) return false;
    if (!sideOnTemplate(p, 0)) return false;
    if (!sideOnTemplate(p, 1)) return false;
    return true;
  }
  VCE_HD bool hasNoOp(const int32_t* p) const {  // Path.hasNoOp :227-231
    const int c = p[L.pathBase() + PT_CSINCE], b = p[L.pathBase() + PT_BSINCE];
    return (b == 0 && c > 0) || (c == 0 && b > 0);
  }
  VCE_HD bool pathMatches(const int32_t* p) {  // Path.matches :346 -> HalfPath.matches :278-286 -> HaplotypePlayback.matches :316
    VCE_NOUNROLL
    for (int h = 0; h < L.H; ++h) {
      const int32_t* a = p + L.hap(0, h);
      const int32_t* b = p + L.hap(1, h);
      if (hapFinished(a) || hapFinished(b)) continue;
      // both on the template at the same position read the same base: no reference access needed
      if (a[HP_PIA] < 0 && b[HP_PIA] < 0 && a[HP_TPOS] == b[HP_TPOS]) continue;
      if (hapNt(a, 0) != hapNt(b, 1)) return false;
    }
    return true;
  }
  VCE_HD int pathCompare(const int32_t* a, const int32_t* b) const { return pathCompareS(a, b, K); }
  VCE_HDN static int pathCompareS(const int32_t* a, const int32_t* b, const RegionConst* k) {  // Path.compareTo :351-357
    const PathLayout& pl = k->L;
    VCE_NOUNROLL
    for (int side = 0; side < 2; ++side) {
      VCE_NOUNROLL
      for (int h = 0; h < pl.H; ++h) {
        const int c = hapCompareS(a + pl.hap(side, h), b + pl.hap(side, h), k->S + side);
        if (c != 0) return c;
      }
    }
    return 0;
  }
  VCE_HD int lastSync(const int32_t* p) const {
    const int ns = p[L.pathBase() + PT_NSYNC];
    return ns == 0 ? 0 : p[L.pathBase() + PT_SYNC0 + ns - 1];
  }

  // PathPreference.better(first, second): returns true when `first` wins.
  VCE_HD bool firstIsBetter(const int32_t* f, const int32_t* s) {
    const int fc = f[L.sideBase(0) + SD_INCLCNT], sc = s[L.sideBase(0) + SD_INCLCNT];
    const int fb = f[L.sideBase(1) + SD_INCLCNT], sb = s[L.sideBase(1) + SD_INCLCNT];
    const bool prefixNonEmpty = R.prefixAlleleId != kPrefixEmpty;
    if (cfg.preference == 0) {  // MaxSumBoth.java:43-82
      const int fs = fc + fb, ss = sc + sb;
      if (fs == ss) {
        if (fb == 0 || sb == 0) usedPrefix |= 1;   // bit0: the prefix's emptiness matters, bit1: its allele id does
        const bool fiNonNull = prefixNonEmpty || fb > 0;
        const bool siNonNull = prefixNonEmpty || sb > 0;
        if (fiNonNull && siNonNull) {
          int fd = f[L.pathBase() + PT_BSINCE] - f[L.pathBase() + PT_CSINCE];
          int sd = s[L.pathBase() + PT_BSINCE] - s[L.pathBase() + PT_CSINCE];
          fd = fd < 0 ? -fd : fd;
          sd = sd < 0 ? -sd : sd;
          if (fd != sd) return fd < sd;
          const int syncDelta = lastSync(f) - lastSync(s);
          if (syncDelta != 0) return syncDelta > 0;
          if ((fb == 0) != (sb == 0)) usedPrefix |= 2;
          const int fid = fb > 0 ? f[L.sideBase(1) + SD_LASTID] : R.prefixAlleleId;
          const int sid = sb > 0 ? s[L.sideBase(1) + SD_LASTID] : R.prefixAlleleId;
          return fid < sid;
        }
      }
      return fs > ss;
    }
    // MaxCallsMinBaseline.java:42-61 (never depends on the prefix: equal local sizes <=> equal total sizes,
    // and with both local lists empty the outcome is `second` whether or not the prefix is empty)
    if (fc == sc) {
      if (fb == sb) {
        if (fb > 0) return f[L.sideBase(1) + SD_LASTID] < s[L.sideBase(1) + SD_LASTID];
      }
      return fb < sb;
    }
    return fc > sc;
  }

  // ---------------------------------------------------------------- decisions
  VCE_HD void setDecision(int32_t* p, int side, int idx, int code) const {
    int32_t* d = p + L.sideBase(side) + SD_DEC0;
    const int bit = idx * L.decBits;
    const unsigned mask = (L.decBits == 4 ? 0xfu : 0xffu) << (bit & 31);
    d[bit >> 5] = (int32_t) ((((unsigned) d[bit >> 5]) & ~mask) | (((unsigned) code) << (bit & 31)));
  }
  VCE_HD int getDecision(const int32_t* p, int side, int idx) const {
    const int32_t* d = p + L.sideBase(side) + SD_DEC0;
    const int bit = idx * L.decBits;
    return (int) ((((unsigned) d[bit >> 5]) >> (bit & 31)) & (L.decBits == 4 ? 0xfu : 0xffu));
  }

  // ---------------------------------------------------------------- warp helpers
  VCE_HD void copyPath(int dst, int src) {  // cooperative
    int32_t* d = P(dst);
    const int32_t* s = P(src);
    Quad* d4 = reinterpret_cast<Quad*>(d);
    const Quad* s4 = reinterpret_cast<const Quad*>(s);
    VCE_NOUNROLL
    for (int i = w.lane(); i < (L.W >> 2); i += WP::L) d4[i] = s4[i];
  }
  VCE_HD void pushFree(int slot) {  // lane 0 writes; callers sync before the next pop
    if (w.lane() == 0) freeList[nFree] = slot;
    ++nFree;
  }

  // Position of `c` in the sorted frontier: pos = #paths < c, eqPos = index of the path equal to c (or -1).
  VCE_HD void findPos(int c, int& pos, int& eqPos) {
    const int32_t* pc = P(c);
    pos = 0;
    eqPos = -1;
    int lo = 0;
    // The new path is almost always near the front (children of the least path): scan from the front in
    // warp-wide chunks; every lane compares one resident path against the new one.
    while (lo < n) {
      const int i = lo + w.lane();
      int cmp = 1;  // "existing > new" for lanes beyond the end
      if (i < n) cmp = pathCompare(P(ORD(i)), pc);
      const unsigned less = w.ballot(cmp < 0);
      const unsigned eq = w.ballot(cmp == 0);
      if (eq) { eqPos = lo + vffs(eq); pos = eqPos; return; }
      const int k = vpopc(less);
      pos = lo + k;
      if (k < WP::L) return;
      lo += WP::L;
      if (lo < n && n - lo > 8 * WP::L) {
        // long frontier: WP::L-ary search over the remainder instead of a linear scan
        int a = lo, b = n;  // answer in [a, b]
        while (b - a > WP::L) {
          const int step = (b - a + WP::L) / (WP::L + 1);
          const int j = a + (w.lane() + 1) * step - 1;
          int cj = 1;
          if (j < b) cj = pathCompare(P(ORD(j)), pc);
          const unsigned lj = w.ballot(cj < 0);
          const unsigned ej = w.ballot(cj == 0);
          if (ej) { eqPos = a + (vffs(ej) + 1) * step - 1; pos = eqPos; return; }
          const int kk = vpopc(lj);  // probes 0..kk-1 are less (sorted => prefix)
          const int na = a + kk * step;
          const int nb = (kk < WP::L && a + (kk + 1) * step - 1 < b) ? a + (kk + 1) * step - 1 : b;
          a = na;
          b = nb;
        }
        lo = a;
        // finish with the linear chunk scan restricted to [a, b)
        while (lo < b) {
          const int i2 = lo + w.lane();
          int c2 = 1;
          if (i2 < b) c2 = pathCompare(P(ORD(i2)), pc);
          const unsigned l2 = w.ballot(c2 < 0);
          const unsigned e2 = w.ballot(c2 == 0);
          if (e2) { eqPos = lo + vffs(e2); pos = eqPos; return; }
          const int k2 = vpopc(l2);
          pos = lo + k2;
          if (k2 < WP::L) return;
          lo += WP::L;
        }
        pos = b;
        return;
      }
    }
  }

  // ---- chunked frontier (see `chunked`)
  VCE_HD void chunkInit(int firstSlot) {  // frontier = { firstSlot } (firstSlot < 0: empty)
    nbMax = (ordCap - 0) / (kBlkStride + 3);
    VCE_NOUNROLL
    for (int i = w.lane(); i < nbMax; i += WP::L) cfree()[i] = nbMax - 1 - i;  // pop order 0, 1, 2, ...
    w.sync();
    nbFree = nbMax - 1;
    nb = 1;
    dh = 0;
    if (w.lane() == 0) {
      cdir()[0] = 0;
      int32_t* B = cblk(0);
      B[0] = 0;
      B[1] = firstSlot >= 0 ? 1 : 0;
      B[2] = firstSlot;
    }
    w.sync();
  }
  VCE_HD int chunkPop() {  // pollFirst; n > 0
    int32_t* B = cblk(cdir()[dh]);
    const int st = B[0], cnt = B[1];
    const int head = B[2 + st];
    w.sync();
    if (cnt == 1 && nb > 1) {  // first block drained: release it
      if (w.lane() == 0) cfree()[nbFree] = cdir()[dh];
      ++nbFree;
      ++dh;
      --nb;
    } else if (w.lane() == 0) {
      B[0] = cnt == 1 ? 0 : st + 1;
      B[1] = cnt - 1;
    }
    w.sync();
    return head;
  }
  VCE_HD int chunkLastSlot(int bi) const {  // greatest path of directory entry bi
    const int32_t* B = cblk(cdir()[dh + bi]);
    return B[2 + B[0] + B[1] - 1];
  }
  // Block bi (directory index) and local position pos of path c: pos = #paths of the block that are less than c;
  // eq = true when the path at pos equals c.
  VCE_HD void chunkFind(int c, int& bi, int& pos, bool& eq) {
    const int32_t* pc = P(c);
    eq = false;
    // 1. the first block whose greatest path is >= c (the last block if there is none)
    int lo = 0, hi = nb - 1;  // answer in [lo, hi]
    while (lo < hi) {
      // probe up to WP::L block maxima spread over [lo, hi)
      const int span = hi - lo;
      const int step = (span + WP::L - 1) / WP::L;
      const int j = lo + w.lane() * step;   // probes lo, lo+step, ...
      int ge = 1;
      if (j < hi) ge = pathCompare(P(chunkLastSlot(j)), pc) >= 0 ? 1 : 0;
      const unsigned m = w.ballot(j < hi && ge);
      const unsigned valid = w.ballot(j < hi);
      if (m != 0) {
        const int k = vffs(m);             // first probed block with max >= c
        hi = lo + k * step;
        lo = k == 0 ? lo : lo + (k - 1) * step + 1;
        if (lo > hi) lo = hi;
      } else {
        lo = lo + (vpopc(valid) - 1) * step + 1;  // beyond the last probe
        if (lo > hi) lo = hi;
      }
    }
    bi = lo;
    // 2. position inside the block
    const int32_t* B = cblk(cdir()[dh + bi]);
    const int st = B[0], cnt = B[1];
    pos = cnt;
    VCE_NOUNROLL
    for (int base = 0; base < cnt; base += WP::L) {
      const int i = base + w.lane();
      int cmp = 1;
      if (i < cnt) cmp = pathCompare(P(B[2 + st + i]), pc);
      const unsigned less = w.ballot(cmp < 0);
      const unsigned e = w.ballot(cmp == 0);
      if (e) { pos = base + vffs(e); eq = true; return; }
      const int k = vpopc(less);
      if (k < WP::L) { pos = base + k; return; }
    }
  }
  VCE_HD void chunkShiftUp(int32_t* a, int from, int to) {  // a[from..to) -> a[from+1..to+1), highest chunk first
    VCE_NOUNROLL
    for (int base = ((to - from - 1) / WP::L) * WP::L; base >= 0 && to > from; base -= WP::L) {
      const int i = from + base + w.lane();
      int v = 0;
      const bool act = i < to;
      if (act) v = a[i];
      w.sync();
      if (act) a[i + 1] = v;
      w.sync();
    }
  }
  VCE_HD void chunkShiftDown(int32_t* a, int from, int to) {  // a[from..to) -> a[from-1..to-1)
    VCE_NOUNROLL
    for (int base = 0; from + base < to; base += WP::L) {
      const int i = from + base + w.lane();
      int v = 0;
      const bool act = i < to;
      if (act) v = a[i];
      w.sync();
      if (act) a[i - 1] = v;
      w.sync();
    }
  }
  VCE_HD bool chunkInsert(int bi, int pos, int c) {
    if (n >= cap) return false;
    int32_t* B = cblk(cdir()[dh + bi]);
    if (B[1] >= kBlk) {  // full: move the upper half to a new block right after this one
      if (nbFree <= 0) return false;
      const int nbk = cfree()[nbFree - 1];
      --nbFree;
      int32_t* N = cblk(nbk);
      const int st = B[0], half = kBlk / 2;
      VCE_NOUNROLL
      for (int i = w.lane(); i < half; i += WP::L) N[2 + i] = B[2 + st + half + i];
      w.sync();
      if (w.lane() == 0) { N[0] = 0; N[1] = half; B[1] = half; }
      // directory: open entry bi + 1
      if (dh + nb >= 2 * nbMax) {  // out of room at the top: compact to the bottom
        VCE_NOUNROLL
        for (int base = 0; base < nb; base += WP::L) {
          const int i = base + w.lane();
          int v = 0;
          const bool act = i < nb;
          if (act) v = cdir()[dh + i];
          w.sync();
          if (act) cdir()[i] = v;
          w.sync();
        }
        dh = 0;
      }
      chunkShiftUp(cdir() + dh, bi + 1, nb);
      if (w.lane() == 0) cdir()[dh + bi + 1] = nbk;
      ++nb;
      w.sync();
      if (pos > half) { ++bi; pos -= half; B = N; }
    }
    const int st = B[0], cnt = B[1];
    if (st > 0 && pos <= cnt - pos) {  // cheaper to move the front part down
      chunkShiftDown(B + 2, st, st + pos);
      if (w.lane() == 0) { B[2 + st - 1 + pos] = c; B[0] = st - 1; B[1] = cnt + 1; }
    } else {
      if (st + cnt >= kBlk) {  // no room at the top of the block: compact it first
        chunkShiftDownBy(B + 2, st, st + cnt, st);
        if (w.lane() == 0) B[0] = 0;
        w.sync();
        chunkShiftUp(B + 2, pos, cnt);
        if (w.lane() == 0) { B[2 + pos] = c; B[1] = cnt + 1; }
      } else {
        chunkShiftUp(B + 2, st + pos, st + cnt);
        if (w.lane() == 0) { B[2 + st + pos] = c; B[1] = cnt + 1; }
      }
    }
    ++n;
    if (n > maxFrontier) maxFrontier = n;
    w.sync();
    return true;
  }
  VCE_HD void chunkShiftDownBy(int32_t* a, int from, int to, int by) {  // a[from..to) -> a[from-by..to-by), by <= from
    VCE_NOUNROLL
    for (int base = 0; from + base < to; base += WP::L) {
      const int i = from + base + w.lane();
      int v = 0;
      const bool act = i < to;
      if (act) v = a[i];
      w.sync();
      if (act) a[i - by] = v;
      w.sync();
    }
  }
  VCE_HD bool addIfBetterChunked(int c) {
    int bi, pos;
    bool eq;
    w.sync();
    chunkFind(c, bi, pos, eq);
    if (eq) {
      int32_t* B = cblk(cdir()[dh + bi]);
      const int other = B[2 + B[0] + pos];
      const bool cWins = firstIsBetter(P(c), P(other));
      w.sync();
      if (cWins) {
        if (w.lane() == 0) B[2 + B[0] + pos] = c;
        pushFree(other);
      } else {
        pushFree(c);
      }
      w.sync();
      return true;
    }
    return chunkInsert(bi, pos, c);
  }

  // PathFinder.addIfBetter :321-348.  Returns false on frontier overflow.
  VCE_HD bool addIfBetter(int c) {
    if (chunked) return addIfBetterChunked(c);
    int pos, eqPos;
    w.sync();
    findPos(c, pos, eqPos);
    if (eqPos >= 0) {
      const int other = ORD(eqPos);
      const bool cWins = firstIsBetter(P(c), P(other));
      w.sync();
      if (cWins) {
        if (w.lane() == 0) ORD(eqPos) = c;
        pushFree(other);
      } else {
        pushFree(c);
      }
      w.sync();
      return true;
    }
    if (n >= cap) return false;
    if (oh > 0 && pos <= n - pos) {
      // shift [0, pos) down by one (cheap: new paths land near the front)
      VCE_NOUNROLL
      for (int base = 0; base < pos; base += WP::L) {
        const int i = base + w.lane();
        int v = 0;
        const bool act = i < pos;
        if (act) v = ORD(i);
        w.sync();
        if (act) order[oh + i - 1] = v;
        w.sync();
      }
      --oh;
    } else {
      if (oh + n >= ordCap) {  // out of room at the top: compact to the bottom
        VCE_NOUNROLL
        for (int base = 0; base < n; base += WP::L) {
          const int i = base + w.lane();
          int v = 0;
          const bool act = i < n;
          if (act) v = ORD(i);
          w.sync();
          if (act) order[i] = v;
          w.sync();
        }
        oh = 0;
      }
      // shift [pos, n) up by one, highest chunk first
      VCE_NOUNROLL
      for (int base = ((n - pos - 1) / WP::L) * WP::L; base >= 0; base -= WP::L) {
        const int i = pos + base + w.lane();
        int v = 0;
        const bool act = i < n;
        if (act) v = ORD(i);
        w.sync();
        if (act) ORD(i + 1) = v;
        w.sync();
      }
    }
    if (w.lane() == 0) ORD(pos) = c;
    ++n;
    if (n > maxFrontier) maxFrontier = n;
    w.sync();
    return true;
  }

  // ---------------------------------------------------------------- PathFinder pieces run by lane 0
  VCE_HD int sideCount(int side) const { return side == 0 ? R.cCount : R.bCount; }
  VCE_HD int sideFirst(int side) const { return side == 0 ? R.cFirst : R.bFirst; }
  VCE_HD int sideNextStart(int side) const { return side == 0 ? R.cNextStart : R.bNextStart; }

  // PathFinder.nextVariant :296-309.  Returns the local index, -1 for none, -2 when the variant that is due
  // lies beyond this region.
  VCE_HD int nextVariant(const int32_t* p, int side) const {
    const int nextIdx = p[L.sideBase(side) + SD_VARIDX] + 1;
    if (nextIdx >= sideCount(side)) {
      const int ns = sideNextStart(side);
      if (ns == kNoNext) return -1;
      return ns <= sidePos(p, side) + 1 ? -2 : -1;
    }
    const int st = S[side].vStart[sideFirst(side) + nextIdx];
    if (st <= sidePos(p, side) + 1) return nextIdx;
    if (sideWants(p, side) && st <= p[L.sideBase(side) + SD_VAREND]) return nextIdx;
    return -1;
  }
  VCE_HD int futureVariantPos(const int32_t* p, int side) const {  // :285-292
    const int nextIdx = p[L.sideBase(side) + SD_VARIDX] + 1;
    if (nextIdx >= sideCount(side)) {
      const int ns = sideNextStart(side);
      return ns == kNoNext ? R.tmplLen - 1 : ns;
    }
    return S[side].vStart[sideFirst(side) + nextIdx];
  }
  VCE_HD void pathMoveForward(int32_t* p, int position) {  // Path.moveForward :325-328
    VCE_NOUNROLL
    for (int side = 0; side < 2; ++side)
      VCE_NOUNROLL
      for (int h = 0; h < L.H; ++h) hapMoveForward(p + L.hap(side, h), side, position);
  }
  VCE_HD void skipToNextVariant(int32_t* p) {  // :271-282
    const int aNext = futureVariantPos(p, 0);
    const int bNext = futureVariantPos(p, 1);
    int m = aNext < bNext ? aNext : bNext;
    const int lastTemplatePos = R.tmplLen - 1;
    m = m < lastTemplatePos ? m : lastTemplatePos;
    const int nextPos = m - 1;
    if (nextPos > sidePos(p, 0)) pathMoveForward(p, nextPos);
  }
  VCE_HD void skipVariantsTo(int32_t* p, int side, int maxPos) {  // :255-264
    int varIndex = p[L.sideBase(side) + SD_VARIDX];
    const int cnt = sideCount(side);
    while (varIndex < cnt && (varIndex == -1 || S[side].vStart[sideFirst(side) + varIndex] < maxPos)) ++varIndex;
    --varIndex;
    p[L.sideBase(side) + SD_VARIDX] = varIndex;
    const int mv = maxPos < R.tmplLen - 1 ? maxPos : R.tmplLen - 1;
    VCE_NOUNROLL
    for (int h = 0; h < L.H; ++h) hapMoveForward(p + L.hap(side, h), side, mv);
  }
  VCE_HD void pathStep(int32_t* p) {  // Path.step :330-341
    const int trailing = sideTrailing(p, 0);
    if (trailing == -1) {
      VCE_NOUNROLL
      for (int side = 0; side < 2; ++side)
        VCE_NOUNROLL
        for (int h = 0; h < L.H; ++h) hapStep(p + L.hap(side, h), side);
    } else {
      hapStep(p + L.hap(0, trailing), 0);
      hapStep(p + L.hap(1, trailing), 1);
    }
  }
  VCE_HD bool sideIsNew(const int32_t* p, int side, int vGlobal, int row) const {  // HalfPath.isNew :140-151
    if (S[side].vStart[vGlobal] >= p[L.sideBase(side) + SD_INCLEND]) return true;
    VCE_NOUNROLL
    for (int h = 0; h < L.H; ++h) {
      if (!hapIsNew(p + L.hap(side, h), side, S[side].oSlot[(size_t) row * L.H + h])) return false;
    }
    return true;
  }

  VCE_HD void initPath(int32_t* p) {
    VCE_NOUNROLL
    for (int i = 0; i < L.W; ++i) p[i] = 0;
    VCE_NOUNROLL
    for (int side = 0; side < 2; ++side) {
      VCE_NOUNROLL
      for (int h = 0; h < L.H; ++h) {
        int32_t* hp = p + L.hap(side, h);
        hp[HP_TPOS] = R.startPos;
        hp[HP_PIA] = -1;
        hp[HP_LASTEND] = -1;
        hp[HP_NEXT] = -1;
      }
      p[L.sideBase(side) + SD_VARIDX] = -1;
    }
  }

  // ---------------------------------------------------------------- main loop: PathFinder.bestPath :129-214
  // Returns a kRegion* status.  On kRegionClean / kRegionFinished `resultSlot` holds the surviving path.
  VCE_HD int run(int& resultSlot) {
    n = 0; iters = 0; curMaxPos = 0; lastSyncPos = 0; bestValid = 0; usedPrefix = 0; nWarn = 0; maxFrontier = 1;
    totalIters = 0;
    nFree = 0;
    VCE_NOUNROLL
    for (int i = w.lane(); i < nSlots; i += WP::L) freeList[i] = nSlots - 1 - i;  // pop order 0,1,2,...
    nFree = nSlots;
    if (w.lane() == 0) {
      VCE_NOUNROLL
      for (int i = 0; i < 16; ++i) ctl[i] = 0;
      initPath(P(0));
      if (!chunked) order[0] = 0;
    }
    oh = 0;
    nFree -= 1;  // slot 0 taken
    n = 1;
    w.sync();
    if (chunked) chunkInit(0);
    resultSlot = -1;

    while (n > 0) {
      ++iters;  // :145 currentIterations++
      ++totalIters;
      int head;
      if (chunked) {
        head = chunkPop();
      } else {
        head = ORD(0);
        ++oh;  // pollFirst
      }
      --n;
      bool headIsLS = false;
      if (n == 0) {  // :151-161 only one path in play
        lastSyncPos = sidePos(P(head), 0);
        iters = 0;
        headIsLS = true;
      } else if (n > cfg.maxPaths || iters > cfg.maxIterations) {  // :162-169 too complex
        if (w.lane() == 0 && warnOut != nullptr && nWarn < warnCap) {
          warnOut[nWarn] = WarnRec{regionId, n, iters, lastSyncPos + 1, curMaxPos + 2};
        }
        ++nWarn;
        w.sync();
        VCE_NOUNROLL
        for (int i = w.lane(); i < nSlots; i += WP::L) freeList[i] = nSlots - 1 - i;
        nFree = nSlots;
        n = 0;
        oh = 0;
        iters = 0;
        head = 0;
        nFree -= 1;
        w.sync();
        if (chunked) chunkInit(-1);
        copyPath(head, slotLS());
        w.sync();
        if (w.lane() == 0) {
          skipVariantsTo(P(head), 0, curMaxPos + 1);
          skipVariantsTo(P(head), 1, curMaxPos + 1);
        }
        w.sync();
        if (ctl[CTL_FLAGS] & FL_ERRMOVE) return kRegionErrMove;
        headIsLS = true;
      }

      // ---- lane 0 decides what happens to the head
      if (w.lane() == 0) {
        int32_t* hp = P(head);
        int act = 0, side = 0, vidx = -1;
        const int cv = nextVariant(hp, 0);   // :170 calls first
        if (cv >= 0) { act = ACT_ENQUEUE; side = 0; vidx = cv; }
        else if (cv == -2) { act = ACT_BOUNDARY_SPILL; }
        else {
          const int bv = nextVariant(hp, 1);  // :173 then baseline
          if (bv >= 0) { act = ACT_ENQUEUE; side = 1; vidx = bv; }
          else if (bv == -2) { act = ACT_BOUNDARY_SPILL; }
        }
        if (act == ACT_BOUNDARY_SPILL) {
          // The next variant belongs to the following region.  The cut is exact only if this is the single
          // surviving path, in sync, standing right before that variant, with nothing of this region left.
          const int cn = R.cNextStart, bn = R.bNextStart;
          const int boundary = (cn < bn ? cn : bn) - 1;
          if (n == 0 && inSync(hp) && sidePos(hp, 0) == boundary
              && hp[L.sideBase(0) + SD_VARIDX] + 1 >= R.cCount && hp[L.sideBase(1) + SD_VARIDX] + 1 >= R.bCount) {
            act = ACT_BOUNDARY_CLEAN;
          }
        } else if (act == ACT_ENQUEUE) {
          const int vg = sideFirst(side) + vidx;
          const int st = S[side].vStart[vg];
          ctl[CTL_AUX] = st;
          ctl[CTL_AUX + 1] = 0;
          // Path.addVariant :288-300 sync point handling on the parent
          if (inSync(hp)) {
            int32_t* pb = hp + L.pathBase();
            const int syncPos = sidePos(hp, 0);
            const int ns = pb[PT_NSYNC];
            if (!(ns > 0 && pb[PT_SYNC0 + ns - 1] == syncPos)) {
              if (ns >= L.SMAX) flag(FL_OVERFLOW);
              else { pb[PT_SYNC0 + ns] = syncPos; pb[PT_NSYNC] = ns + 1; ctl[CTL_AUX + 1] = 1; }
            }
            pb[PT_CSINCE] = 0;
            pb[PT_BSINCE] = 0;
          }
        } else {
          pathStep(hp);  // :177
          act = ACT_STEP_DROP;
          bool alive = true;
          if (inSync(hp)) {  // :182-190
            if (cfg.pruneNoOps && hasNoOp(hp)) alive = false;
            else skipToNextVariant(hp);
          }
          if (alive && pathMatches(hp)) {  // :192
            if (sideFinished(hp, 0) && sideFinished(hp, 1)) {  // :193
              if (!(cfg.pruneNoOps && hasNoOp(hp))) {
                int32_t* pb = hp + L.pathBase();
                const int ns = pb[PT_NSYNC];
                if (ns >= L.SMAX) flag(FL_OVERFLOW);
                else { pb[PT_SYNC0 + ns] = sidePos(hp, 0); pb[PT_NSYNC] = ns + 1; }  // :199
                act = ACT_STEP_FINISH;
              }
            } else {
              act = ACT_STEP_REINSERT;
            }
          }
        }
        ctl[CTL_ACT] = act;
        ctl[CTL_SIDE] = side;
        ctl[CTL_VIDX] = vidx;
      }
      w.sync();
      const int act = ctl[CTL_ACT];
      const int fl = ctl[CTL_FLAGS];
      if (fl & FL_ERRORDER) return kRegionErrOrder;
      if (fl & FL_ERRMOVE) return kRegionErrMove;
      if (fl & FL_OVERFLOW) return kRegionOverflow;
      if (fl & FL_WINMISS) return kRegionWinMiss;

      if (act == ACT_BOUNDARY_CLEAN) { resultSlot = head; return kRegionClean; }
      if (act == ACT_BOUNDARY_SPILL) return kRegionSpill;

      if (act == ACT_ENQUEUE) {
        const int side = ctl[CTL_SIDE];
        const int vidx = ctl[CTL_VIDX];
        const int vg = sideFirst(side) + vidx;
        const int st = ctl[CTL_AUX];
        curMaxPos = curMaxPos > st ? curMaxPos : st;  // :236
        const int row0 = S[side].oOff[vg];
        const int nOr = S[side].oOff[vg + 1] - row0;
        const int nChild = 1 + nOr;
        if (nChild > maxChildren || nChild > nFree) return kRegionOverflow;
        // children: cooperative copy of the parent, then lane k specialises child k
        VCE_NOUNROLL
        for (int k = 0; k < nChild; ++k) copyPath(freeList[nFree - 1 - k], head);
        w.sync();
        VCE_NOUNROLL
        for (int k = w.lane(); k < nChild; k += WP::L) {
          const int cs = freeList[nFree - 1 - k];
          int32_t* cp = P(cs);
          int32_t* sb = cp + L.sideBase(side);
          const int vend = S[side].vEnd[vg];
          int valid = 1;
          if (k == 0) {  // HalfPath.exclude :153-158
            sb[SD_VAREND] = sb[SD_VAREND] > vend ? sb[SD_VAREND] : vend;
            sb[SD_VARIDX] = vidx;
            setDecision(cp, side, vidx, 1);
          } else {
            const int row = row0 + k - 1;
            if (sideIsNew(P(head), side, vg, row)) {  // HalfPath.include :160-170, Path.include :267-275
              sb[SD_VAREND] = sb[SD_VAREND] > vend ? sb[SD_VAREND] : vend;
              sb[SD_VARIDX] = vidx;
              sb[SD_INCLEND] = sb[SD_INCLEND] > vend ? sb[SD_INCLEND] : vend;
              sb[SD_INCLCNT] += 1;
              sb[SD_LASTID] = S[side].oIds[(size_t) row * 4];
              setDecision(cp, side, vidx, 1 + k);
              VCE_NOUNROLL
              for (int h = 0; h < L.H; ++h) hapAdd(cp + L.hap(side, h), side, S[side].oSlot[(size_t) row * L.H + h]);
              cp[L.pathBase() + (side == 0 ? PT_CSINCE : PT_BSINCE)] += 1;
            } else {
              valid = 0;
            }
          }
          childValid(k) = valid;
          childSlot(k) = cs;
        }
        w.sync();
        if (ctl[CTL_FLAGS] & FL_OVERFLOW) return kRegionOverflow;
        if (headIsLS) {
          // lastSyncPath is the PARENT object: it got the since-sync reset (Path.java:299-300) but the new sync
          // point belongs to the children only (the parent's mSyncPointList is final, Path.java:296,305)
          copyPath(slotLS(), head);
          w.sync();
          if (w.lane() == 0 && ctl[CTL_AUX + 1]) P(slotLS())[L.pathBase() + PT_NSYNC] -= 1;
          w.sync();
        }
        // claim the child slots, release the parent
        nFree -= nChild;
        pushFree(head);
        w.sync();
        VCE_NOUNROLL
        for (int k = 0; k < nChild; ++k) {
          const int cs = childSlot(k);
          if (childValid(k)) {
            if (!addIfBetter(cs)) return kRegionOverflow;
          } else {
            pushFree(cs);
            w.sync();
          }
        }
        continue;
      }

      if (act == ACT_STEP_DROP) {
        pushFree(head);
        w.sync();
        continue;
      }
      if (act == ACT_STEP_FINISH) {  // :199-201
        bool take = true;
        if (bestValid) take = !firstIsBetter(P(slotBest()), P(head));
        w.sync();
        if (take) { copyPath(slotBest(), head); bestValid = 1; }
        pushFree(head);
        w.sync();
        continue;
      }
      // ACT_STEP_REINSERT :203
      if (!addIfBetter(head)) return kRegionOverflow;
    }
    if (R.cNextStart != kNoNext || R.bNextStart != kNoNext) {
      // every path died before reaching the boundary although more variants follow on this sequence:
      // the reference keeps going with whatever finished (nothing can have), so hand it to the merged run.
      return kRegionSpill;
    }
    if (!bestValid) return kRegionErrNull;
    resultSlot = slotBest();
    return kRegionFinished;
  }

  // ---------------------------------------------------------------- epilogue: decisions, sync points, weights
  // Path.calculateWeights E/Path.java:385-453 restricted to this region: a variant belongs to the first sync point
  // at or after its start; variants after the region's last sync point belong to the closing sync point that the
  // next region (or the finish) contributes, which holds only this region's trailing variants.
  VCE_HD void writeResults(int slot, int32_t* syncOut, RegionResult* rr) {
    const int32_t* p = P(slot);
    const int ns = p[L.pathBase() + PT_NSYNC];
    const int32_t* sy = p + L.pathBase() + PT_SYNC0;
    VCE_NOUNROLL
    for (int i = w.lane(); i < ns; i += WP::L) syncOut[i] = sy[i];
    VCE_NOUNROLL
    for (int side = 0; side < 2; ++side) {
      const int cnt = sideCount(side), first = sideFirst(side);
      VCE_NOUNROLL
      for (int i = w.lane(); i < cnt; i += WP::L) S[side].outDecision[first + i] = (uint8_t) getDecision(p, side, i);
    }
    // weights for included calls
    VCE_NOUNROLL
    for (int i = w.lane(); i < R.cCount; i += WP::L) {
      double wgt = 0.0;
      if (getDecision(p, 0, i) >= 2) {
        const int st = S[0].vStart[R.cFirst + i];
        int k = 0;
        while (k < ns && sy[k] < st) ++k;   // interval id; ns = closing interval
        const int lo = k == 0 ? -0x7fffffff : sy[k - 1];  // interval = starts in (lo, hi]
        const int hi = k < ns ? sy[k] : 0x7fffffff;
        int calledTP = 0, baseTP = 0, baseInside = 0;
        VCE_NOUNROLL
        for (int j = 0; j < R.cCount; ++j) {
          if (getDecision(p, 0, j) >= 2) {
            const int sj = S[0].vStart[R.cFirst + j];
            if (sj > lo && sj <= hi) ++calledTP;
          }
        }
        VCE_NOUNROLL
        for (int j = 0; j < R.bCount; ++j) {
          if (getDecision(p, 1, j) >= 2) {
            const int sj = S[1].vStart[R.bFirst + j];
            if (sj > lo && sj <= hi) {
              ++baseTP;
              if (!(S[1].vFlags[R.bFirst + j] & 0x10)) ++baseInside;
            }
          }
        }
        wgt = baseTP == 0 ? 0.0 : (double) baseInside / (double) calledTP;
      }
      S[0].outWeight[R.cFirst + i] = wgt;
    }
    if (w.lane() == 0) {
      rr->nSync = ns;
      rr->lastBaseAlleleId = p[L.sideBase(1) + SD_INCLCNT] > 0 ? p[L.sideBase(1) + SD_LASTID] : kPrefixEmpty;
    }
  }
};

}  // namespace vce
#endif
