// vce_cta.h -- the bounded best-first haplotype search for ONE large region, run by ONE thread block (CTA).
//
// Same behavioural contract as vce_search.h (bit-exact PathFinder.bestPath, E/PathFinder.java:129-348, including the
// too-complex skip semantics :162-169 on the identical iteration); the path semantics themselves (playback, half
// paths, sync points, tie-breaks) are the RegionSearch members of vce_search.h, run on path records that live in
// SHARED memory.  What is new here is everything around them -- the part that made dense clusters (BASELINE configs[2],
// frontiers of tens of thousands of paths) slow in the warp-per-region kernels:
//
//   * FIXED-SIZE SORT KEYS.  Path.compareTo (E/Path.java:351-357) walks four HaplotypePlayback states and, for every
//     pending allele, Allele.compareTo (start, end, length, reversed bases -- E/Allele.java:95-119).  Per region all
//     alleles of a side are ranked once under that order; a path's position in the frontier is then the lexicographic
//     order of a 64-byte key  (templatePos, rank(next), posInAllele, ranks of the queue)  x 4 haplotypes: a
//     comparison is a few word compares, no table chase.
//   * THE FRONTIER (java.util.TreeSet) is a two-level sorted structure: a directory of blocks of up to 128 entries
//     (key + tie-break fields + record slot).  The directory with every block's largest key sits in shared memory;
//     the blocks live in a per-CTA arena in HBM and are worked on in a small write-back CACHE OF BLOCKS IN SHARED
//     MEMORY (the front block, which serves every pollFirst and most inserts, stays resident).  Blocks move between
//     HBM and shared memory with 1-D bulk TMA copies (cp.async.bulk + mbarrier) on the device.
//   * PATH RECORDS of frontier members live in an HBM pool; the head being expanded and its children are built in
//     shared-memory work slots.  A path popped right after it was inserted is still in its work slot (tag match), and
//     the record of the next front entry is prefetched into registers while the current head is processed.
//   * Equal paths are merged by PathFinder.addIfBetter (:321-348) from the tie-break fields stored beside the key, so a
//     duplicate never costs a record read.
//
// Everything is written against a warp policy (vce_search.h) so that the same source runs on the device and, for the
// host-logic tests only, on the CPU with one lane.  Anything outside the fixed shapes (ploidy > 2, more than four
// pending alleles per haplotype, alleles of 65 535 bases and more, too many sync points, arena exhausted) reports
// kRegionOverflow and is re-run by the large-capacity warp kernel, so the shapes never change a result.
#ifndef VCE_CTA_H_
#define VCE_CTA_H_

#include "vce_device.h"
#include "vce_search.h"

#if defined(VCE_CTA_DEBUG) && defined(__CUDA_ARCH__)
#include <cstdio>
#define VCE_CTA_CHECK(cond, code, a, b)                                                                          \
  do {                                                                                                           \
    if (!(cond)) {                                                                                               \
      printf("[k_cta] check %d failed: block %d lane %d values %d %d\n", (int) (code), (int) blockIdx.x, (int) (threadIdx.x & 31), (int) (a), (int) (b)); \
      __trap();                                                                                                  \
    }                                                                                                            \
  } while (0)
#else
#define VCE_CTA_CHECK(cond, code, a, b) do { } while (0)
#endif

namespace vce {

struct CtaShape {
  int nb;         // frontier blocks (directory entries in shared memory)
  int ncache;     // blocks cached in shared memory (2 or 3)
  int tabBytes;   // shared memory for the region's tables
  int winBytes;   // shared memory for the reference window
  VCE_HD int capacity() const { return nb * 64 - 64; }   // paths that always fit (split blocks stay at least half full)
};

namespace cta {
constexpr int kBlk = 128;            // entries per block
constexpr int kKeyW = 16;            // key words (4 per haplotype, calls h0 h1, baseline h0 h1)
constexpr int kAuxW = 4;             // tie-break words
constexpr int kQ = 8;                // pending alleles per haplotype a key holds with 8-bit ranks (4 with 16-bit ranks)
constexpr int kSmaxCap = 44;         // sync points per record
constexpr int kMaxW = 128;           // record words (one 16-byte chunk per lane)
constexpr int kMaxWork = 8;          // work slots in shared memory that hold the head and its children (+ last-sync, best)
// block layout (32-bit words)
constexpr int B_START = 0, B_CNT = 1, B_FREE = 2 /* 4 words */, B_ORD = 8 /* 256 bytes */, B_SLOT = 72 /* kBlk */,
              B_AUX = 200 /* kBlk * kAuxW */, B_KEY = 712 /* kBlk * kKeyW */, B_WORDS = 2760;
static_assert(B_SLOT == B_ORD + 64 && B_AUX == B_SLOT + kBlk && B_KEY == B_AUX + kBlk * kAuxW && B_WORDS == B_KEY + kBlk * kKeyW, "block layout");
static_assert((B_WORDS * 4) % 16 == 0, "blocks move as 16-byte chunks");
constexpr int kFreeCache = 64;       // free record slots kept in shared memory
}  // namespace cta

// Shared-memory words a CTA needs for `shape` (everything except the RegionConst that precedes it).
VCE_HD int ctaSmemWords(const CtaShape& s) {
  using namespace cta;
  int wds = 0;
  wds += s.nb * kKeyW;                 // largest key of every block
  wds += (2 * s.nb + s.nb + 1) / 2 + 2;  // directory (deque of uint16) + free block stack (uint16)
  wds = (wds + 3) & ~3;                // (carve() starts the cached blocks on a 16-byte boundary)
  wds += s.ncache * B_WORDS;           // cached blocks
  wds += (kMaxWork + 2) * kMaxW;       // work slots
  wds += kKeyW + kAuxW + 12;           // key of the path being inserted
  wds += kFreeCache;                   // free record slots
  wds += 16 + 2 * cta::kMaxWork + 4;   // ctl mailbox of the RegionSearch
  wds = (wds + 3) & ~3;
  wds += (s.tabBytes + s.winBytes) / 4;
  return wds;
}

// Bytes of HBM arena a CTA needs for `shape` with records of W words and `poolCap` record slots.
VCE_HD long long ctaArenaWords(const CtaShape& s, int W, long long poolCap) {
  return (long long) s.nb * cta::B_WORDS + poolCap * W + poolCap + 64;
}

// Host policy (tests only): one lane, blocks move with memcpy.
struct HostLaneMove : HostLane {
  void loadBlock(uint32_t* dst, const uint32_t* src, int bytes) const { for (int i = 0; i < bytes / 4; ++i) dst[i] = src[i]; }
  void storeBlock(uint32_t* dst, const uint32_t* src, int bytes) const { for (int i = 0; i < bytes / 4; ++i) dst[i] = src[i]; }
  void storeWait() const {}
};

// Record layout of a region in the CTA kernel: queue capacity of the keys, as many sync points as a 128-word record
// allows.  Returns false when no layout fits (the region then runs in the large-capacity warp kernel).
// Ranks fit 8 bits (and a key holds eight pending alleles per haplotype) when neither side has more than 254 allele slots.
VCE_HD bool ctaNarrowRanks(int cSlotN, int bSlotN) { return cSlotN <= 254 && bSlotN <= 254; }
VCE_HD bool ctaLayout(PathLayout& L, int H, int cCount, int bCount, int decBits, bool narrowRanks) {
  const int kv = cCount > bCount ? cCount : bCount;
  const int q = narrowRanks ? cta::kQ : cta::kQ / 2;
  int smax = cCount + bCount + 2;
  if (smax > cta::kSmaxCap) smax = cta::kSmaxCap;
  L.init(H, q, kv < 1 ? 1 : kv, smax, decBits);
  while (L.W > cta::kMaxW && smax > 8) {
    smax -= 4;
    L.init(H, q, kv < 1 ? 1 : kv, smax, decBits);
  }
  return L.W <= cta::kMaxW;
}

template <class WP>
struct CtaSearch {
  typedef RegionSearch<WP, false> RS;
  RS rs;
  WP w;
  CtaShape shape;
  // shared memory
  uint32_t* sep;        // [nb][kKeyW]
  uint16_t* dir;        // [2 * nb] deque of block ids
  uint16_t* blkFree;    // [nb]
  uint32_t* cacheMem;   // [ncache][B_WORDS]
  uint32_t* keyNew;     // [kKeyW + kAuxW]
  int32_t* freeCache;   // [kFreeCache]
  uint16_t* rank[2];    // [slots of the side] rank under Allele.compareTo
  // HBM arena
  uint32_t* gBlocks;    // [nb][B_WORDS]
  int32_t* gPool;       // [poolCap][W]
  int32_t* gFree;       // [poolCap] stack of recycled record slots
  int poolCap;
  // warp-uniform state
  int dh, nbUsed, nbFree;              // directory head offset, blocks in use, free block ids
  int cid[3], cdirty[3], cuse[3];      // cache: block id (-1 = empty), dirty, last use
  int useClock;
  int nFreeCache, nFreeStack, bump;    // record slot allocator
  int tag[cta::kMaxWork];              // record slot whose content a work slot still holds (-1 = none)
  int nWork;                           // work slots that can carry a tag (head + children)
  int initialPending;
  // prefetched record of the next front entry: one 16-byte chunk per lane
  int pfSlot;
  Quad pf[WP::L == 1 ? cta::kMaxW / 4 : 1];
  bool unsupported;
  bool narrow;          // 8-bit ranks (ctaNarrowRanks)

  VCE_HD explicit CtaSearch(RegionConst* k) : rs(k) {}

  // ---------------------------------------------------------------------------------------------- keys
  VCE_HD static int keyCmp(const uint32_t* a, const uint32_t* b) {
    VCE_NOUNROLL
    for (int i = 0; i < cta::kKeyW; ++i) {
      const uint32_t x = a[i], y = b[i];
      if (x != y) return x < y ? -1 : 1;
    }
    return 0;
  }
  VCE_HD int rankOf(int side, int slot) const {
    VCE_CTA_CHECK(slot - (side == 0 ? rs.R.cSlot0 : rs.R.bSlot0) >= 0 && slot - (side == 0 ? rs.R.cSlot0 : rs.R.bSlot0) < (side == 0 ? rs.R.cSlotN : rs.R.bSlotN), 1, slot, side);
    return rank[side][slot - (side == 0 ? rs.R.cSlot0 : rs.R.bSlot0)];
  }
  // Key and tie-break fields of the path record p (cooperative; call sync() before reading them).
  VCE_HD void makeKey(const int32_t* p, uint32_t* key) {
    const PathLayout& L = rs.L;
    VCE_NOUNROLL
    for (int q = w.lane(); q < 4; q += WP::L) {
      const int side = q >> 1, h = q & 1;
      uint32_t k0 = 0, k1 = 0, k2 = 0, k3 = 0;
      if (h < L.H) {
        const int32_t* hp = p + L.hap(side, h);
        k0 = (uint32_t) (hp[HP_TPOS] + 2);
        const int nx = hp[HP_NEXT];
        if (nx >= 0) {
          k1 = ((uint32_t) (rankOf(side, nx) + 1) << 16) | (uint32_t) (hp[HP_PIA] + 1);
          const int ql = hp[HP_QINFO] & 0xffff;
          // the pending alleles, most significant first, absent = 0: eight 8-bit ranks or four 16-bit ranks
          const int bits = narrow ? 8 : 16;
          VCE_NOUNROLL
          for (int i = 0; i < ql; ++i) {
            const uint32_t rk = (uint32_t) (rankOf(side, hp[HP_Q0 + i]) + 1);
            const int bit = i * bits;
            if (bit < 32) k2 |= rk << (32 - bits - bit);
            else k3 |= rk << (64 - bits - bit);
          }
        }
      }
      key[q * 4 + 0] = k0; key[q * 4 + 1] = k1; key[q * 4 + 2] = k2; key[q * 4 + 3] = k3;
    }
    if (w.lane() == 0) {
      uint32_t* aux = key + cta::kKeyW;
      aux[0] = (uint32_t) p[L.sideBase(0) + SD_INCLCNT] | ((uint32_t) p[L.sideBase(1) + SD_INCLCNT] << 16);
      aux[1] = ((uint32_t) p[L.pathBase() + PT_CSINCE] & 0xffffu) | ((uint32_t) p[L.pathBase() + PT_BSINCE] << 16);
      aux[2] = (uint32_t) rs.lastSync(p);
      aux[3] = (uint32_t) p[L.sideBase(1) + SD_LASTID];
    }
  }
  // PathPreference.better(first, second) on the stored tie-break fields: RegionSearch::firstIsBetter, same order of tests.
  VCE_HD bool betterAux(const uint32_t* f, const uint32_t* s) {
    const int fc = (int) (f[0] & 0xffffu), fb = (int) (f[0] >> 16), sc = (int) (s[0] & 0xffffu), sb = (int) (s[0] >> 16);
    const bool prefixNonEmpty = rs.R.prefixAlleleId != kPrefixEmpty;
    if (rs.cfg.preference == 0) {  // MaxSumBoth.java:43-82
      const int fs = fc + fb, ss = sc + sb;
      if (fs == ss) {
        if (fb == 0 || sb == 0) rs.usedPrefix |= 1;
        const bool fiNonNull = prefixNonEmpty || fb > 0;
        const bool siNonNull = prefixNonEmpty || sb > 0;
        if (fiNonNull && siNonNull) {
          int fd = (int) (f[1] >> 16) - (int) (f[1] & 0xffffu);
          int sd = (int) (s[1] >> 16) - (int) (s[1] & 0xffffu);
          fd = fd < 0 ? -fd : fd;
          sd = sd < 0 ? -sd : sd;
          if (fd != sd) return fd < sd;
          const int syncDelta = (int) f[2] - (int) s[2];
          if (syncDelta != 0) return syncDelta > 0;
          if ((fb == 0) != (sb == 0)) rs.usedPrefix |= 2;
          const int fid = fb > 0 ? (int) f[3] : rs.R.prefixAlleleId;
          const int sid = sb > 0 ? (int) s[3] : rs.R.prefixAlleleId;
          return fid < sid;
        }
      }
      return fs > ss;
    }
    if (fc == sc) {  // MaxCallsMinBaseline.java:42-61
      if (fb == sb) {
        if (fb > 0) return (int) f[3] < (int) s[3];
      }
      return fb < sb;
    }
    return fc > sc;
  }

  // Ranks of the region's alleles under Allele.compareTo, per side: rank = number of alleles that compare less (equal
  // alleles share a rank).  Also checks what the keys cannot hold.
  VCE_HD void computeRanks() {
    unsupported = false;
    narrow = ctaNarrowRanks(rs.R.cSlotN, rs.R.bSlotN);
    VCE_NOUNROLL
    for (int side = 0; side < 2; ++side) {
      const int s0 = side == 0 ? rs.R.cSlot0 : rs.R.bSlot0, sn = side == 0 ? rs.R.cSlotN : rs.R.bSlotN;
      if (sn >= 65000) unsupported = true;
      VCE_NOUNROLL
      for (int a = w.lane(); a < sn; a += WP::L) {
        int r = 0;
        if (rs.alLen(side, s0 + a) >= 65000) r = -1;
        VCE_NOUNROLL
        for (int b = 0; b < sn && r >= 0; ++b) {
          if (b != a && RS::alleleCompareS(rs.S + side, s0 + b, s0 + a) < 0) ++r;
        }
        rank[side][a] = (uint16_t) (r < 0 ? 0xffff : r);
      }
    }
    w.sync();
    VCE_NOUNROLL
    for (int side = 0; side < 2; ++side) {
      const int sn = side == 0 ? rs.R.cSlotN : rs.R.bSlotN;
      bool bad = false;
      VCE_NOUNROLL
      for (int a = w.lane(); a < sn; a += WP::L) if (rank[side][a] == 0xffff) bad = true;
      if (w.ballot(bad) != 0) unsupported = true;
    }
  }

  // ---------------------------------------------------------------------------------------------- block cache
  VCE_HD uint32_t* cacheBlock(int c) const { return cacheMem + (size_t) c * cta::B_WORDS; }
  VCE_HD uint32_t* globalBlock(int id) const {
    VCE_CTA_CHECK(id >= 0 && id < shape.nb, 2, id, shape.nb);
    return gBlocks + (size_t) id * cta::B_WORDS;
  }
  // Moves are plain cooperative 16-byte copies here; the device backend replaces them with bulk TMA copies (blockMove).
  VCE_HD void copyWords16(uint32_t* dst, const uint32_t* src, int words) {
    Quad* d = reinterpret_cast<Quad*>(dst);
    const Quad* s = reinterpret_cast<const Quad*>(src);
    VCE_NOUNROLL
    for (int i = w.lane(); i < (words >> 2); i += WP::L) d[i] = s[i];
  }
  VCE_HD void writeBack(int c) {
    if (cid[c] >= 0 && cdirty[c]) {
      w.sync();
      w.storeBlock(globalBlock(cid[c]), cacheBlock(c), cta::B_WORDS * 4);
    }
    cdirty[c] = 0;
  }
  // Cache slot holding block `id` (loaded if necessary); `keep` = a slot that must not be evicted (-1 = none).
  long long statHit = 0, statMiss = 0, statInsFront = 0, statInsFar = 0, statMaxQ = 0;   // diagnostics (emulation / debug builds)
  VCE_HD int ensureCached(int id, int keep, bool load) {
    ++useClock;
    for (int c = 0; c < shape.ncache; ++c) if (cid[c] == id) { cuse[c] = useClock; ++statHit; return c; }
    if (load) ++statMiss;
    int victim = -1;
    for (int c = 0; c < shape.ncache; ++c) {
      if (c == keep) continue;
      if (cid[c] < 0) { victim = c; break; }
      if (victim < 0 || cuse[c] < cuse[victim]) victim = c;
    }
    VCE_CTA_CHECK(victim >= 0 && victim < shape.ncache, 3, victim, keep);
    writeBack(victim);
    w.storeWait();   // the slot's old content has been read out
    cid[victim] = id;
    cuse[victim] = useClock;
    cdirty[victim] = 0;
    if (load) {
      w.loadBlock(cacheBlock(victim), globalBlock(id), cta::B_WORDS * 4);
      w.sync();
    }
    return victim;
  }
  VCE_HD void dropCached(int id) {
    for (int c = 0; c < shape.ncache; ++c) if (cid[c] == id) { cid[c] = -1; cdirty[c] = 0; }
  }

  // ---------------------------------------------------------------------------------------------- record slots
  VCE_HD int allocRecord() {  // -1 when the pool is exhausted
    if (nFreeCache == 0) {
      if (nFreeStack > 0) {
        const int take = nFreeStack < 32 ? nFreeStack : 32;
        VCE_NOUNROLL
        for (int i = w.lane(); i < take; i += WP::L) freeCache[i] = gFree[nFreeStack - take + i];
        nFreeStack -= take;
        nFreeCache = take;
        w.sync();
      } else {
        if (bump >= poolCap) return -1;
        return bump++;
      }
    }
    return freeCache[--nFreeCache];
  }
  VCE_HD void invalidate(int slot) {
    for (int k = 0; k < nWork; ++k) if (tag[k] == slot) tag[k] = -1;
    if (pfSlot == slot) pfSlot = -1;
  }
  VCE_HD void freeRecord(int slot) {
    invalidate(slot);
    if (nFreeCache == cta::kFreeCache) {  // spill the older half
      w.sync();
      VCE_NOUNROLL
      for (int i = w.lane(); i < 32; i += WP::L) gFree[nFreeStack + i] = freeCache[i];
      w.sync();
      VCE_NOUNROLL
      for (int i = w.lane(); i < cta::kFreeCache - 32; i += WP::L) {
        const int v = freeCache[32 + i];
        freeCache[i] = v;   // (WP::L lanes move disjoint elements downwards in steps of L: read-before-write holds per pass)
      }
      nFreeStack += 32;
      nFreeCache -= 32;
      w.sync();
    }
    VCE_CTA_CHECK(nFreeCache >= 0 && nFreeCache < cta::kFreeCache && nFreeStack >= 0 && nFreeStack <= poolCap, 7, nFreeCache, nFreeStack);
    if (w.lane() == 0) freeCache[nFreeCache] = slot;
    ++nFreeCache;
    w.sync();
  }
  VCE_HD void storeRecord(int slot, const int32_t* p) {  // lane j writes chunk j (and later reads the same chunk)
    VCE_CTA_CHECK(slot >= 0 && slot < poolCap, 4, slot, poolCap);
    Quad* d = reinterpret_cast<Quad*>(gPool + (size_t) slot * rs.L.W);
    const Quad* s = reinterpret_cast<const Quad*>(p);
    VCE_NOUNROLL
    for (int i = w.lane(); i < (rs.L.W >> 2); i += WP::L) d[i] = s[i];
  }
  VCE_HD void loadRecord(int32_t* p, int slot) {
    VCE_CTA_CHECK(slot >= 0 && slot < poolCap, 5, slot, poolCap);
    const Quad* s = reinterpret_cast<const Quad*>(gPool + (size_t) slot * rs.L.W);
    Quad* d = reinterpret_cast<Quad*>(p);
    VCE_NOUNROLL
    for (int i = w.lane(); i < (rs.L.W >> 2); i += WP::L) d[i] = s[i];
  }
  VCE_HD void prefetchRecord(int slot) {
    VCE_CTA_CHECK(slot >= 0 && slot < poolCap, 6, slot, poolCap);
    pfSlot = slot;
    const Quad* s = reinterpret_cast<const Quad*>(gPool + (size_t) slot * rs.L.W);
    if (WP::L == 1) {
      for (int i = 0; i < (rs.L.W >> 2); ++i) pf[i] = s[i];
    } else if (w.lane() < (rs.L.W >> 2)) {
      pf[0] = s[w.lane()];
    }
  }
  VCE_HD void takePrefetched(int32_t* p) {
    Quad* d = reinterpret_cast<Quad*>(p);
    if (WP::L == 1) {
      for (int i = 0; i < (rs.L.W >> 2); ++i) d[i] = pf[i];
    } else if (w.lane() < (rs.L.W >> 2)) {
      d[w.lane()] = pf[0];
    }
  }

  // ---------------------------------------------------------------------------------------------- frontier
  VCE_HD void frontierReset() {
    VCE_NOUNROLL
    for (int i = w.lane(); i < shape.nb; i += WP::L) blkFree[i] = (uint16_t) (shape.nb - 1 - i);  // pop order 0, 1, 2, ...
    w.storeWait();
    w.sync();
    nbFree = shape.nb - 1;   // block 0 taken
    nbUsed = 1;
    dh = 0;
    for (int c = 0; c < 3; ++c) { cid[c] = -1; cdirty[c] = 0; cuse[c] = 0; }
    useClock = 0;
    cid[0] = 0;
    cdirty[0] = 1;
    if (w.lane() == 0) {
      dir[0] = 0;
      uint32_t* B = cacheBlock(0);
      B[cta::B_START] = 0;
      B[cta::B_CNT] = 0;
      B[cta::B_FREE] = B[cta::B_FREE + 1] = B[cta::B_FREE + 2] = B[cta::B_FREE + 3] = 0xffffffffu;
    }
    nFreeCache = 0; nFreeStack = 0; bump = 0;
    for (int k = 0; k < cta::kMaxWork; ++k) tag[k] = -1;
    pfSlot = -1;
    w.sync();
  }
  VCE_HD static uint8_t* ordOf(uint32_t* B) { return reinterpret_cast<uint8_t*>(B + cta::B_ORD); }
  VCE_HD static int entryAt(uint32_t* B, int i) { return ordOf(B)[B[cta::B_START] + i]; }   // logical position -> entry
  VCE_HD static uint32_t* keyOf(uint32_t* B, int e) { return B + cta::B_KEY + e * cta::kKeyW; }
  VCE_HD static uint32_t* auxOf(uint32_t* B, int e) { return B + cta::B_AUX + e * cta::kAuxW; }

  // First i in [0, n) for which pred(i) holds (pred is monotone: false ... false true ... true), n if none.  Every round
  // the lanes probe the last elements of WP::L equal segments of the open range.
  template <class Pred>
  VCE_HD int lowerBound(int n, Pred pred) {
    int lo = 0, hi = n;   // everything below lo is false, everything from hi on is true
    while (lo < hi) {
      const int span = hi - lo;
      const int step = (span + WP::L - 1) / WP::L;
      const int first = lo + w.lane() * step;
      const bool act = first < hi;
      int j = first + step - 1;
      if (j > hi - 1) j = hi - 1;
      const bool t = act && pred(j);
      const unsigned m = w.ballot(t);
      if (m == 0) { lo = hi; break; }
      const int k = vffs(m);
      int nh = lo + k * step + step - 1;
      if (nh > hi - 1) nh = hi - 1;
      lo = lo + k * step;
      hi = nh;   // element nh is true
    }
    return lo;
  }
  // Directory index (0-based from the head) of the first block whose largest key is >= key; the last block if none.
  VCE_HD int findBlock(const uint32_t* key) {
    const uint32_t* sp = sep;
    const uint16_t* d = dir + dh;
    return lowerBound(nbUsed - 1, [sp, d, key](int i) { return keyCmp(sp + (size_t) d[i] * cta::kKeyW, key) >= 0; });
  }
  // Logical position in block B of the first entry whose key is >= key (cnt if none); eq: that entry's key equals key.
  VCE_HD int findInBlock(uint32_t* B, const uint32_t* key, bool& eq) {
    const int cnt = (int) B[cta::B_CNT];
    const int pos = lowerBound(cnt, [B, key](int i) { return keyCmp(keyOf(B, entryAt(B, i)), key) >= 0; });
    eq = pos < cnt && keyCmp(keyOf(B, entryAt(B, pos)), key) == 0;
    return pos;
  }

  VCE_HD void setSep(int id, const uint32_t* key) {
    VCE_NOUNROLL
    for (int i = w.lane(); i < cta::kKeyW; i += WP::L) sep[(size_t) id * cta::kKeyW + i] = key[i];
  }
  // Opens position `pos` (logical) in the order deque of block B; the caller writes the entry id there.
  VCE_HD void ordOpen(uint32_t* B, int pos) {
    uint8_t* ord = ordOf(B);
    int st = (int) B[cta::B_START];
    const int cnt = (int) B[cta::B_CNT];
    w.sync();   // every lane has read the header before lane 0 rewrites it (the shift loops below may be empty)
    if (st > 0 && pos <= cnt - pos) {  // move the front part down
      VCE_NOUNROLL
      for (int base = 0; base < pos; base += WP::L) {
        const int i = base + w.lane();
        int v = 0;
        const bool act = i < pos;
        if (act) v = ord[st + i];
        w.sync();
        if (act) ord[st + i - 1] = (uint8_t) v;
        w.sync();
      }
      if (w.lane() == 0) B[cta::B_START] = (uint32_t) (st - 1);
    } else {
      if (st + cnt >= 256) {  // no room at the top: compact to the bottom first
        VCE_NOUNROLL
        for (int base = 0; base < cnt; base += WP::L) {
          const int i = base + w.lane();
          int v = 0;
          const bool act = i < cnt;
          if (act) v = ord[st + i];
          w.sync();
          if (act) ord[i] = (uint8_t) v;
          w.sync();
        }
        st = 0;
        if (w.lane() == 0) B[cta::B_START] = 0;
      }
      VCE_NOUNROLL
      for (int base = ((cnt - pos - 1) / WP::L) * WP::L; base >= 0 && cnt > pos; base -= WP::L) {  // highest chunk first
        const int i = pos + base + w.lane();
        int v = 0;
        const bool act = i < cnt;
        if (act) v = ord[st + i];
        w.sync();
        if (act) ord[st + i + 1] = (uint8_t) v;
        w.sync();
      }
    }
    if (w.lane() == 0) B[cta::B_CNT] = (uint32_t) (cnt + 1);
    w.sync();
  }
  VCE_HD static int takeFreeEntry(uint32_t* B) {  // lane 0
    for (int q = 0; q < 4; ++q) {
      const uint32_t m = B[cta::B_FREE + q];
      if (m != 0) {
        const int b = vffs(m);
        B[cta::B_FREE + q] = m & ~(1u << b);
        return q * 32 + b;
      }
    }
    return -1;
  }
  // Opens the directory entry right after directory index bi for block id `id`.
  VCE_HD void dirInsertAfter(int bi, int id) {
    if (dh + nbUsed >= 2 * shape.nb) {  // out of room at the top: compact to the bottom
      VCE_NOUNROLL
      for (int base = 0; base < nbUsed; base += WP::L) {
        const int i = base + w.lane();
        int v = 0;
        const bool act = i < nbUsed;
        if (act) v = dir[dh + i];
        w.sync();
        if (act) dir[i] = (uint16_t) v;
        w.sync();
      }
      dh = 0;
    }
    const int from = bi + 1, to = nbUsed;
    VCE_NOUNROLL
    for (int base = ((to - from - 1) / WP::L) * WP::L; base >= 0 && to > from; base -= WP::L) {
      const int i = from + base + w.lane();
      int v = 0;
      const bool act = i < to;
      if (act) v = dir[dh + i];
      w.sync();
      if (act) dir[dh + i + 1] = (uint16_t) v;
      w.sync();
    }
    if (w.lane() == 0) dir[dh + bi + 1] = (uint16_t) id;
    ++nbUsed;
    w.sync();
  }

  // PathFinder.addIfBetter (:321-348) for the path record in work slot `ws`.  Returns false when the arena is exhausted.
  VCE_HD bool frontierInsert(int ws) {
    const int32_t* p = rs.P(ws);
    w.sync();
    makeKey(p, keyNew);
    w.sync();
    const uint32_t* auxNew = keyNew + cta::kKeyW;
    int bi = findBlock(keyNew);
    if (bi == 0) ++statInsFront; else ++statInsFar;
    VCE_CTA_CHECK(bi >= 0 && bi < nbUsed && dh >= 0 && dh + nbUsed <= 2 * shape.nb, 8, bi, nbUsed);
    int id = dir[dh + bi];
    int c = ensureCached(id, -1, true);
    uint32_t* B = cacheBlock(c);
    bool eq;
    VCE_CTA_CHECK((int) B[cta::B_CNT] >= 0 && (int) B[cta::B_CNT] <= cta::kBlk && (int) B[cta::B_START] + (int) B[cta::B_CNT] <= 256, 9, B[cta::B_CNT], B[cta::B_START]);
    int pos = findInBlock(B, keyNew, eq);
    VCE_CTA_CHECK(pos >= 0 && pos <= (int) B[cta::B_CNT], 10, pos, B[cta::B_CNT]);
    if (eq) {
      const int e = entryAt(B, pos);
      const bool newWins = betterAux(auxNew, auxOf(B, e));
      if (newWins) {
        // the new path takes over the entry and the record slot of the one it replaces
        const int slot = (int) B[cta::B_SLOT + e];
        invalidate(slot);
        w.sync();
        VCE_NOUNROLL
        for (int i = w.lane(); i < cta::kAuxW; i += WP::L) auxOf(B, e)[i] = auxNew[i];
        storeRecord(slot, p);
        if (ws < nWork) tag[ws] = slot;
        cdirty[c] = 1;
        w.sync();
      }
      return true;
    }
    if (rs.n >= shape.capacity()) return false;
    const int slot = allocRecord();
    if (slot < 0) return false;
    if ((int) B[cta::B_CNT] >= cta::kBlk) {
      // full: the upper half becomes a new block right after this one
      if (nbFree <= 0) return false;
      const int nid = blkFree[nbFree - 1];
      VCE_CTA_CHECK(nid >= 0 && nid < shape.nb && nbFree <= shape.nb, 15, nid, nbFree);
      --nbFree;
      const int c2 = ensureCached(nid, c, false);
      uint32_t* N = cacheBlock(c2);
      const int half = cta::kBlk / 2;
      // entries by order: logical half .. kBlk-1 move to entries 0 .. half-1 of the new block
      VCE_NOUNROLL
      for (int i = w.lane(); i < half; i += WP::L) {
        const int e = entryAt(B, half + i);
        N[cta::B_SLOT + i] = B[cta::B_SLOT + e];
        for (int q = 0; q < cta::kAuxW; ++q) auxOf(N, i)[q] = auxOf(B, e)[q];
        for (int q = 0; q < cta::kKeyW; ++q) keyOf(N, i)[q] = keyOf(B, e)[q];
        ordOf(N)[i] = (uint8_t) i;
      }
      w.sync();
      if (w.lane() == 0) {
        N[cta::B_START] = 0;
        N[cta::B_CNT] = (uint32_t) half;
        N[cta::B_FREE] = 0; N[cta::B_FREE + 1] = 0; N[cta::B_FREE + 2] = 0xffffffffu; N[cta::B_FREE + 3] = 0xffffffffu;
        // entries that left B are free again
        uint32_t fm[4] = {B[cta::B_FREE], B[cta::B_FREE + 1], B[cta::B_FREE + 2], B[cta::B_FREE + 3]};
        for (int i = 0; i < half; ++i) {
          const int e = entryAt(B, half + i);
          fm[e >> 5] |= 1u << (e & 31);
        }
        B[cta::B_FREE] = fm[0]; B[cta::B_FREE + 1] = fm[1]; B[cta::B_FREE + 2] = fm[2]; B[cta::B_FREE + 3] = fm[3];
        B[cta::B_CNT] = (uint32_t) half;
      }
      w.sync();
      // largest keys: the new block inherits the old one, the lower half ends at its entry half-1
      setSep(nid, sep + (size_t) id * cta::kKeyW);
      w.sync();
      setSep(id, keyOf(B, entryAt(B, half - 1)));
      dirInsertAfter(bi, nid);
      cdirty[c] = 1;
      cdirty[c2] = 1;
      if (pos > half) { pos -= half; bi += 1; id = nid; c = c2; B = N; }
    }
    const int cnt = (int) B[cta::B_CNT];
    int e = 0;
    if (w.lane() == 0) e = takeFreeEntry(B);
    e = w.bcast(e, 0);
    VCE_CTA_CHECK(e >= 0 && e < cta::kBlk, 11, e, cnt);
    ordOpen(B, pos);
    if (w.lane() == 0) {
      ordOf(B)[B[cta::B_START] + pos] = (uint8_t) e;
      B[cta::B_SLOT + e] = (uint32_t) slot;
    }
    VCE_NOUNROLL
    for (int i = w.lane(); i < cta::kKeyW; i += WP::L) keyOf(B, e)[i] = keyNew[i];
    VCE_NOUNROLL
    for (int i = w.lane(); i < cta::kAuxW; i += WP::L) auxOf(B, e)[i] = auxNew[i];
    if (pos == cnt) setSep(id, keyNew);   // the new path is the block's largest
    cdirty[c] = 1;
    storeRecord(slot, p);
    if (ws < nWork) tag[ws] = slot;
    ++rs.n;
    if (rs.n > rs.maxFrontier) rs.maxFrontier = rs.n;
    w.sync();
    return true;
  }

  // TreeSet.pollFirst: the least path's record ends up in work slot 0.  rs.n > 0.
  VCE_HD void popHead() {
    int id = dir[dh];
    int c = ensureCached(id, -1, true);
    uint32_t* B = cacheBlock(c);
    const int st = (int) B[cta::B_START], cnt = (int) B[cta::B_CNT];
    VCE_CTA_CHECK(cnt >= 1 && cnt <= cta::kBlk && st >= 0 && st + cnt <= 256, 12, cnt, st);
    const int e = ordOf(B)[st];
    VCE_CTA_CHECK(e >= 0 && e < cta::kBlk, 13, e, st);
    const int slot = (int) B[cta::B_SLOT + e];
    VCE_CTA_CHECK(slot >= 0 && slot < poolCap, 14, slot, e);
    // the record: still in a work slot, prefetched, or from the pool
    int hit = -1;
    for (int k = 0; k < nWork; ++k) if (tag[k] == slot) { hit = k; break; }
    w.sync();
    if (hit > 0) {
      rs.copyPath(0, hit);
    } else if (hit < 0) {
      if (pfSlot == slot) takePrefetched(rs.P(0));
      else loadRecord(rs.P(0), slot);
    }
    for (int k = 0; k < nWork; ++k) tag[k] = k == 0 ? -1 : tag[k];
    w.sync();
    if (w.lane() == 0) {
      B[cta::B_START] = (uint32_t) (cnt == 1 ? 0 : st + 1);
      B[cta::B_CNT] = (uint32_t) (cnt - 1);
      B[cta::B_FREE + (e >> 5)] |= 1u << (e & 31);
    }
    cdirty[c] = 1;
    --rs.n;
    freeRecord(slot);
    if (cnt == 1 && nbUsed > 1) {  // front block drained: release it
      if (w.lane() == 0) blkFree[nbFree] = (uint16_t) id;
      ++nbFree;
      ++dh;
      --nbUsed;
      dropCached(id);
      w.sync();
    }
    // request the record of the next front entry while this head is being processed
    pfSlot = -1;
    if (rs.n > 0) {
      id = dir[dh];
      int cc = -1;
      for (int q = 0; q < shape.ncache; ++q) if (cid[q] == id) cc = q;
      if (cc >= 0) {
        uint32_t* F = cacheBlock(cc);
        const int ns = (int) F[cta::B_SLOT + ordOf(F)[F[cta::B_START]]];
        bool inWork = false;
        for (int k = 0; k < nWork; ++k) if (tag[k] == ns) inWork = true;
        if (!inWork) prefetchRecord(ns);
      }
    }
  }

  // ---------------------------------------------------------------------------------------------- set-up
  // Shared-memory carve-up behind the RegionConst (see ctaSmemWords); returns the first word after the fixed part.
  VCE_HD uint32_t* carve(uint32_t* base) {
    using namespace cta;
    sep = base; base += (size_t) shape.nb * kKeyW;
    dir = reinterpret_cast<uint16_t*>(base);
    blkFree = dir + 2 * shape.nb;
    base += (2 * shape.nb + shape.nb + 1) / 2 + 2;
    base = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(base) + 15) & ~(uintptr_t) 15);
    cacheMem = base; base += (size_t) shape.ncache * B_WORDS;
    rs.slots = reinterpret_cast<int32_t*>(base); base += (kMaxWork + 2) * kMaxW;
    keyNew = base; base += kKeyW + kAuxW + 12;
    freeCache = reinterpret_cast<int32_t*>(base); base += kFreeCache;
    rs.ctl = reinterpret_cast<int32_t*>(base); base += 16 + 2 * cta::kMaxWork + 4;
    base = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(base) + 15) & ~(uintptr_t) 15);
    return base;
  }
  VCE_HD void bindArena(int32_t* arena, long long words) {
    gBlocks = reinterpret_cast<uint32_t*>(arena);
    const long long blockWords = (long long) shape.nb * cta::B_WORDS;
    const long long rest = words - blockWords - 64;
    long long cap = rest / (rs.L.W + 1);
    if (cap > 0x7fffff00LL) cap = 0x7fffff00LL;
    poolCap = cap > 0 ? (int) cap : 0;
    gPool = arena + blockWords;
    gFree = gPool + (size_t) poolCap * rs.L.W;
  }

  // ---------------------------------------------------------------------------------------------- main loop
  // PathFinder.bestPath :129-214 -- RegionSearch::run with this frontier.  Work slots: 0 = head, 1 .. maxChildren = the
  // children, then the last-sync path and the best finished path (rs.slotLS(), rs.slotBest()).
  VCE_HD int run(int& resultSlot) {
    typedef RS R_;
    const PathLayout& L = rs.L;
    rs.n = 0; rs.iters = 0; rs.curMaxPos = 0; rs.lastSyncPos = 0; rs.bestValid = 0; rs.usedPrefix = 0; rs.nWarn = 0;
    rs.maxFrontier = 1; rs.totalIters = 0;
    resultSlot = -1;
    nWork = 1 + rs.maxChildren;
    rs.nSlots = nWork;   // slotLS = nWork, slotBest = nWork + 1
    if (L.H > 2 || L.W > cta::kMaxW || nWork > cta::kMaxWork || unsupported) return kRegionOverflow;
    frontierReset();
    if (w.lane() == 0) {
      VCE_NOUNROLL
      for (int i = 0; i < 16; ++i) rs.ctl[i] = 0;
      rs.initPath(rs.P(0));
    }
    w.sync();
    rs.n = 1;
    initialPending = 1;

    while (rs.n > 0) {
      ++rs.iters;
      ++rs.totalIters;
      if (initialPending) { initialPending = 0; --rs.n; }
      else popHead();
      const int head = 0;
      bool headIsLS = false;
      if (rs.n == 0) {  // :151-161 only one path in play
        rs.lastSyncPos = rs.sidePos(rs.P(head), 0);
        rs.iters = 0;
        headIsLS = true;
      } else if (rs.n > rs.cfg.maxPaths || rs.iters > rs.cfg.maxIterations) {  // :162-169 too complex
        if (w.lane() == 0 && rs.warnOut != nullptr && rs.nWarn < rs.warnCap) {
          rs.warnOut[rs.nWarn] = WarnRec{rs.regionId, rs.n, rs.iters, rs.lastSyncPos + 1, rs.curMaxPos + 2};
        }
        ++rs.nWarn;
        w.sync();
        frontierReset();
        rs.n = 0;
        rs.iters = 0;
        rs.copyPath(head, rs.slotLS());
        w.sync();
        if (w.lane() == 0) {
          rs.skipVariantsTo(rs.P(head), 0, rs.curMaxPos + 1);
          rs.skipVariantsTo(rs.P(head), 1, rs.curMaxPos + 1);
        }
        w.sync();
        if (rs.ctl[R_::CTL_FLAGS] & R_::FL_ERRMOVE) return kRegionErrMove;
        headIsLS = true;
      }

      // ---- lane 0 decides what happens to the head (same decisions as RegionSearch::run)
      if (w.lane() == 0) {
        int32_t* hp = rs.P(head);
        int act = 0, side = 0, vidx = -1;
        const int cv = rs.nextVariant(hp, 0);   // :170 calls first
        if (cv >= 0) { act = R_::ACT_ENQUEUE; side = 0; vidx = cv; }
        else if (cv == -2) { act = R_::ACT_BOUNDARY_SPILL; }
        else {
          const int bv = rs.nextVariant(hp, 1);  // :173 then baseline
          if (bv >= 0) { act = R_::ACT_ENQUEUE; side = 1; vidx = bv; }
          else if (bv == -2) { act = R_::ACT_BOUNDARY_SPILL; }
        }
        if (act == R_::ACT_BOUNDARY_SPILL) {
          const int cn = rs.R.cNextStart, bn = rs.R.bNextStart;
          const int boundary = (cn < bn ? cn : bn) - 1;
          if (rs.n == 0 && rs.inSync(hp) && rs.sidePos(hp, 0) == boundary
              && hp[L.sideBase(0) + SD_VARIDX] + 1 >= rs.R.cCount && hp[L.sideBase(1) + SD_VARIDX] + 1 >= rs.R.bCount) {
            act = R_::ACT_BOUNDARY_CLEAN;
          }
        } else if (act == R_::ACT_ENQUEUE) {
          const int vg = rs.sideFirst(side) + vidx;
          const int st = rs.S[side].vStart[vg];
          rs.ctl[R_::CTL_AUX] = st;
          rs.ctl[R_::CTL_AUX + 1] = 0;
          if (rs.inSync(hp)) {  // Path.addVariant :288-300
            int32_t* pb = hp + L.pathBase();
            const int syncPos = rs.sidePos(hp, 0);
            const int ns = pb[PT_NSYNC];
            if (!(ns > 0 && pb[PT_SYNC0 + ns - 1] == syncPos)) {
              if (ns >= L.SMAX) rs.flag(R_::FL_OVERFLOW);
              else { pb[PT_SYNC0 + ns] = syncPos; pb[PT_NSYNC] = ns + 1; rs.ctl[R_::CTL_AUX + 1] = 1; }
            }
            pb[PT_CSINCE] = 0;
            pb[PT_BSINCE] = 0;
          }
        } else {
          rs.pathStep(hp);  // :177
          act = R_::ACT_STEP_DROP;
          bool alive = true;
          if (rs.inSync(hp)) {  // :182-190
            if (rs.cfg.pruneNoOps && rs.hasNoOp(hp)) alive = false;
            else rs.skipToNextVariant(hp);
          }
          if (alive && rs.pathMatches(hp)) {  // :192
            if (rs.sideFinished(hp, 0) && rs.sideFinished(hp, 1)) {  // :193
              if (!(rs.cfg.pruneNoOps && rs.hasNoOp(hp))) {
                int32_t* pb = hp + L.pathBase();
                const int ns = pb[PT_NSYNC];
                if (ns >= L.SMAX) rs.flag(R_::FL_OVERFLOW);
                else { pb[PT_SYNC0 + ns] = rs.sidePos(hp, 0); pb[PT_NSYNC] = ns + 1; }  // :199
                act = R_::ACT_STEP_FINISH;
              }
            } else {
              act = R_::ACT_STEP_REINSERT;
            }
          }
        }
        rs.ctl[R_::CTL_ACT] = act;
        rs.ctl[R_::CTL_SIDE] = side;
        rs.ctl[R_::CTL_VIDX] = vidx;
      }
      w.sync();
      const int act = rs.ctl[R_::CTL_ACT];
      const int fl = rs.ctl[R_::CTL_FLAGS];
      if (fl & R_::FL_ERRORDER) return kRegionErrOrder;
      if (fl & R_::FL_ERRMOVE) return kRegionErrMove;
      if (fl & R_::FL_OVERFLOW) return kRegionOverflow;
      if (fl & R_::FL_WINMISS) return kRegionWinMiss;

      if (act == R_::ACT_BOUNDARY_CLEAN) { resultSlot = head; return kRegionClean; }
      if (act == R_::ACT_BOUNDARY_SPILL) return kRegionSpill;

      if (act == R_::ACT_ENQUEUE) {
        const int side = rs.ctl[R_::CTL_SIDE];
        const int vidx = rs.ctl[R_::CTL_VIDX];
        const int vg = rs.sideFirst(side) + vidx;
        const int st = rs.ctl[R_::CTL_AUX];
        rs.curMaxPos = rs.curMaxPos > st ? rs.curMaxPos : st;  // :236
        const int row0 = rs.S[side].oOff[vg];
        const int nOr = rs.S[side].oOff[vg + 1] - row0;
        const int nChild = 1 + nOr;
        if (nChild > rs.maxChildren) return kRegionOverflow;
        for (int k = 0; k < nChild; ++k) { tag[1 + k] = -1; rs.copyPath(1 + k, head); }
        w.sync();
        VCE_NOUNROLL
        for (int k = w.lane(); k < nChild; k += WP::L) {
          int32_t* cp = rs.P(1 + k);
          int32_t* sb = cp + L.sideBase(side);
          const int vend = rs.S[side].vEnd[vg];
          int valid = 1;
          if (k == 0) {  // HalfPath.exclude :153-158
            sb[SD_VAREND] = sb[SD_VAREND] > vend ? sb[SD_VAREND] : vend;
            sb[SD_VARIDX] = vidx;
            rs.setDecision(cp, side, vidx, 1);
          } else {
            const int row = row0 + k - 1;
            if (rs.sideIsNew(rs.P(head), side, vg, row)) {  // HalfPath.include :160-170, Path.include :267-275
              sb[SD_VAREND] = sb[SD_VAREND] > vend ? sb[SD_VAREND] : vend;
              sb[SD_VARIDX] = vidx;
              sb[SD_INCLEND] = sb[SD_INCLEND] > vend ? sb[SD_INCLEND] : vend;
              sb[SD_INCLCNT] += 1;
              sb[SD_LASTID] = rs.S[side].oIds[(size_t) row * 4];
              rs.setDecision(cp, side, vidx, 1 + k);
              VCE_NOUNROLL
              for (int h = 0; h < L.H; ++h) rs.hapAdd(cp + L.hap(side, h), side, rs.S[side].oSlot[(size_t) row * L.H + h]);
              cp[L.pathBase() + (side == 0 ? PT_CSINCE : PT_BSINCE)] += 1;
            } else {
              valid = 0;
            }
          }
          rs.childValid(k) = valid;
        }
        w.sync();
        if (rs.ctl[R_::CTL_FLAGS] & R_::FL_OVERFLOW) return kRegionOverflow;
        if (headIsLS) {
          // lastSyncPath is the PARENT object (see RegionSearch::run)
          rs.copyPath(rs.slotLS(), head);
          w.sync();
          if (w.lane() == 0 && rs.ctl[R_::CTL_AUX + 1]) rs.P(rs.slotLS())[L.pathBase() + PT_NSYNC] -= 1;
          w.sync();
        }
        VCE_NOUNROLL
        for (int k = 0; k < nChild; ++k) {
          if (rs.childValid(k)) {
            if (!frontierInsert(1 + k)) return kRegionOverflow;
          }
        }
        continue;
      }
      if (act == R_::ACT_STEP_DROP) continue;
      if (act == R_::ACT_STEP_FINISH) {  // :199-201
        bool take = true;
        if (rs.bestValid) take = !rs.firstIsBetter(rs.P(rs.slotBest()), rs.P(head));
        w.sync();
        if (take) { rs.copyPath(rs.slotBest(), head); rs.bestValid = 1; }
        w.sync();
        continue;
      }
      // ACT_STEP_REINSERT :203
      if (!frontierInsert(head)) return kRegionOverflow;
    }
    if (rs.R.cNextStart != kNoNext || rs.R.bNextStart != kNoNext) return kRegionSpill;
    if (!rs.bestValid) return kRegionErrNull;
    resultSlot = rs.slotBest();
    return kRegionFinished;
  }
};

}  // namespace vce
#endif
