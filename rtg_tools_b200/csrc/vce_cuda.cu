// vce_cuda.cu -- the CUDA backend: device buffers, K1 (orientation expansion), K2/K3 (region search with the
// tie-break / weight epilogue) for sm_100a.  Compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
//
// Kernels
//   k_orient_count / k_orient_fill : one thread per variant, Orientor.orientations (vce_orient.h)
//   k_search<false>                : one warp per region, path records + frontier in shared memory, the region's
//                                    reference window staged into shared memory, fixed small capacities
//   k_search<true>                 : one warp per region, per-warp arena in HBM, capacities up to the reference's
//                                    own search budget (max-paths / max-iterations)
// Both search kernels pull regions from a global atomic work counter so that long clusters do not stall a block.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <string>
#include <vector>

#include <mutex>

#include "vce_cuda.h"
#include "vce_micro.h"
#include "vce_orient.h"
#include "vce_search.h"
#include "vce_cta.h"
#include "vce_final_cuda.cuh"

namespace vce {

// ------------------------------------------------------------------------------------------------ K1
struct OrientParams {
  int n, kind, H, gtPloidy;
  const int8_t* gt;
  const uint8_t* vFlags;
  const int32_t* slot0;
  const int32_t* nSl;
  const uint8_t* aNull;
};

__device__ __forceinline__ VariantGt loadGt(const OrientParams& p, int v) {
  VariantGt g;
  g.gtPloidy = p.gtPloidy;
  const char4 q = reinterpret_cast<const char4*>(p.gt)[v];
  g.gt[0] = q.x; g.gt[1] = q.y; g.gt[2] = q.z; g.gt[3] = q.w;
  g.phased = p.vFlags[v] & 1;
  g.slot0 = p.slot0[v];
  g.nSlots = p.nSl[v];
  g.aNull = p.aNull;
  return g;
}

__global__ void k_orient_count(const OrientParams p, int32_t* __restrict__ cnt, int32_t* __restrict__ maxOut) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  if (v < p.n) {
    const VariantGt g = loadGt(p, v);
    OrientCount oc;
    orientVariant(p.kind, p.H, g, oc);
    c = oc.n;
    cnt[v] = c;
  } else if (v == p.n) {
    cnt[v] = 0;
  }
  // block max -> one atomic per warp
  for (int o = 16; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
  if ((threadIdx.x & 31) == 0 && c > 0) atomicMax(maxOut, c);
}

__global__ void k_orient_fill(const OrientParams p, const int32_t* __restrict__ oOff, int32_t* __restrict__ oSlot,
                              int8_t* __restrict__ oIds, uint8_t* __restrict__ oOrig) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= p.n) return;
  const VariantGt g = loadGt(p, v);
  OrientFill f{oOff[v], p.H, g.slot0, g.nSlots, g.aNull, oSlot, oIds, oOrig};
  orientVariant(p.kind, p.H, g, f);
}

// ------------------------------------------------------------------------------------------------ K2 / K3
struct SearchParams {
  const RegionDesc* regs;
  int nRegs;
  SideArrays S[2];
  const uint8_t* windows;
  SearchConfig cfg;
  RegionResult* out;
  const int32_t* syncOff;
  int32_t* syncOut;
  WarnRec* warns;        // global list
  int* warnCount;
  int warnCap;
  WarnRec* warnScratch;  // kWarnPerRegion per warp of the grid
  int* counter;          // work queue
  int H;
  int maxChildren;
  int32_t* arena;        // general kernel only
  long long arenaWordsPerWarp;
  int winSmemBytes;      // shared memory per warp for the reference window
  FastShape fs;          // shared-memory kernel: capacities of the tier
  int wordsPerWarp;      // shared-memory words per warp (whole layout)
  int tabBytes;          // shared memory per warp for the region's tables
  int regionBase;        // index of regs[0] within the batch (warnings carry batch-wide region indices)
};

constexpr int kWarnPerRegion = 64;
constexpr int kFastWarps = 8;
#ifndef VCE_GEN_MINB
#define VCE_GEN_MINB 3   // resident blocks per SM the large-capacity kernel is compiled for
#endif

// Copies the region's slices of the variant / orientation / allele tables into shared memory and points the search at
// them (the pointers are re-based so that the search keeps indexing with the batch-wide indices).  The serial part of
// the search is a chain of dependent table reads: from HBM each costs ~700 cycles, from shared memory ~30.
// Returns false (nothing changed) when the slices do not fit `cap` bytes.
template <class RS>
__device__ __noinline__ bool stageTables(RS& rs, const SideArrays* G, int H, uint8_t* buf, int cap, int lane) {
  const RegionDesc& R = rs.R;
  int need = 0;
  int rows0[2], rowsN[2], nt0[2], ntN[2];
  for (int side = 0; side < 2; ++side) {
    const int first = side == 0 ? R.cFirst : R.bFirst, cnt = side == 0 ? R.cCount : R.bCount;
    const int sl0 = side == 0 ? R.cSlot0 : R.bSlot0, slN = side == 0 ? R.cSlotN : R.bSlotN;
    rows0[side] = cnt > 0 ? G[side].oOff[first] : 0;
    rowsN[side] = cnt > 0 ? G[side].oOff[first + cnt] - rows0[side] : 0;
    nt0[side] = slN > 0 ? G[side].aNtOff[sl0] : 0;
    ntN[side] = slN > 0 ? G[side].aNtOff[sl0 + slN] - nt0[side] : 0;
    need += cnt * 8 + (cnt + 1) * 4 + ((cnt + 3) & ~3);            // vStart, vEnd, oOff, vFlags
    need += rowsN[side] * H * 4 + rowsN[side] * 4;                   // oSlot, oIds
    need += slN * 8 + (slN + 1) * 4 + ((ntN[side] + 3) & ~3);        // aStart, aEnd, aNtOff, nt
  }
  if (need > cap) return false;
  uint8_t* at = buf;
  for (int side = 0; side < 2; ++side) {
    const int first = side == 0 ? R.cFirst : R.bFirst, cnt = side == 0 ? R.cCount : R.bCount;
    const int sl0 = side == 0 ? R.cSlot0 : R.bSlot0, slN = side == 0 ? R.cSlotN : R.bSlotN;
    const SideArrays& g = G[side];
    SideArrays& S = rs.S[side];
    int32_t* vS = reinterpret_cast<int32_t*>(at); at += cnt * 4;
    int32_t* vE = reinterpret_cast<int32_t*>(at); at += cnt * 4;
    int32_t* oO = reinterpret_cast<int32_t*>(at); at += (cnt + 1) * 4;
    int32_t* oS = reinterpret_cast<int32_t*>(at); at += rowsN[side] * H * 4;
    int32_t* aS = reinterpret_cast<int32_t*>(at); at += slN * 4;
    int32_t* aE = reinterpret_cast<int32_t*>(at); at += slN * 4;
    int32_t* aN = reinterpret_cast<int32_t*>(at); at += (slN + 1) * 4;
    int8_t* oI = reinterpret_cast<int8_t*>(at); at += rowsN[side] * 4;
    uint8_t* vF = at; at += (cnt + 3) & ~3;
    uint8_t* nt = at; at += (ntN[side] + 3) & ~3;
    _Pragma("unroll 1")
    for (int i = lane; i < cnt; i += 32) { vS[i] = g.vStart[first + i]; vE[i] = g.vEnd[first + i]; vF[i] = g.vFlags[first + i]; }
    _Pragma("unroll 1")
    for (int i = lane; i <= cnt; i += 32) oO[i] = cnt > 0 ? g.oOff[first + i] : 0;
    _Pragma("unroll 1")
    for (int i = lane; i < rowsN[side] * H; i += 32) oS[i] = g.oSlot[(size_t) rows0[side] * H + i];
    _Pragma("unroll 1")
    for (int i = lane; i < rowsN[side] * 4; i += 32) oI[i] = g.oIds[(size_t) rows0[side] * 4 + i];
    _Pragma("unroll 1")
    for (int i = lane; i < slN; i += 32) { aS[i] = g.aStart[sl0 + i]; aE[i] = g.aEnd[sl0 + i]; }
    _Pragma("unroll 1")
    for (int i = lane; i <= slN; i += 32) aN[i] = slN > 0 ? g.aNtOff[sl0 + i] : 0;
    _Pragma("unroll 1")
    for (int i = lane; i < ntN[side]; i += 32) nt[i] = g.nt[nt0[side] + i];
    if (lane == 0) {
      S.vStart = vS - first; S.vEnd = vE - first; S.vFlags = vF - first; S.oOff = oO - first;
      S.oSlot = oS - (ptrdiff_t) rows0[side] * H; S.oIds = oI - (ptrdiff_t) rows0[side] * 4;
      S.aStart = aS - sl0; S.aEnd = aE - sl0; S.aNtOff = aN - sl0; S.nt = nt - nt0[side];
    }
  }
  __syncwarp();
  return true;
}

template <bool GENERAL>
__global__ void __launch_bounds__(kFastWarps * 32, GENERAL ? VCE_GEN_MINB : 1) k_search(const __grid_constant__ SearchParams p) {
  extern __shared__ __align__(16) int32_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;

  // per-warp shared memory: [RegionConst | tier layout ...] (fast tiers) or [RegionConst | tables] (general)
  const int kcWords = (int) ((sizeof(RegionConst) + 15) / 16) * 4;
  int32_t* warpBase = GENERAL ? smem + (size_t) warp * (kcWords + p.tabBytes / 4) : smem + (size_t) warp * p.wordsPerWarp;
  RegionConst* kc = reinterpret_cast<RegionConst*>(warpBase);
  RegionSearch<DeviceWarp, GENERAL> rs(kc);
  if (lane == 0) kc->cfg = p.cfg;
  rs.warnOut = p.warnScratch + (size_t) gwarp * kWarnPerRegion;
  rs.warnCap = kWarnPerRegion;
  uint8_t* winS = nullptr;
  uint8_t* tabS = nullptr;
  if (!GENERAL) {
    if (lane == 0) kc->L.init(p.H, p.fs.q, p.fs.kv, p.fs.smax, 4);
    __syncwarp();
    rs.maxChildren = p.fs.maxChildren;
    rs.cap = p.fs.cap;
    rs.nSlots = p.fs.slots();
    rs.ordCap = p.fs.ordCap();
    PathLayout LM;
    LM.init(2, p.fs.q, p.fs.kv, p.fs.smax, 4);
    // ctl is padded so that what follows it (tables, window, the next warp's block) stays 16-byte aligned
    const int preWords = (rs.nSlots + 2) * LM.W + rs.ordCap + rs.nSlots;
    const int ctlWords = ((preWords + 16 + 2 * p.fs.maxChildren + 3) & ~3) - preWords;
    int32_t* base = warpBase + kcWords;
    rs.slots = base;
    rs.order = base + (rs.nSlots + 2) * LM.W;
    rs.freeList = rs.order + rs.ordCap;
    rs.ctl = rs.freeList + rs.nSlots;
    tabS = reinterpret_cast<uint8_t*>(rs.ctl + ctlWords);
    winS = tabS + p.tabBytes;
  } else {
    tabS = reinterpret_cast<uint8_t*>(warpBase + kcWords);
  }

  while (true) {
    int r = 0;
    if (lane == 0) r = atomicAdd(p.counter, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= p.nRegs) break;
    __syncwarp();
    if (lane == 0) {
      kc->R = p.regs[r];
      kc->S[0] = p.S[0];
      kc->S[1] = p.S[1];
      if (GENERAL) kc->L.init(p.H, kc->R.qMax, max(1, max(kc->R.cCount, kc->R.bCount)), kc->R.cCount + kc->R.bCount + 2, kc->R.decBits);
    }
    rs.regionId = p.regionBase + r;
    __syncwarp();
    stageTables(rs, p.S, p.H, tabS, p.tabBytes, lane);
    if (GENERAL) {
      rs.maxChildren = p.maxChildren;
      const long long ctlWords = 16 + 2 * p.maxChildren;
      // capacity that fits this warp's arena:  (nSlots+2)*W + ordCap + nSlots + ctl <= arenaWords,
      // nSlots = cap + 1 + MC, ordCap = chunkWords(cap) <= 3*cap + 1024 (two-level frontier)
      long long cap = p.cfg.maxPaths + 1 + p.maxChildren;
      const long long fixed = (long long) (3 + p.maxChildren) * rs.L.W + 1024 + (1 + p.maxChildren) + ctlWords;
      const long long fit = (p.arenaWordsPerWarp - fixed) / (rs.L.W + 4);
      if (cap > fit) cap = fit;
      if (cap < 1) cap = 1;
      rs.cap = (int) cap;
      rs.nSlots = rs.cap + 1 + p.maxChildren;
      rs.ordCap = RegionSearch<DeviceWarp, GENERAL>::chunkWords(rs.cap);
      int32_t* base = p.arena + (size_t) gwarp * p.arenaWordsPerWarp;
      rs.slots = base;
      rs.order = base + (size_t) (rs.nSlots + 2) * rs.L.W;
      rs.freeList = rs.order + rs.ordCap;
      rs.ctl = rs.freeList + rs.nSlots;
      rs.win = p.windows + rs.R.winOff;
    } else {
      // stage the reference window in shared memory (16-byte vector loads; the window buffer is 16 B aligned and padded)
      const uint8_t* gw = p.windows + rs.R.winOff;
      if (rs.R.winLen <= p.winSmemBytes) {
        const int nvec = (rs.R.winLen + 15) >> 4;
        const uint4* src = reinterpret_cast<const uint4*>(gw);
        uint4* dst = reinterpret_cast<uint4*>(winS);
        for (int i = lane; i < nvec; i += 32) dst[i] = __ldg(src + i);
        rs.win = winS;
      } else {
        rs.win = gw;
      }
    }
    __syncwarp();
    int resultSlot = -1;
    const int st = rs.run(resultSlot);
    RegionResult rr;
    rr.status = st;
    rr.usedPrefix = rs.usedPrefix;
    rr.nSync = 0;
    rr.nWarn = rs.nWarn;
    rr.lastBaseAlleleId = kPrefixEmpty;
    rr.maxFrontier = rs.maxFrontier;
    rr.iterations = rs.totalIters;
    rr.pad = 0;
    __syncwarp();
    if (st == kRegionClean || st == kRegionFinished) {
      rs.writeResults(resultSlot, p.syncOut + p.syncOff[r], &rr);
      if (lane == 0 && rs.nWarn > 0) {
        const int nw = min(rs.nWarn, kWarnPerRegion);
        const int at = atomicAdd(p.warnCount, nw);
        for (int i = 0; i < nw; ++i) if (at + i < p.warnCap) p.warns[at + i] = rs.warnOut[i];
      }
    }
    if (lane == 0) p.out[r] = rr;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ K2 / K3, one region per CTA
// Warp policy of the CTA kernel (vce_cta.h): frontier blocks move between the CTA's HBM arena and its shared-memory cache
// as 1-D bulk TMA copies (cp.async.bulk, completion on an mbarrier for loads, bulk groups for stores); tma = 0 keeps plain
// cooperative 16-byte copies (diagnostics).
struct DeviceWarpMove : DeviceWarp {
  uint32_t mbar;     // shared-memory address of the CTA's mbarrier
  uint32_t phase;
  int tma;
  __device__ __forceinline__ void loadBlock(uint32_t* dst, const uint32_t* src, int bytes) {
    if (tma) {
      const uint32_t d = (uint32_t) __cvta_generic_to_shared(dst);
      // the slot was last touched through the generic proxy; earlier bulk stores of this block have to be complete
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane() == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"((uint32_t) bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(d), "l"(src), "r"((uint32_t) bytes), "r"(mbar) : "memory");
      }
      uint32_t done = 0;
      while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
      }
      phase ^= 1u;
    } else {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      __syncwarp();
      for (int i = lane(); i < (bytes >> 4); i += 32) d4[i] = s4[i];
    }
  }
  __device__ __forceinline__ void storeBlock(uint32_t* dst, const uint32_t* src, int bytes) {
    if (tma) {
      const uint32_t s = (uint32_t) __cvta_generic_to_shared(src);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of every lane, before the async read
      __syncwarp();
      if (lane() == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(s), "r"((uint32_t) bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      for (int i = lane(); i < (bytes >> 4); i += 32) d4[i] = s4[i];   // lane i <-> chunk i, as loadBlock reads them back
    }
  }
  // The shared-memory sources of earlier storeBlock calls may be overwritten after this.
  __device__ __forceinline__ void storeWait() {
    if (tma) {
      if (lane() == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
};

struct CtaParams {
  SearchParams sp;
  CtaShape shape;
  long long arenaWordsPerCta;
  int tma;
};

__global__ void __launch_bounds__(32) k_cta(const __grid_constant__ CtaParams cp) {
  extern __shared__ __align__(16) int32_t smem[];
  const SearchParams& p = cp.sp;
  const int lane = threadIdx.x & 31;
  const int kcWords = (int) ((sizeof(RegionConst) + 15) / 16) * 4;
  RegionConst* kc = reinterpret_cast<RegionConst*>(smem);
  CtaSearch<DeviceWarpMove> cs(kc);
  cs.shape = cp.shape;
  uint32_t* rest = cs.carve(reinterpret_cast<uint32_t*>(smem + kcWords));
  uint64_t* mbarP = reinterpret_cast<uint64_t*>(rest);
  rest += 4;
  uint8_t* tabS = reinterpret_cast<uint8_t*>(rest);
  uint8_t* winS = tabS + cp.shape.tabBytes;
  cs.w.mbar = (uint32_t) __cvta_generic_to_shared(mbarP);
  cs.w.phase = 0;
  cs.w.tma = cp.tma;
  cs.rs.w = cs.w;
  if (lane == 0) {
    kc->cfg = p.cfg;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(cs.w.mbar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  cs.rs.warnOut = p.warnScratch + (size_t) blockIdx.x * kWarnPerRegion;
  cs.rs.warnCap = kWarnPerRegion;
  cs.rs.maxChildren = p.maxChildren;

  while (true) {
    int r = 0;
    if (lane == 0) r = atomicAdd(p.counter, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= p.nRegs) break;
    __syncwarp();
    if (lane == 0) {
      kc->R = p.regs[r];
      kc->S[0] = p.S[0];
      kc->S[1] = p.S[1];
    }
    __syncwarp();
    bool ok = true;
    {
      PathLayout PL;
      ok = ctaLayout(PL, p.H, kc->R.cCount, kc->R.bCount, kc->R.decBits ? kc->R.decBits : 4, ctaNarrowRanks(kc->R.cSlotN, kc->R.bSlotN));
      if (lane == 0) kc->L = PL;
    }
    cs.rs.regionId = p.regionBase + r;
    // the allele ranks take the head of the table area; the region's tables follow (or stay in HBM when they do not fit)
    const int nSl = kc->R.cSlotN + kc->R.bSlotN;
    const int rankBytes = (nSl * 2 + 15) & ~15;
    if (rankBytes > cp.shape.tabBytes) ok = false;
    cs.rank[0] = reinterpret_cast<uint16_t*>(tabS);
    cs.rank[1] = cs.rank[0] + kc->R.cSlotN;
    __syncwarp();
    int st = kRegionOverflow;
    int resultSlot = -1;
    cs.rs.usedPrefix = 0; cs.rs.nWarn = 0; cs.rs.maxFrontier = 0; cs.rs.totalIters = 0;
    if (ok) {
      stageTables(cs.rs, p.S, p.H, tabS + rankBytes, cp.shape.tabBytes - rankBytes, lane);
      const uint8_t* gw = p.windows + cs.rs.R.winOff;
      if (cs.rs.R.winLen > 0 && cs.rs.R.winLen <= cp.shape.winBytes) {   // reference window -> shared memory (the buffer is 16 B aligned and padded)
        cs.w.loadBlock(reinterpret_cast<uint32_t*>(winS), reinterpret_cast<const uint32_t*>(gw), (cs.rs.R.winLen + 15) & ~15);
        cs.rs.win = winS;
      } else {
        cs.rs.win = gw;
      }
      __syncwarp();
      cs.bindArena(p.arena + (size_t) blockIdx.x * cp.arenaWordsPerCta, cp.arenaWordsPerCta);
      cs.computeRanks();
      st = cs.run(resultSlot);
    }
    RegionResult rr;
    rr.status = st;
    rr.usedPrefix = cs.rs.usedPrefix;
    rr.nSync = 0;
    rr.nWarn = cs.rs.nWarn;
    rr.lastBaseAlleleId = kPrefixEmpty;
    rr.maxFrontier = cs.rs.maxFrontier;
    rr.iterations = cs.rs.totalIters;
    rr.pad = 0;
    __syncwarp();
    if (st == kRegionClean || st == kRegionFinished) {
      cs.rs.writeResults(resultSlot, p.syncOut + p.syncOff[r], &rr);
      if (lane == 0 && cs.rs.nWarn > 0) {
        const int nw = min(cs.rs.nWarn, kWarnPerRegion);
        const int at = atomicAdd(p.warnCount, nw);
        for (int i = 0; i < nw; ++i) if (at + i < p.warnCap) p.warns[at + i] = cs.rs.warnOut[i];
      }
    }
    if (lane == 0) p.out[r] = rr;
    // no bulk copy of this region may still be reading the cache when the next region resets it
    cs.w.storeWait();
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ K1+K2+K3, one region per thread
// Thread-per-region search (vce_micro.h).  Every lane owns a shared-memory slice; a lane that settles its region
// takes the next one (grid-stride), so the warp keeps all lanes busy on "one search iteration" per loop trip while
// the regions of its lanes are at different stages.
struct MicroChunkDesc {
  const uint8_t* packets;
  const uint32_t* offsets;
  uint8_t* results;
  int n;
  int first;   // global index of the chunk's first region within the launch
  const uint32_t* order;   // optional: position in the launch order -> packet index (grouped by shape class)
};
struct MicroParams {
  MicroChunkDesc one;            // single-chunk launch
  const MicroChunkDesc* table;   // multi-chunk launch (nullptr = use `one`)
  int nChunks;
  int total;
  int refill;                    // idle lanes of a warp take new regions once at least this many lanes are idle
  MicroConfig cfg;
};
constexpr int kMicroThreads = 64;
constexpr uint32_t kNoPacket = 0xffffffffu;   // offsets[] entry of a region the builder could not pack
constexpr int kClassBins = 0x8000;             // shape classes (microPacketClass is 15 bits)

// Device-side digest (vce_packet.h): one thread packs one region's packet from the uploaded slices of the caller's arrays.
struct BuildParams {
  PacketView pv;              // pointers into the device copy of the slab, moved back by the slices' first indices
  const RegRange* regs;
  const uint32_t* winOff;
  const uint8_t* windows;
  uint8_t* packets;           // [n * stride]
  uint32_t* offsets;          // [n] packet offset in 4-byte words, kNoPacket when the region does not fit
  uint32_t* keys;             // optional [n]: shape class of each packet (kClassBins - 1 for the rejected ones)
  int n, stride;
  const uint32_t* t2;         // resident template (vce_template_register): 2 bits per base, nullptr = windows[] holds bytes
  const uint32_t* nmask;      // one bit per 16-base granule with a non-ACGT base
};
template <class T>
__global__ void __launch_bounds__(128) k_build(const __grid_constant__ BuildParams p) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.n) return;
  const RegRange g = p.regs[q];
  int32_t winStart = 0, winLen = 0, nextStart[2];
  int bytes = 0;
  uint8_t* out = p.packets + (size_t) q * p.stride;
  if (microWindow(p.pv, g, &winStart, &winLen, nextStart)) {
    if (p.t2 != nullptr) bytes = microBuildPacketW<T>(p.pv, g, winStart, winLen, nextStart, WinPacked{p.t2, p.nmask, winStart}, out);
    else bytes = microBuildPacket<T>(p.pv, g, winStart, winLen, nextStart, p.windows + p.winOff[q], out);
  }
  p.offsets[q] = bytes > 0 ? (uint32_t) (((size_t) q * p.stride) >> 2) : kNoPacket;
  if (p.keys != nullptr) p.keys[q] = bytes > 0 ? microPacketClass(out) : (uint32_t) (kClassBins - 1);
}

// Counting sort of one chunk's packets by shape class, one block: order[] lists the packet indices class by class, so
// that the 32 regions a warp of k_micro takes together behave alike (the host digest does the same for staged jobs).
constexpr int kOrderThreads = 1024;
__global__ void __launch_bounds__(kOrderThreads) k_class_order(const uint32_t* __restrict__ keys, uint32_t* __restrict__ order, int n) {
  extern __shared__ uint32_t hist[];   // kClassBins counters, then one partial sum per warp
  uint32_t* warpSum = hist + kClassBins;
  const int tid = threadIdx.x;
  for (int i = tid; i < kClassBins; i += kOrderThreads) hist[i] = 0;
  __syncthreads();
  for (int q = tid; q < n; q += kOrderThreads) atomicAdd(&hist[keys[q]], 1u);
  __syncthreads();
  // exclusive scan: every thread owns kClassBins / kOrderThreads consecutive bins
  constexpr int per = kClassBins / kOrderThreads;
  uint32_t local = 0;
  for (int i = 0; i < per; ++i) local += hist[tid * per + i];
  uint32_t incl = local;
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= d) incl += v;
  }
  if ((tid & 31) == 31) warpSum[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    const uint32_t w = warpSum[tid];
    uint32_t wi = w;
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
      if (tid >= d) wi += v;
    }
    warpSum[tid] = wi - w;   // exclusive
  }
  __syncthreads();
  uint32_t run = warpSum[tid >> 5] + incl - local;
  for (int i = 0; i < per; ++i) {
    const uint32_t c = hist[tid * per + i];
    hist[tid * per + i] = run;
    run += c;
  }
  __syncthreads();
  for (int q = tid; q < n; q += kOrderThreads) order[atomicAdd(&hist[keys[q]], 1u)] = (uint32_t) q;
}

template <class T>
__global__ void __launch_bounds__(kMicroThreads) k_micro(const __grid_constant__ MicroParams p) {
  extern __shared__ __align__(16) uint8_t msm[];
  MicroSearch<T> ms;
  ms.ws = msm + (size_t) threadIdx.x * T::WS_BYTES;
  ms.cfg = p.cfg;
  const int stride = gridDim.x * blockDim.x;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  bool running = false;
  int slot = -1;
  uint8_t* resOut = nullptr;
  while (true) {
    __syncwarp();
    // Lanes that start their regions together run in lock step as long as the regions behave alike (most regions of a
    // genome are alike), so idle lanes wait until a group of them can start together.
    const unsigned idle = __ballot_sync(0xffffffffu, !running);
    const bool take = __popc(idle) >= p.refill || idle == 0xffffffffu;
    if (take && !running && r < p.total) {
      MicroChunkDesc c = p.one;
      if (p.table != nullptr) {
        int lo = 0, hi = p.nChunks - 1;  // last chunk whose first region is <= r
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (__ldg(&p.table[mid].first) <= r) lo = mid; else hi = mid - 1;
        }
        c = p.table[lo];
      }
      int q = r - c.first;
      if (c.order != nullptr) q = (int) __ldg(c.order + q);
      const uint32_t off = __ldg(c.offsets + q);
      resOut = c.results + (size_t) q * T::RES_BYTES;
      if (off == kNoPacket) {
        // the device-side builder found that the region does not fit a packet: hand it back right away
        uint32_t* o = reinterpret_cast<uint32_t*>(resOut);
#pragma unroll
        for (int i = 0; i < T::RES_BYTES / 4; ++i) o[i] = i == 0 ? (uint32_t) kMicroFallback << (8 * MR_STATUS) : 0u;
        r += stride;
      } else {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(c.packets) + off;
        uint32_t* dst = reinterpret_cast<uint32_t*>(ms.ws + T::OFF_PKT);
        const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
        const int nw = (int) (w2 >> 24);
        dst[0] = w0; dst[1] = w1; dst[2] = w2;
        for (int i = 3; i < nw; ++i) dst[i] = __ldg(src + i);
        ms.begin();
        running = true;
      }
    }
    if (__all_sync(0xffffffffu, !running && r >= p.total)) break;  // nothing left for any lane
    if (running) {
      const int st = ms.iterate(slot);
      if (st != kMicroPending) {
        uint32_t tmp[T::RES_BYTES / 4];
        ms.writeResult(st, slot, reinterpret_cast<uint8_t*>(tmp));
        uint32_t* o = reinterpret_cast<uint32_t*>(resOut);
#pragma unroll
        for (int i = 0; i < T::RES_BYTES / 4; ++i) o[i] = tmp[i];
        running = false;
        r += stride;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backend
#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      char b_[512];                                                                          \
      snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      setErr(b_);                                                                            \
      return VCE_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct AlleleDev {
  DevBuf<int32_t> aStart, aEnd, aNtOff;
  DevBuf<uint8_t> aNull, nt;
  void release() { aStart.release(); aEnd.release(); aNtOff.release(); aNull.release(); nt.release(); }
};

struct SideDev {  // one side of one pass
  DevBuf<int32_t> vStart, vEnd, slot0, nSl, oCnt, oOff, oSlot;
  DevBuf<uint8_t> vFlags, oOrig, decision;
  DevBuf<int8_t> gt, oIds;
  DevBuf<double> weight;
  size_t nVar = 0, nRows = 0;
  void release() {
    vStart.release(); vEnd.release(); slot0.release(); nSl.release(); oCnt.release(); oOff.release(); oSlot.release();
    vFlags.release(); oOrig.release(); decision.release(); gt.release(); oIds.release(); weight.release();
  }
};

class CudaBackend : public Backend {
 public:
  explicit CudaBackend(int device) : device_(device) {}
  ~CudaBackend() override { destroy(); }

  int init() {
    CUDA_TRY(cudaSetDevice(device_));
    // the thread-per-region chunks feed the host pipeline and go first; the warp-per-region batches fill the gaps
    int prLow = 0, prHigh = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
    CUDA_TRY(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, prLow));
    CUDA_TRY(cudaEventCreate(&ev0_));
    CUDA_TRY(cudaEventCreate(&ev1_));
    CUDA_TRY(cudaEventCreate(&evT0_));
    CUDA_TRY(cudaEventCreate(&evT1_));
    CUDA_TRY(cudaDeviceGetAttribute(&smCount_, cudaDevAttrMultiProcessorCount, device_));
    int maxSmem = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device_));
    maxSmemOptin_ = maxSmem;
    CUDA_TRY(cudaFuncSetAttribute(k_search<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
    CUDA_TRY(cudaFuncSetAttribute(k_search<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (4096 + 1024) * kFastWarps));
    CUDA_TRY(cudaFuncSetAttribute(k_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
    if (const char* e = std::getenv("VCE_MID_KERNEL")) oldMid_ = std::string(e) != "cta";   // "cta": mid tier through the CTA kernel (slower there, measured)
    if (const char* e = std::getenv("VCE_CTA_BIG")) ctaBig_ = std::atoi(e) != 0;
    if (const char* e = std::getenv("VCE_CTA_TMA")) ctaTma_ = std::atoi(e) != 0;
    CUDA_TRY(cudaFuncSetAttribute(k_class_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (kClassBins + 32) * (int) sizeof(uint32_t)));
    if (const char* e = std::getenv("VCE_CLASS_ORDER")) classOrder_ = std::atoi(e) != 0;
    CUDA_TRY(cudaFuncSetAttribute(k_micro<MicroA>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMicroThreads * MicroA::WS_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(k_micro<MicroB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMicroThreads * MicroB::WS_BYTES));
    for (int i = 0; i < kMicroLanes; ++i) CUDA_TRY(cudaStreamCreateWithPriority(&mStream_[i], cudaStreamNonBlocking, prHigh));
    for (int i = 0; i < 2; ++i) { CUDA_TRY(cudaEventCreate(&evM0_[i])); CUDA_TRY(cudaEventCreate(&evM1_[i])); }
    {
      int perSm = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_micro<MicroA>, kMicroThreads, kMicroThreads * MicroA::WS_BYTES));
      microBlocksPerSm_[0] = perSm < 1 ? 1 : perSm;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_micro<MicroB>, kMicroThreads, kMicroThreads * MicroB::WS_BYTES));
      microBlocksPerSm_[1] = perSm < 1 ? 1 : perSm;
      if (const char* e = std::getenv("VCE_MICRO_REFILL")) microRefill_ = std::max(1, std::min(32, std::atoi(e)));
    }
    ready_ = true;
    return VCE_OK;
  }
  void destroy() {
    if (!ready_) return;
    cudaSetDevice(device_);
    // nothing of this backend may still be running when its buffers go (a failed call leaves work in flight)
    cudaStreamSynchronize(stream_);
    for (int i = 0; i < kMicroLanes; ++i) if (mStream_[i]) cudaStreamSynchronize(mStream_[i]);
    for (int s = 0; s < 2; ++s) { al_[s].release(); pdev_[0][s].release(); pdev_[1][s].release(); }
    sRegs_.release(); sSyncOff_.release(); sWindows_.release();
    regs_.release(); rres_.release(); syncOff_.release(); syncBuf_.release(); windows_.release(); warns_.release();
    warnScratch_.release(); ints_.release(); arena_.release(); scanTmp_.release();
    mTableC_[0].release(); mTableC_[1].release();
    cudaEventDestroy(ev0_);
    cudaEventDestroy(ev1_);
    cudaEventDestroy(evT0_);
    cudaEventDestroy(evT1_);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(evM0_[i]); cudaEventDestroy(evM1_[i]); }
    for (cudaEvent_t e : mEvents_) cudaEventDestroy(e);
    mEvents_.clear();
    for (int i = 0; i < kMicroLanes; ++i) cudaStreamDestroy(mStream_[i]);
    for (DevBlock& b : dBlocks_) cudaFree(b.p);
    dBlocks_.clear();
    cudaStreamDestroy(stream_);
    ready_ = false;
  }

  int setAlleles(int side, const AlleleTable& t) override {
    CUDA_TRY(cudaSetDevice(device_));
    AlleleDev& d = al_[side];
    CUDA_TRY(d.aStart.ensure(t.aStart.size() + 1));
    CUDA_TRY(d.aEnd.ensure(t.aEnd.size() + 1));
    CUDA_TRY(d.aNtOff.ensure(t.aNtOff.size() + 1));
    CUDA_TRY(d.aNull.ensure(t.aNull.size() + 1));
    CUDA_TRY(d.nt.ensure(t.nt.size() + 16));
    CUDA_TRY(cudaMemcpyAsync(d.aStart.p, t.aStart.data(), t.aStart.size() * 4, cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaMemcpyAsync(d.aEnd.p, t.aEnd.data(), t.aEnd.size() * 4, cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaMemcpyAsync(d.aNtOff.p, t.aNtOff.data(), t.aNtOff.size() * 4, cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaMemcpyAsync(d.aNull.p, t.aNull.data(), t.aNull.size(), cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaMemcpyAsync(d.nt.p, t.nt.data(), t.nt.size(), cudaMemcpyHostToDevice, stream_));
    stats.h2dBytes += (int64_t) (t.aStart.size() * 4 + t.aEnd.size() * 4 + t.aNtOff.size() * 4 + t.aNull.size() + t.nt.size());
    return VCE_OK;
  }

  int uploadPass(PassData& pd) override {
    CUDA_TRY(cudaSetDevice(device_));
    CUDA_TRY(ints_.ensure(16));
    for (int side = 0; side < 2; ++side) {
      HostSide& hs = pd.side[side];
      SideDev& d = pdev_[pd.pass][side];
      const size_t n = hs.n();
      d.nVar = n;
      CUDA_TRY(d.vStart.ensure(n + 1)); CUDA_TRY(d.vEnd.ensure(n + 1)); CUDA_TRY(d.slot0.ensure(n + 1));
      CUDA_TRY(d.nSl.ensure(n + 1)); CUDA_TRY(d.vFlags.ensure(n + 1)); CUDA_TRY(d.gt.ensure(4 * n + 4));
      CUDA_TRY(d.oCnt.ensure(n + 1)); CUDA_TRY(d.oOff.ensure(n + 1)); CUDA_TRY(d.decision.ensure(n + 1));
      CUDA_TRY(d.weight.ensure(n + 1));
      if (n) {
        CUDA_TRY(cudaMemcpyAsync(d.vStart.p, hs.vStart.data(), n * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_TRY(cudaMemcpyAsync(d.vEnd.p, hs.vEnd.data(), n * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_TRY(cudaMemcpyAsync(d.slot0.p, hs.slot0.data(), n * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_TRY(cudaMemcpyAsync(d.nSl.p, hs.nSl.data(), n * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_TRY(cudaMemcpyAsync(d.vFlags.p, hs.vFlags.data(), n, cudaMemcpyHostToDevice, stream_));
        CUDA_TRY(cudaMemcpyAsync(d.gt.p, hs.gt.data(), n * 4, cudaMemcpyHostToDevice, stream_));
        stats.h2dBytes += (int64_t) (n * 21);
      }
    }
    return resetPass(pd);
  }

  int resetPass(PassData& pd) override {
    CUDA_TRY(cudaSetDevice(device_));
    for (int side = 0; side < 2; ++side) {
      SideDev& d = pdev_[pd.pass][side];
      CUDA_TRY(cudaMemsetAsync(d.decision.p, 0, d.nVar + 1, stream_));
      CUDA_TRY(cudaMemsetAsync(d.weight.p, 0, (d.nVar + 1) * sizeof(double), stream_));
    }
    return VCE_OK;
  }

  int orient(PassData& pd) override {
    CUDA_TRY(cudaSetDevice(device_));
    curPass_ = pd.pass;
    // K1: count -> exclusive scan -> fill
    for (int side = 0; side < 2; ++side) {
      HostSide& hs = pd.side[side];
      SideDev& d = pdev_[pd.pass][side];
      const size_t n = d.nVar;
      OrientParams op{(int) n, pd.kind[side], pd.H, hs.gtPloidy, d.gt.p, d.vFlags.p, d.slot0.p, d.nSl.p, al_[side].aNull.p};
      const int tb = 256;
      const int nb = (int) ((n + 1 + tb - 1) / tb);
      CUDA_TRY(cudaMemsetAsync(ints_.p + side, 0, sizeof(int), stream_));
      k_orient_count<<<nb, tb, 0, stream_>>>(op, d.oCnt.p, ints_.p + side);
      ++stats.launches;
      size_t tmpBytes = 0;
      CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, d.oCnt.p, d.oOff.p, (int) (n + 1), stream_));
      CUDA_TRY(scanTmp_.ensure(tmpBytes + 16));
      CUDA_TRY(cub::DeviceScan::ExclusiveSum(scanTmp_.p, tmpBytes, d.oCnt.p, d.oOff.p, (int) (n + 1), stream_));
      ++stats.launches;
      // the row total is bounded on the host so that no round trip is needed before the fill
      const size_t rowBound = n * (size_t) std::max(1, pd.maxOrient[side]);
      CUDA_TRY(d.oSlot.ensure(rowBound * pd.H + 4));
      CUDA_TRY(d.oIds.ensure(rowBound * 4 + 4));
      CUDA_TRY(d.oOrig.ensure(rowBound + 4));
      if (n) {
        k_orient_fill<<<(int) ((n + tb - 1) / tb), tb, 0, stream_>>>(op, d.oOff.p, d.oSlot.p, d.oIds.p, d.oOrig.p);
        ++stats.launches;
      }
      CUDA_TRY(cudaGetLastError());
    }
    return VCE_OK;
  }

  int stage(const std::vector<RegionDesc>& regs, const std::vector<uint8_t>& windows, const std::vector<int32_t>& syncOff,
            size_t syncTotal) override {
    CUDA_TRY(cudaSetDevice(device_));
    const size_t nr = regs.size();
    CUDA_TRY(sRegs_.ensure(nr + 1)); CUDA_TRY(sSyncOff_.ensure(nr + 1)); CUDA_TRY(sWindows_.ensure(windows.size() + 64));
    CUDA_TRY(ensureSync(syncTotal));
    CUDA_TRY(cudaMemcpyAsync(sRegs_.p, regs.data(), nr * sizeof(RegionDesc), cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaMemcpyAsync(sSyncOff_.p, syncOff.data(), nr * 4, cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaMemcpyAsync(sWindows_.p, windows.data(), windows.size(), cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaStreamSynchronize(stream_));
    stats.h2dBytes += (int64_t) (nr * (sizeof(RegionDesc) + 4) + windows.size());
    return VCE_OK;
  }

  // HBM for the search arenas of one launch: most of what is free (a B200 has 180 GB and this engine's inputs are small), so
  // that a batch of huge regions keeps every SM full of resident regions
  long long arenaBudgetBytes() {
    size_t freeB = 0, totalB = 0;
    long long budget = 24ll << 30;
    if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
      const long long avail = (long long) freeB + (long long) arena_.cap * 4;
      budget = std::max<long long>(4ll << 30, std::min<long long>(96ll << 30, avail * 6 / 10));
    }
    if (const char* e = std::getenv("VCE_ARENA_GB")) budget = std::max(1, std::atoi(e)) * (1ll << 30);
    return budget;
  }

  cudaError_t ensureSync(size_t total) {
    // the device sync buffer mirrors the host one (which grows when regions are merged)
    if (total > syncBuf_.cap) {
      DevBuf<int32_t> nb;
      cudaError_t e = nb.ensure(total * 2 + 1024);
      if (e != cudaSuccess) return e;
      if (syncBuf_.p && syncBufUsed_) {
        e = cudaMemcpyAsync(nb.p, syncBuf_.p, syncBufUsed_ * 4, cudaMemcpyDeviceToDevice, stream_);
        if (e != cudaSuccess) return e;
      }
      e = cudaStreamSynchronize(stream_);
      if (e != cudaSuccess) return e;
      syncBuf_.release();
      syncBuf_ = nb;
    }
    syncBufUsed_ = total;
    return cudaSuccess;
  }

  SideArrays sideArrays(int side) const {
    const SideDev& d = pdev_[curPass_][side];
    const AlleleDev& a = al_[side];
    SideArrays S;
    S.vStart = d.vStart.p; S.vEnd = d.vEnd.p; S.oOff = d.oOff.p; S.oSlot = d.oSlot.p; S.oIds = d.oIds.p; S.oOrig = d.oOrig.p;
    S.aStart = a.aStart.p; S.aEnd = a.aEnd.p; S.aNtOff = a.aNtOff.p; S.nt = a.nt.p; S.vFlags = d.vFlags.p;
    S.outDecision = d.decision.p; S.outWeight = d.weight.p;
    return S;
  }

  int search(PassData& pd, const SearchConfig& sc, const std::vector<RegionDesc>& regs, const std::vector<uint8_t>& windows,
             int tier, bool /*staged*/, std::vector<RegionResult>& out, const std::vector<int32_t>& syncOff, size_t syncTotal,
             std::vector<WarnRec>& warns) override {
    int tierFirst[kTierGeneral + 2];
    for (int t = 0; t <= kTierGeneral + 1; ++t) tierFirst[t] = t <= tier ? 0 : (int) regs.size();
    const int rc = searchStart(pd, sc, regs, windows, tierFirst, syncOff, syncTotal);
    if (rc) return rc;
    return searchFinish(out, warns);
  }

  int searchStart(PassData& pd, const SearchConfig& sc, const std::vector<RegionDesc>& regs, const std::vector<uint8_t>& windows,
                  const int* tierFirst, const std::vector<int32_t>& syncOff, size_t syncTotal) override {
    CUDA_TRY(cudaSetDevice(device_));
    curPass_ = pd.pass;
    const size_t nr = regs.size();
    inFlight_ = nr;
    inFlightLaunches_ = 0;
    if (nr == 0) return VCE_OK;
    CUDA_TRY(rres_.ensure(nr));
    CUDA_TRY(ints_.ensure(16));
    CUDA_TRY(ensureSync(syncTotal));
    const int warnCap = 1 << 16;
    CUDA_TRY(warns_.ensure(warnCap));
    CUDA_TRY(regs_.ensure(nr)); CUDA_TRY(syncOff_.ensure(nr)); CUDA_TRY(windows_.ensure(windows.size() + 64));
    CUDA_TRY(cudaMemcpyAsync(regs_.p, regs.data(), nr * sizeof(RegionDesc), cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaMemcpyAsync(syncOff_.p, syncOff.data(), nr * 4, cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaMemcpyAsync(windows_.p, windows.data(), windows.size(), cudaMemcpyHostToDevice, stream_));
    stats.h2dBytes += (int64_t) (nr * (sizeof(RegionDesc) + 4) + windows.size());
    static_assert(kTierGeneral + 1 <= 8, "work counters live in ints_[4..11]");
    CUDA_TRY(cudaMemsetAsync(ints_.p + 4, 0, 9 * sizeof(int), stream_));  // [4..11] work counters per tier, [12] warning count

    // arena / scratch sizes over all tiers first (buffers must not move between the launches)
    int blocksT[kTierGeneral + 1] = {0}, threadsT[kTierGeneral + 1] = {0};
    size_t smemT[kTierGeneral + 1] = {0};
    SearchParams spT[kTierGeneral + 1];
    CtaParams cpT[kTierGeneral + 1];
    bool isCta[kTierGeneral + 1] = {false};
    size_t arenaWords = 0;
    int maxWarps = 0;
    for (int tier = 0; tier <= kTierGeneral; ++tier) {
      const int r0 = tierFirst[tier], r1 = tierFirst[tier + 1];
      if (r1 <= r0) continue;
      const bool general = tier >= kTierMid;
      const size_t nt = (size_t) (r1 - r0);
      SearchParams& sp = spT[tier];
      std::memset(&sp, 0, sizeof(sp));
      // The CTA kernel (vce_cta.h) is the LATENCY tier: one warp per region with ~100 KB of shared memory, two regions per SM.
      // It takes the full-budget tier while the batch fits one wave of CTAs; a larger batch of huge regions is a THROUGHPUT
      // problem and runs in the warp kernel with up to 24 resident regions per SM (measured, profiles/r02_*: per iteration the
      // two kernels cost about the same, ~13 us, so concurrency decides).
      isCta[tier] = (tier == kTierMid && !oldMid_) || (tier == kTierCta && ctaBig_ && (long long) nt <= (long long) smCount_ * 2);
      sp.regs = regs_.p + r0; sp.nRegs = (int) nt;
      sp.S[0] = sideArrays(0); sp.S[1] = sideArrays(1);
      sp.windows = windows_.p; sp.cfg = sc; sp.out = rres_.p + r0; sp.syncOff = syncOff_.p + r0; sp.syncOut = syncBuf_.p;
      sp.warns = warns_.p; sp.warnCount = ints_.p + 12; sp.warnCap = warnCap; sp.counter = ints_.p + 4 + tier;
      sp.H = pd.H;
      sp.maxChildren = 1 + std::max(pd.maxOrient[0], pd.maxOrient[1]);
      sp.regionBase = r0;
      int blocks = 0, threads = kFastWarps * 32;
      size_t smemBytes = 0;
      if (!general) {
        const FastShape fs = fastShape(tier);
        sp.fs = fs;
        PathLayout LM;
        LM.init(2, fs.q, fs.kv, fs.smax, 4);
        const int preWords = (fs.slots() + 2) * LM.W + fs.ordCap() + fs.slots();
        const int ctlWords = ((preWords + 16 + 2 * fs.maxChildren + 3) & ~3) - preWords;
        sp.winSmemBytes = 256;
        sp.tabBytes = fs.tabBytes;
        const int kcWords = (int) ((sizeof(RegionConst) + 15) / 16) * 4;
        sp.wordsPerWarp = kcWords + preWords + ctlWords + (sp.tabBytes + sp.winSmemBytes) / 4;
        const size_t bytesPerWarp = (size_t) sp.wordsPerWarp * 4;
        int warpsPerBlock = (int) std::min<size_t>(kFastWarps, (size_t) maxSmemOptin_ / bytesPerWarp);
        if (warpsPerBlock < 1) { setErr("fast kernel shared memory exceeds the device limit"); return VCE_ERR_CUDA; }
        // prefer several smaller blocks per SM over one large one
        if (warpsPerBlock < kFastWarps) warpsPerBlock = std::max(1, (int) ((size_t) maxSmemOptin_ / bytesPerWarp) / 2);
        threads = warpsPerBlock * 32;
        smemBytes = bytesPerWarp * warpsPerBlock;
        // moderate batches: one region per warp, short-lived blocks, so that the block scheduler can favour the
        // high-priority streams; huge batches: resident blocks pulling from the work queue
        const int need = (int) ((nt + warpsPerBlock - 1) / warpsPerBlock);
        blocks = need;
        if (need > 16384) {
          int perSm = 0;
          CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_search<false>, threads, smemBytes));
          if (perSm < 1) perSm = 1;
          blocks = std::min(need, smCount_ * perSm);
        }
      } else if (isCta[tier]) {
        // one region per CTA (vce_cta.h): frontier directory + block cache + work records in shared memory, blocks and
        // path records in a per-CTA arena
        CtaParams& cp = cpT[tier];
        cp.shape = ctaShape(tier);
        cp.tma = ctaTma_ ? 1 : 0;
        const int kcWords = (int) ((sizeof(RegionConst) + 15) / 16) * 4;
        smemBytes = ((size_t) kcWords + (size_t) ctaSmemWords(cp.shape) + 4) * 4;
        threads = 32;
        int perSm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_cta, threads, smemBytes));
        if (perSm < 1) { setErr("CTA kernel shared memory exceeds the device limit"); return VCE_ERR_CUDA; }
        const long long mc = sp.maxChildren;
        const long long pool = std::min<long long>(cp.shape.capacity(), (long long) sc.maxPaths + 1 + mc) + 64;
        long long words = (ctaArenaWords(cp.shape, cta::kMaxW, pool) + 3) & ~3LL;
        const long long budgetWords = arenaBudgetBytes() / 4;
        long long ctasWanted = std::min<long long>((long long) nt, (long long) smCount_ * perSm);
        ctasWanted = std::min(ctasWanted, std::max<long long>(1, budgetWords / words));
        blocks = (int) ctasWanted;
        arenaWords = std::max(arenaWords, (size_t) blocks * (size_t) words);
        cp.arenaWordsPerCta = words;
      } else {
        // per-warp arena sized for the largest record layout in this launch at the full search budget
        long long maxW = 0;
        for (int q = r0; q < r1; ++q) maxW = std::max(maxW, (long long) regs[q].layoutW);
        const long long mc = sp.maxChildren;
        long long capFull = (long long) sc.maxPaths + 1 + mc;
        if (tier == kTierMid) capFull = std::min<long long>(capFull, kMidCap);   // (only with VCE_MID_KERNEL=old)
        // resident blocks per SM pulling regions from the work queue (registers allow 3 blocks of 8 warps; a 4th
        // queued block evens out the tail)
        static const int bps = std::getenv("VCE_GENERAL_BPS") ? std::max(1, std::atoi(std::getenv("VCE_GENERAL_BPS"))) : 4;
        long long words = (capFull + 1 + mc + 2) * maxW + (3 * capFull + 1024) + (capFull + 1 + mc) + 16 + 2 * mc + 64;
        const long long budgetWords = arenaBudgetBytes() / 4;
        long long warpsWanted = std::min<long long>((long long) nt, (long long) smCount_ * kFastWarps * bps);
        words = (words + 3) & ~3LL;
        if (words > budgetWords) words = budgetWords;  // a single region larger than the budget runs with reduced capacity
        long long warpsFit = std::max<long long>(1, budgetWords / words);
        long long nw = std::min(warpsWanted, warpsFit);
        blocks = (int) ((nw + kFastWarps - 1) / kFastWarps);
        words &= ~3LL;   // records stay 16-byte aligned
        arenaWords = std::max(arenaWords, (size_t) blocks * kFastWarps * (size_t) words);
        sp.arenaWordsPerWarp = words;
        sp.tabBytes = 4096;
        const int kcWords = (int) ((sizeof(RegionConst) + 15) / 16) * 4;
        smemBytes = ((size_t) sp.tabBytes + (size_t) kcWords * 4) * kFastWarps;
      }
      blocksT[tier] = blocks; threadsT[tier] = threads; smemT[tier] = smemBytes;
      maxWarps = std::max(maxWarps, blocks * std::max(1, threads / 32));
    }
    CUDA_TRY(warnScratch_.ensure((size_t) std::max(1, maxWarps) * kWarnPerRegion));
    if (arenaWords > 0) CUDA_TRY(arena_.ensure(arenaWords));   // shared by the launches below (they run in stream order)
    CUDA_TRY(cudaEventRecord(ev0_, stream_));
    tlMark(stream_, "wide batch: inputs resident, kernels next; regions", (int) nr);
    for (int tier = 0; tier <= kTierGeneral; ++tier) {
      if (blocksT[tier] == 0) continue;
      spT[tier].warnScratch = warnScratch_.p;
      spT[tier].arena = arena_.p;
      if (isCta[tier]) {
        cpT[tier].sp = spT[tier];
        k_cta<<<blocksT[tier], threadsT[tier], smemT[tier], stream_>>>(cpT[tier]);
      } else if (tier >= kTierMid) {
        k_search<true><<<blocksT[tier], threadsT[tier], smemT[tier], stream_>>>(spT[tier]);
      } else {
        k_search<false><<<blocksT[tier], threadsT[tier], smemT[tier], stream_>>>(spT[tier]);
      }
      CUDA_TRY(cudaGetLastError());
      ++stats.launches;
      ++inFlightLaunches_;
      if (std::getenv("VCE_TIMING")) {
        fprintf(stderr, "[vce]     k_search tier %d: %d regions, %d blocks x %d threads, %zu B shared memory per block\n", tier,
                tierFirst[tier + 1] - tierFirst[tier], blocksT[tier], threadsT[tier], smemT[tier]);
      }
    }
    CUDA_TRY(cudaEventRecord(ev1_, stream_));
    tlMark(stream_, "wide batch: kernels done; regions", (int) nr);
    return VCE_OK;
  }

  int searchFinish(std::vector<RegionResult>& out, std::vector<WarnRec>& warns) override {
    CUDA_TRY(cudaSetDevice(device_));
    const size_t nr = inFlight_;
    out.resize(nr);
    if (nr == 0) return VCE_OK;
    int nWarn = 0;
    const int warnCap = 1 << 16;
    CUDA_TRY(cudaMemcpyAsync(out.data(), rres_.p, nr * sizeof(RegionResult), cudaMemcpyDeviceToHost, stream_));
    CUDA_TRY(cudaMemcpyAsync(&nWarn, ints_.p + 12, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CUDA_TRY(cudaStreamSynchronize(stream_));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, ev0_, ev1_));
    stats.searchMs += ms;
    stats.d2hBytes += (int64_t) (nr * sizeof(RegionResult));
    if (nWarn > 0) {
      nWarn = std::min(nWarn, warnCap);
      const size_t at = warns.size();
      warns.resize(at + nWarn);
      CUDA_TRY(cudaMemcpy(warns.data() + at, warns_.p, (size_t) nWarn * sizeof(WarnRec), cudaMemcpyDeviceToHost));
      // keep each region's warnings in emission order
      std::stable_sort(warns.begin() + at, warns.end(), [](const WarnRec& a, const WarnRec& b) { return a.region < b.region; });
    }
    inFlight_ = 0;
    return VCE_OK;
  }
  size_t inFlight_ = 0;
  int inFlightLaunches_ = 0;

  int fetchPass(PassData& pd, std::vector<int32_t>& syncBuf) override {
    CUDA_TRY(cudaSetDevice(device_));
    for (int side = 0; side < 2; ++side) {
      HostSide& hs = pd.side[side];
      SideDev& d = pdev_[pd.pass][side];
      const size_t n = hs.n();
      int32_t rows = 0;
      CUDA_TRY(cudaMemcpyAsync(&rows, d.oOff.p + n, 4, cudaMemcpyDeviceToHost, stream_));
      CUDA_TRY(cudaStreamSynchronize(stream_));
      d.nRows = (size_t) rows;
      hs.decision.resize(n);
      hs.weight.resize(n);
      if (n) {
        CUDA_TRY(cudaMemcpyAsync(hs.decision.data(), d.decision.p, n, cudaMemcpyDeviceToHost, stream_));
        CUDA_TRY(cudaMemcpyAsync(hs.weight.data(), d.weight.p, n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
      }
      if (!pd.wantOrientTables) continue;
      hs.oOff.resize(n + 1);
      hs.oIds.resize(d.nRows * 4 + 4);
      hs.oOrig.resize(d.nRows + 4);
      CUDA_TRY(cudaMemcpyAsync(hs.oOff.data(), d.oOff.p, (n + 1) * 4, cudaMemcpyDeviceToHost, stream_));
      if (d.nRows) {
        CUDA_TRY(cudaMemcpyAsync(hs.oIds.data(), d.oIds.p, d.nRows * 4, cudaMemcpyDeviceToHost, stream_));
        CUDA_TRY(cudaMemcpyAsync(hs.oOrig.data(), d.oOrig.p, d.nRows, cudaMemcpyDeviceToHost, stream_));
      }
    }
    if (!syncBuf.empty()) CUDA_TRY(cudaMemcpyAsync(syncBuf.data(), syncBuf_.p, syncBuf.size() * 4, cudaMemcpyDeviceToHost, stream_));
    CUDA_TRY(cudaStreamSynchronize(stream_));
    for (int side = 0; side < 2; ++side) {
      const size_t n = pd.side[side].n();
      stats.d2hBytes += (int64_t) (n * 9 + (n + 1) * 4 + pdev_[pd.pass][side].nRows * 5);
    }
    stats.d2hBytes += (int64_t) syncBuf.size() * 4;
    return VCE_OK;
  }


  // ---------------------------------------------------------------- thread-per-region path
  void* hostAlloc(size_t bytes) override {
    cudaSetDevice(device_);
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { setErr("cudaHostAlloc failed"); return nullptr; }
    return p;
  }
  void hostFree(void* p) override {
    cudaSetDevice(device_);
    cudaFreeHost(p);
  }
  int microReset() override {
    std::lock_guard<std::mutex> lk(mMu_);
    for (DevBlock& b : dBlocks_) b.used = 0;
    mNextEvent_ = 0;
    mNextLane_ = 0;
    pendingB_.clear();
    lastMicroMs_ = 0;
    lastMicroRegions_ = 0;
    return VCE_OK;
  }
  void microMark() override {
    std::lock_guard<std::mutex> lk(mMu_);
    dMarks_.clear();
    for (DevBlock& b : dBlocks_) dMarks_.push_back(b.used);
    mMarkEvent_ = mNextEvent_;
  }
  void microRewind() override {
    std::lock_guard<std::mutex> lk(mMu_);
    for (size_t i = 0; i < dBlocks_.size(); ++i) dBlocks_[i].used = i < dMarks_.size() ? dMarks_[i] : 0;
    mNextEvent_ = mMarkEvent_;
  }
  void* devAlloc(size_t bytes) {  // caller holds mMu_
    bytes = (bytes + 255) & ~(size_t) 255;
    for (DevBlock& b : dBlocks_) {
      if (b.cap - b.used >= bytes) {
        void* p = b.p + b.used;
        b.used += bytes;
        return p;
      }
    }
    const size_t cap = std::max(bytes, (size_t) 256 << 20);
    void* p = nullptr;
    if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr;
    dBlocks_.push_back(DevBlock{static_cast<uint8_t*>(p), cap, bytes});
    return p;
  }
  int microUpload(MicroJob& j) override {
    CUDA_TRY(cudaSetDevice(device_));
    {
      std::lock_guard<std::mutex> lk(mMu_);
      j.dPackets = devAlloc(j.packetBytes + 16);
      j.dOffsets = devAlloc((size_t) j.n * 4);
      j.dResults = devAlloc((size_t) j.n * j.resBytes);
      j.lane = mNextLane_;
      mNextLane_ = (mNextLane_ + 1) % kMicroLanes;
      if ((int) mEvents_.size() <= mNextEvent_) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { setErr("cudaEventCreate failed"); return VCE_ERR_CUDA; }
        mEvents_.push_back(e);
      }
      j.event = mNextEvent_++;
      j.evHandle = mEvents_[j.event];
      stats.h2dBytes += (int64_t) (j.packetBytes + (size_t) j.n * 4);
    }
    if (!j.dPackets || !j.dOffsets || !j.dResults) { setErr("out of device memory for region packets"); return VCE_ERR_CUDA; }
    cudaStream_t st = mStream_[j.lane];
    uploadsPending_.store(true);
    CUDA_TRY(cudaMemcpyAsync(j.dPackets, j.packets, j.packetBytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(j.dOffsets, j.offsets, (size_t) j.n * 4, cudaMemcpyHostToDevice, st));
    return VCE_OK;
  }
  int microGrid(int n, int cls) const {
    const int need = (n + kMicroThreads - 1) / kMicroThreads;
    return std::max(1, std::min(need, smCount_ * microBlocksPerSm_[cls]));
  }
  void microLaunch(const MicroParams& mp, int cls, cudaStream_t st) {
    if (cls == 0) k_micro<MicroA><<<microGrid(mp.total, 0), kMicroThreads, kMicroThreads * MicroA::WS_BYTES, st>>>(mp);
    else k_micro<MicroB><<<microGrid(mp.total, 1), kMicroThreads, kMicroThreads * MicroB::WS_BYTES, st>>>(mp);
  }
  int microRun(MicroJob& j, const MicroConfig& mc) override {
    CUDA_TRY(cudaSetDevice(device_));
    cudaStream_t st = mStream_[j.lane];
    MicroParams mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.one = MicroChunkDesc{static_cast<const uint8_t*>(j.dPackets), static_cast<const uint32_t*>(j.dOffsets),
                            static_cast<uint8_t*>(j.dResults), j.n, 0, nullptr};
    mp.table = nullptr;
    mp.nChunks = 1;
    mp.total = j.n;
    mp.refill = microRefill_;
    mp.cfg = mc;
    microLaunch(mp, j.cls, st);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(j.results, j.dResults, (size_t) j.n * j.resBytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(j.evHandle), st));
    {
      std::lock_guard<std::mutex> lk(mMu_);
      ++stats.launches;
      stats.d2hBytes += (int64_t) j.n * j.resBytes;
    }
    return VCE_OK;
  }
  int microRunAll(std::vector<MicroJob*>& jobs, const MicroConfig& mc) override {
    CUDA_TRY(cudaSetDevice(device_));
    if (jobs.empty()) return VCE_OK;
    // every upload has to be complete before the one launch that reads them all
    if (uploadsPending_.exchange(false)) {
      for (int i = 0; i < kMicroLanes; ++i) CUDA_TRY(cudaStreamSynchronize(mStream_[i]));
    }
    std::vector<MicroChunkDesc> table(jobs.size());
    int total = 0;
    for (size_t i = 0; i < jobs.size(); ++i) {
      MicroJob& j = *jobs[i];
      table[i] = MicroChunkDesc{static_cast<const uint8_t*>(j.dPackets), static_cast<const uint32_t*>(j.dOffsets),
                                static_cast<uint8_t*>(j.dResults), j.n, total, nullptr};
      total += j.n;
    }
    const int pc = jobs[0]->cls & 1;
    cudaStream_t st = mStream_[pc];   // the two packet classes run side by side
    CUDA_TRY(mTableC_[pc].ensure(table.size()));
    CUDA_TRY(cudaMemcpyAsync(mTableC_[pc].p, table.data(), table.size() * sizeof(MicroChunkDesc), cudaMemcpyHostToDevice, st));
    MicroParams mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.one = table[0];
    mp.table = mTableC_[pc].p;
    mp.nChunks = (int) table.size();
    mp.total = total;
    mp.refill = microRefill_;
    mp.cfg = mc;
    if (microPending_[pc]) microCollect();   // the event pair of this class is still waiting to be read
    CUDA_TRY(cudaEventRecord(evM0_[pc], st));
    microLaunch(mp, jobs[0]->cls, st);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(evM1_[pc], st));
    microPending_[pc] = true;
    microPendingTotal_[pc] = total;
    for (MicroJob* j : jobs) {
      CUDA_TRY(cudaMemcpyAsync(j->results, j->dResults, (size_t) j->n * j->resBytes, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(j->evHandle), st));
      j->lane = pc;
      stats.d2hBytes += (int64_t) j->n * j->resBytes;
    }
    ++stats.launches;
    return VCE_OK;   // no wait here: the caller decodes job by job (microWait) and reads the device times afterwards
  }
  // Device times of the launches started by microRunAll (blocks until they have finished).
  void microCollect() override {
    cudaSetDevice(device_);
    for (int pc = 0; pc < 2; ++pc) {
      if (!microPending_[pc]) continue;
      microPending_[pc] = false;
      float ms = 0;
      if (cudaEventSynchronize(evM1_[pc]) != cudaSuccess || cudaEventElapsedTime(&ms, evM0_[pc], evM1_[pc]) != cudaSuccess) continue;
      stats.searchMs += ms;
      if (pc == 0 && microPendingTotal_[pc] >= lastMicroRegions_ / 2) {  // the main MicroA launch (not the small settle retries)
        lastMicroMs_ = ms;
        lastMicroRegions_ = microPendingTotal_[pc];
      }
    }
  }
  bool microPending_[2] = {false, false};
  long long microPendingTotal_[2] = {0, 0};
  // Device-side digest of one chunk: slab up, k_build per packet class, the search kernels, results down.
  int microBuildRun(MicroBuildJob& b, MicroJob& ja, MicroJob& jb, const MicroConfig& mc) override {
    CUDA_TRY(cudaSetDevice(device_));
    MicroJob* jobs[2] = {&ja, &jb};
    const int strides[2] = {(MicroA::PKT_MAX + 15) & ~15, (MicroB::PKT_MAX + 15) & ~15};
    void* dSlab = nullptr;
    uint32_t* dKeys = nullptr;
    uint32_t* dOrder = nullptr;
    int lane = 0;
    {
      std::lock_guard<std::mutex> lk(mMu_);
      dSlab = devAlloc(b.slabBytes + 16);
      lane = mNextLane_;
      mNextLane_ = (mNextLane_ + 1) % kMicroLanes;
      for (int c = 0; c < 2; ++c) {
        MicroJob& j = *jobs[c];
        if (j.n == 0) continue;
        j.dPackets = devAlloc((size_t) j.n * strides[c] + 16);
        j.dOffsets = devAlloc((size_t) j.n * 4);
        j.dResults = devAlloc((size_t) j.n * j.resBytes);
        j.lane = lane;
        if (c == 0 && j.n > kOrderMin && classOrder_) {
          dKeys = static_cast<uint32_t*>(devAlloc((size_t) j.n * 4));
          dOrder = static_cast<uint32_t*>(devAlloc((size_t) j.n * 4));
          if (!dKeys || !dOrder) { setErr("out of device memory for region packets"); return VCE_ERR_CUDA; }
        }
        if ((int) mEvents_.size() <= mNextEvent_) {
          cudaEvent_t e;
          if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { setErr("cudaEventCreate failed"); return VCE_ERR_CUDA; }
          mEvents_.push_back(e);
        }
        j.event = mNextEvent_++;
        j.evHandle = mEvents_[j.event];
        if (!j.dPackets || !j.dOffsets || !j.dResults) { setErr("out of device memory for region packets"); return VCE_ERR_CUDA; }
        if (!b.scatter) stats.d2hBytes += (int64_t) j.n * j.resBytes;
        // k_build, and for MicroA the search launch (+ the class ordering); the MicroB search is counted by the flush
        stats.launches += c == 0 ? (dKeys != nullptr ? 3 : 2) : 1;
      }
      stats.h2dBytes += (int64_t) b.slabBytes;
    }
    if (!dSlab) { setErr("out of device memory for region packets"); return VCE_ERR_CUDA; }
    cudaStream_t st = mStream_[lane];
    CUDA_TRY(cudaMemcpyAsync(dSlab, b.slab, b.slabBytes, cudaMemcpyHostToDevice, st));
    const uint8_t* base = static_cast<const uint8_t*>(dSlab);
    if (b.resident) {   // the sequence's arrays were uploaded whole: the builder waits for those copies
      for (int e = 0; e < 2; ++e) {
        if (b.waitEvents[e] < 0) continue;
        cudaEvent_t ev = nullptr;
        {
          std::lock_guard<std::mutex> lk(mMu_);
          ev = mEvents_[b.waitEvents[e]];
        }
        CUDA_TRY(cudaStreamWaitEvent(st, ev, 0));
      }
    }
    tlMark(st, "chunk: slab + sequence arrays resident, k_build next; lane", lane);
    int q0 = 0;
    for (int c = 0; c < 2; ++c) {
      MicroJob& j = *jobs[c];
      if (j.n > 0) {
        BuildParams bp;
        bp.pv = b.resident ? b.rv : microBuildView(b, base);
        bp.regs = reinterpret_cast<const RegRange*>(base + b.regions) + q0;
        bp.winOff = b.resident ? nullptr : reinterpret_cast<const uint32_t*>(base + b.winOff) + q0;
        bp.windows = b.resident ? nullptr : base + b.windows;
        bp.t2 = b.resident ? b.t2 : nullptr;
        bp.nmask = b.resident ? b.nmask : nullptr;
        bp.packets = static_cast<uint8_t*>(j.dPackets);
        bp.offsets = static_cast<uint32_t*>(j.dOffsets);
        bp.keys = c == 0 ? dKeys : nullptr;
        bp.n = j.n;
        bp.stride = strides[c];
        const int blocks = (j.n + 127) / 128;
        if (c == 0) k_build<MicroA><<<blocks, 128, 0, st>>>(bp);
        else k_build<MicroB><<<blocks, 128, 0, st>>>(bp);
        CUDA_TRY(cudaGetLastError());
        if (c == 0 && dKeys != nullptr) {
          k_class_order<<<1, kOrderThreads, (kClassBins + 32) * sizeof(uint32_t), st>>>(dKeys, dOrder, j.n);
          CUDA_TRY(cudaGetLastError());
        }
        MicroParams mp;
        std::memset(&mp, 0, sizeof(mp));
        mp.one = MicroChunkDesc{static_cast<const uint8_t*>(j.dPackets), static_cast<const uint32_t*>(j.dOffsets),
                                static_cast<uint8_t*>(j.dResults), j.n, 0, nullptr};
        if (c == 0) mp.one.order = dOrder;
        mp.table = nullptr;
        mp.nChunks = 1;
        mp.total = j.n;
        mp.refill = microRefill_;
        mp.cfg = mc;
        if (c == 1) {
          // the few MicroB-shaped regions of a chunk would be a tiny, latency-bound launch of their own: their packets are
          // built now, the search runs once over all chunks (microBuildFlush)
          cudaEvent_t built = nullptr;
          {
            std::lock_guard<std::mutex> lk(mMu_);
            if ((int) mEvents_.size() <= mNextEvent_) {
              cudaEvent_t e;
              if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { setErr("cudaEventCreate failed"); return VCE_ERR_CUDA; }
              mEvents_.push_back(e);
            }
            built = mEvents_[mNextEvent_++];
            pendingB_.push_back(PendingB{&j, built, bp.regs, b.scatter ? reinterpret_cast<const int32_t*>(base + b.regIdx) + q0 : nullptr, b.scatter});
          }
          CUDA_TRY(cudaEventRecord(built, st));
          continue;
        }
        tlMark(st, "chunk: packets built, k_micro<MicroA> next; lane", lane);
        microLaunch(mp, c, st);
        CUDA_TRY(cudaGetLastError());
        tlMark(st, "chunk: k_micro<MicroA> done; regions", j.n);
        if (b.scatter) scatterLaunch(j, bp.regs, reinterpret_cast<const int32_t*>(base + b.regIdx) + q0, st);
        else CUDA_TRY(cudaMemcpyAsync(j.results, j.dResults, (size_t) j.n * j.resBytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(j.evHandle), st));
      }
      q0 += c == 0 ? b.nA : 0;
    }
    return VCE_OK;
  }
  // ---------------------------------------------------------------- device-resident results (vce_final.h)
  template <class T>
  struct PinBuf {   // page-locked host buffer, grows only
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
      if (n <= cap) return cudaSuccess;
      if (p) cudaFreeHost(p);
      p = nullptr;
      cap = 0;
      const size_t want = n + n / 8 + 64;
      cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), want * sizeof(T), cudaHostAllocDefault);
      if (e == cudaSuccess) cap = want;
      return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  };
  bool resultsSupported() const override { return true; }
  int resultsBegin(const ResultsSetup& s) override {
    CUDA_TRY(cudaSetDevice(device_));
    rs_ = s;
    const size_t nC = (size_t) s.nVar[0], nB = (size_t) s.nVar[1], nR = (size_t) s.nReg, nS = s.seq.size();
    CUDA_TRY(rDec_[0].ensure(nC + 16)); CUDA_TRY(rDec_[1].ensure(nB + 16)); CUDA_TRY(rWeight_.ensure(nC + 2));
    CUDA_TRY(rInfo_.ensure(nR + 2)); CUDA_TRY(rCnt_.ensure(nR + 2)); CUDA_TRY(rBase_.ensure(nR + 2)); CUDA_TRY(rRel_.ensure((nR + 2) * kFinalRel));
    CUDA_TRY(rSeq_.ensure(nS + 1)); CUDA_TRY(rIter_.ensure(2));
    std::vector<FinalSeq>& fs = fsHost_;
    fs.assign(nS, FinalSeq());
    for (size_t k = 0; k < nS; ++k) {
      for (int side = 0; side < 2; ++side) {
        fs[k].vStart[side] = s.seq[k].vStart[side];
        fs[k].phased[side] = s.seq[k].phased[side];
        fs[k].statusIn[side] = s.seq[k].statusIn[side];
        fs[k].first[side] = s.seq[k].first[side];
        fs[k].n[side] = s.seq[k].n[side];
      }
      fs[k].regFirst = s.seq[k].regFirst;
      fs[k].nReg = s.seq[k].nReg;
    }
    cudaStream_t st = mStream_[0];
    CUDA_TRY(cudaMemsetAsync(rInfo_.p, 0, (nR + 2) * sizeof(uint32_t), st));
    CUDA_TRY(cudaMemsetAsync(rIter_.p, 0, 2 * sizeof(unsigned long long), st));
    if (nS) CUDA_TRY(cudaMemcpyAsync(rSeq_.p, fs.data(), nS * sizeof(FinalSeq), cudaMemcpyHostToDevice, st));
    // the chunks' scatter kernels run on every lane: they wait for this event, the host does not
    if (!rBeginEv_) CUDA_TRY(cudaEventCreateWithFlags(&rBeginEv_, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(rBeginEv_, st));
    stats.h2dBytes += (int64_t) (nS * sizeof(FinalSeq));
    return VCE_OK;
  }
  void scatterLaunch(const MicroJob& j, const RegRange* regs, const int32_t* regIdx, cudaStream_t st) {
    cudaStreamWaitEvent(st, rBeginEv_, 0);
    ScatterArgs a;
    a.results = static_cast<const uint8_t*>(j.dResults);
    a.regs = regs;
    a.regIdx = regIdx;
    a.packets = static_cast<const uint8_t*>(j.dPackets);
    a.offsets = static_cast<const uint32_t*>(j.dOffsets);
    a.n = j.n;
    a.resBytes = j.resBytes;
    a.kv = j.cls == 0 ? MicroA::KV : MicroB::KV;
    a.smax = j.cls == 0 ? MicroA::SMAX : MicroB::SMAX;
    a.dec[0] = rDec_[0].p; a.dec[1] = rDec_[1].p;
    a.weight = rWeight_.p;
    a.regInfo = rInfo_.p; a.regCnt = rCnt_.p; a.regBase = rBase_.p; a.regRel = rRel_.p;
    a.iterations = rIter_.p;
    k_scatter<<<(j.n + 255) / 256, 256, 0, st>>>(a);
    std::lock_guard<std::mutex> lk(mMu_);
    ++stats.launches;
  }
  int resultsRegInfo(const uint32_t** info, int64_t* iterations) override {
    CUDA_TRY(cudaSetDevice(device_));
    for (int i = 0; i < kMicroLanes; ++i) CUDA_TRY(cudaStreamSynchronize(mStream_[i]));
    const size_t nR = (size_t) rs_.nReg;
    CUDA_TRY(hInfo_.ensure(nR + 2));
    cudaStream_t st = mStream_[0];
    unsigned long long it[2] = {0, 0};
    if (nR) CUDA_TRY(cudaMemcpyAsync(hInfo_.p, rInfo_.p, nR * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(it, rIter_.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    stats.d2hBytes += (int64_t) (nR * sizeof(uint32_t));
    *info = hInfo_.p;
    *iterations = (int64_t) it[0];
    return VCE_OK;
  }
  template <class T>
  int upVec(DevBuf<T>& d, const std::vector<T>& v, cudaStream_t st) {
    CUDA_TRY(d.ensure(v.size() + 4));
    if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    stats.h2dBytes += (int64_t) (v.size() * sizeof(T));
    return VCE_OK;
  }
  int resultsFinish(const ResultsPatch& p, ResultsOut& out) override {
    CUDA_TRY(cudaSetDevice(device_));
    cudaStream_t st = mStream_[0];
    const size_t nC = (size_t) rs_.nVar[0], nB = (size_t) rs_.nVar[1], nR = (size_t) rs_.nReg, nS = rs_.seq.size();
    int rc;
    if ((rc = upVec(pRegIdx_, p.regIdx, st)) || (rc = upVec(pInfo_, p.info, st)) || (rc = upVec(pSyncOff_, p.syncOff, st)) ||
        (rc = upVec(pSyncCnt_, p.syncCnt, st)) || (rc = upVec(pRange_, p.range, st)) || (rc = upVec(pPayOff_, p.payOff, st)) ||
        (rc = upVec(pDec_, p.decPay, st)) || (rc = upVec(pWOff_, p.wOff, st)) || (rc = upVec(pW_, p.wPay, st)) ||
        (rc = upVec(pSync_, p.extSync, st))) {
      return rc;
    }
    if (!p.regIdx.empty()) {
      PatchArgs a;
      a.n = (int) p.regIdx.size();
      a.regIdx = pRegIdx_.p; a.info = pInfo_.p; a.syncOff = pSyncOff_.p; a.syncCnt = pSyncCnt_.p; a.range = pRange_.p;
      a.payOff = pPayOff_.p; a.decPay = pDec_.p; a.wOff = pWOff_.p; a.wPay = pW_.p;
      a.dec[0] = rDec_[0].p; a.dec[1] = rDec_[1].p; a.weight = rWeight_.p;
      a.regInfo = rInfo_.p; a.regCnt = rCnt_.p; a.regBase = rBase_.p;
      k_patch<<<(a.n + 127) / 128, 128, 0, st>>>(a);
      CUDA_TRY(cudaGetLastError());
      ++stats.launches;
    }
    // scratch of the final passes
    const size_t nFlatMax = nC + nB + 2 * nR + 4;
    CUDA_TRY(fRegOff_.ensure(nR + 2)); CUDA_TRY(fKeys_.ensure(nFlatMax)); CUDA_TRY(fUniq_.ensure(nFlatMax)); CUDA_TRY(fPos_.ensure(nFlatMax));
    CUDA_TRY(fSeqSync_.ensure(nS + 2)); CUDA_TRY(fOutC_.ensure(nC * 16 + 16)); CUDA_TRY(fOutB_.ensure(nB * 8 + 16));
    CUDA_TRY(fCounters_.ensure(nS * 8 + 8)); CUDA_TRY(fKept_.ensure(nC + 2)); CUDA_TRY(fMStart_.ensure(nC + 2)); CUDA_TRY(fMFlags_.ensure(nC + 2));
    CUDA_TRY(fMSeq_.ensure(nC + 2)); CUDA_TRY(fMVal_.ensure(nC + 2)); CUDA_TRY(fMMin_.ensure(nC + 2)); CUDA_TRY(fBaseSum_.ensure(nFlatMax));
    CUDA_TRY(fCall0_.ensure(nFlatMax + nS + 2));
    const size_t nBlocksMax = nFlatMax / kPhaseBlock + nS + 2;
    CUDA_TRY(fBlockEnd_.ensure(nBlocksMax * 16)); CUDA_TRY(fBlockCnt_.ensure(nBlocksMax * 16 * 3)); CUDA_TRY(fSeqBlock_.ensure(nS + 2));
    const size_t nSuperMax = nBlocksMax / kPhaseGroup + nS + 2;
    CUDA_TRY(fSuperEnd_.ensure(nSuperMax * 16)); CUDA_TRY(fSuperCnt_.ensure(nSuperMax * 16 * 3)); CUDA_TRY(fSeqSuper_.ensure(nS + 2));
    if (rs_.rocTuples) CUDA_TRY(fRoc_.ensure(nC * 3 + 3));
    CUDA_TRY(ints_.ensure(16));
    FinalArrays A;
    A.nSeq = (int) nS; A.H = rs_.H; A.kind[0] = rs_.kind[0]; A.kind[1] = rs_.kind[1];
    A.seq = rSeq_.p;
    A.nVar[0] = rs_.nVar[0]; A.nVar[1] = rs_.nVar[1]; A.nReg = rs_.nReg;
    A.dec[0] = rDec_[0].p; A.dec[1] = rDec_[1].p; A.weight = rWeight_.p;
    A.regInfo = rInfo_.p; A.regCnt = rCnt_.p; A.regBase = rBase_.p; A.regRel = rRel_.p; A.extSync = pSync_.p;
    A.regOff = fRegOff_.p; A.syncKeys = fKeys_.p; A.syncUniq = fUniq_.p; A.syncPos = fPos_.p; A.seqSyncFirst = fSeqSync_.p;
    A.outCalls = fOutC_.p; A.outBase = fOutB_.p; A.counters = fCounters_.p; A.rocTuples = rs_.rocTuples ? fRoc_.p : nullptr;
    A.keptScan = fKept_.p; A.mStart = fMStart_.p; A.mFlags = fMFlags_.p; A.mSeq = fMSeq_.p; A.mVal = fMVal_.p; A.mMin = fMMin_.p;
    A.regBaseSum = fBaseSum_.p; A.regCall0 = fCall0_.p; A.blockEnd = fBlockEnd_.p; A.blockCnt = fBlockCnt_.p; A.seqBlockFirst = fSeqBlock_.p;
    A.superEnd = fSuperEnd_.p; A.superCnt = fSuperCnt_.p; A.seqSuperFirst = fSeqSuper_.p;
    finalOps_.st = st;
    finalOps_.dCount = ints_.p + 14;
    finalOps_.launches = 0;
    finalOps_.err = cudaSuccess;
    // (the offsets array is scanned in place over nReg + 1 entries: the last one has to be defined)
    CUDA_TRY(cudaMemsetAsync(fRegOff_.p + nR, 0, sizeof(int32_t), st));
    tlMark(st, "final passes next");
    const int nSync = finalPasses(finalOps_, A);
    tlMark(st, "final passes done");
    if (finalOps_.err != cudaSuccess) {
      char b[256];
      snprintf(b, sizeof(b), "final passes failed: %s", cudaGetErrorString(finalOps_.err));
      setErr(b);
      return VCE_ERR_CUDA;
    }
    stats.launches += finalOps_.launches;
    CUDA_TRY(hOutC_.ensure(nC * 16 + 16)); CUDA_TRY(hOutB_.ensure(nB * 8 + 16)); CUDA_TRY(hPos_.ensure((size_t) nSync + 4));
    CUDA_TRY(hSeqSync_.ensure(nS + 2)); CUDA_TRY(hCounters_.ensure(nS * 8 + 8));
    if (nC) CUDA_TRY(cudaMemcpyAsync(hOutC_.p, fOutC_.p, nC * 16, cudaMemcpyDeviceToHost, st));
    if (nB) CUDA_TRY(cudaMemcpyAsync(hOutB_.p, fOutB_.p, nB * 8, cudaMemcpyDeviceToHost, st));
    if (nSync) CUDA_TRY(cudaMemcpyAsync(hPos_.p, fPos_.p, (size_t) nSync * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(hSeqSync_.p, fSeqSync_.p, (nS + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (nS) CUDA_TRY(cudaMemcpyAsync(hCounters_.p, fCounters_.p, nS * 8 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (rs_.rocTuples) {
      CUDA_TRY(hRoc_.ensure(nC * 3 + 3));
      if (nC) CUDA_TRY(cudaMemcpyAsync(hRoc_.p, fRoc_.p, nC * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
      stats.d2hBytes += (int64_t) (nC * 3 * sizeof(double));
    }
    tlMark(st, "results copied back");
    CUDA_TRY(cudaStreamSynchronize(st));
    tlDump();
    stats.d2hBytes += (int64_t) (nC * 16 + nB * 8 + (size_t) nSync * 4 + (nS + 1) * 4 + nS * 64);
    out.outCalls = hOutC_.p; out.outBase = hOutB_.p; out.syncPos = hPos_.p; out.seqSyncFirst = hSeqSync_.p;
    out.counters = hCounters_.p; out.rocTuples = rs_.rocTuples ? hRoc_.p : nullptr; out.nSync = nSync;
    return VCE_OK;
  }
  // ---- VCE_TIMELINE=1: when each stream reached the marked points of one vce_submit (diagnostics; the events are
  // created per mark and destroyed by the dump)
  struct TlMark { cudaEvent_t ev; const char* what; int a; };
  std::vector<TlMark> tl_;
  std::mutex tlMu_;
  bool tlOn_ = std::getenv("VCE_TIMELINE") != nullptr;
  void tlMark(cudaStream_t st, const char* what, int a = 0) {
    if (!tlOn_) return;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    std::lock_guard<std::mutex> lk(tlMu_);
    tl_.push_back(TlMark{e, what, a});
  }
  void tlDump() {
    if (!tlOn_) return;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(tlMu_);
    std::vector<std::pair<float, size_t>> at;
    for (size_t i = 0; i < tl_.size(); ++i) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, tl_[0].ev, tl_[i].ev) == cudaSuccess) at.push_back(std::make_pair(ms, i));
    }
    std::sort(at.begin(), at.end());
    for (const auto& q : at) fprintf(stderr, "[vce timeline] %8.3f ms  %s %d\n", q.first, tl_[q.second].what, tl_[q.second].a);
    for (TlMark& m : tl_) cudaEventDestroy(m.ev);
    tl_.clear();
  }
  ResultsSetup rs_;
  std::vector<FinalSeq> fsHost_;
  cudaEvent_t rBeginEv_ = nullptr;
  DevBuf<uint8_t> rDec_[2], rRel_, pDec_, fOutC_, fOutB_, fMFlags_, fBaseSum_, fBlockEnd_, fSuperEnd_;
  DevBuf<int32_t> fSuperCnt_, fSeqSuper_;
  DevBuf<double> rWeight_, pW_, fRoc_;
  DevBuf<uint32_t> rInfo_, pInfo_;
  DevBuf<int32_t> rCnt_, rBase_, pRegIdx_, pSyncOff_, pSyncCnt_, pPayOff_, pWOff_, pSync_, fRegOff_, fPos_, fSeqSync_, fKept_, fMStart_, fMSeq_,
      fMVal_, fMMin_, fCall0_, fBlockCnt_, fSeqBlock_;
  DevBuf<int64_t> fKeys_, fUniq_, fCounters_;
  DevBuf<RegRange> pRange_;
  DevBuf<FinalSeq> rSeq_;
  DevBuf<unsigned long long> rIter_;
  PinBuf<uint32_t> hInfo_;
  PinBuf<uint8_t> hOutC_, hOutB_;
  PinBuf<int32_t> hPos_, hSeqSync_;
  PinBuf<int64_t> hCounters_;
  PinBuf<double> hRoc_;
  CudaFinalOps finalOps_;

  // ---- resident inputs
  bool templateLookup(const uint8_t* bases, int32_t len, const uint32_t** t2, const uint32_t** nmask) override {
    return templateRegistryLookup(device_, bases, len, t2, nmask);
  }
  int residentUpload(const void* pinned, size_t bytes, void** dev, int* eventId) override {
    CUDA_TRY(cudaSetDevice(device_));
    void* d = nullptr;
    cudaEvent_t ev = nullptr;
    int lane = 0;
    {
      std::lock_guard<std::mutex> lk(mMu_);
      d = devAlloc(bytes + 16);
      lane = mNextLane_;
      mNextLane_ = (mNextLane_ + 1) % kMicroLanes;
      if ((int) mEvents_.size() <= mNextEvent_) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { setErr("cudaEventCreate failed"); return VCE_ERR_CUDA; }
        mEvents_.push_back(e);
      }
      *eventId = mNextEvent_;
      ev = mEvents_[mNextEvent_++];
      stats.h2dBytes += (int64_t) bytes;
    }
    if (!d) { setErr("out of device memory for the resident inputs"); return VCE_ERR_CUDA; }
    tlMark(mStream_[lane], "sequence arrays: copy next; lane", lane);
    CUDA_TRY(cudaMemcpyAsync(d, pinned, bytes, cudaMemcpyHostToDevice, mStream_[lane]));
    CUDA_TRY(cudaEventRecord(ev, mStream_[lane]));
    tlMark(mStream_[lane], "sequence arrays: resident; MB", (int) (bytes >> 20));
    *dev = d;
    return VCE_OK;
  }

  bool hostPinned(const void* p, size_t bytes) const override { return pinnedRegistryCovers(p, bytes); }
  int residentAlloc(size_t bytes, void** dev, int* lane) override {
    CUDA_TRY(cudaSetDevice(device_));
    std::lock_guard<std::mutex> lk(mMu_);
    *dev = devAlloc(bytes + 16);
    *lane = mNextLane_;
    mNextLane_ = (mNextLane_ + 1) % kMicroLanes;
    if (!*dev) { setErr("out of device memory for the resident inputs"); return VCE_ERR_CUDA; }
    return VCE_OK;
  }
  int residentCopy(void* dst, const void* src, size_t bytes, int lane) override {
    if (bytes == 0) return VCE_OK;
    CUDA_TRY(cudaSetDevice(device_));
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, mStream_[lane]));
    std::lock_guard<std::mutex> lk(mMu_);
    stats.h2dBytes += (int64_t) bytes;
    return VCE_OK;
  }
  int residentMark(int lane, int* eventId) override {
    CUDA_TRY(cudaSetDevice(device_));
    cudaEvent_t ev = nullptr;
    {
      std::lock_guard<std::mutex> lk(mMu_);
      if ((int) mEvents_.size() <= mNextEvent_) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { setErr("cudaEventCreate failed"); return VCE_ERR_CUDA; }
        mEvents_.push_back(e);
      }
      *eventId = mNextEvent_;
      ev = mEvents_[mNextEvent_++];
    }
    CUDA_TRY(cudaEventRecord(ev, mStream_[lane]));
    tlMark(mStream_[lane], "sequence arrays (page-locked caller memory): resident; lane", lane);
    return VCE_OK;
  }

  // One search launch over the MicroB packets of every chunk built since the last flush.
  int microBuildFlush(const MicroConfig& mc) override {
    CUDA_TRY(cudaSetDevice(device_));
    std::vector<PendingB> pend;
    {
      std::lock_guard<std::mutex> lk(mMu_);
      pend.swap(pendingB_);
    }
    if (pend.empty()) return VCE_OK;
    cudaStream_t st = mStream_[1];
    std::vector<MicroChunkDesc> table(pend.size());
    int total = 0;
    for (size_t i = 0; i < pend.size(); ++i) {
      MicroJob& j = *pend[i].job;
      CUDA_TRY(cudaStreamWaitEvent(st, pend[i].built, 0));
      table[i] = MicroChunkDesc{static_cast<const uint8_t*>(j.dPackets), static_cast<const uint32_t*>(j.dOffsets),
                                static_cast<uint8_t*>(j.dResults), j.n, total, nullptr};
      total += j.n;
    }
    CUDA_TRY(mTableC_[1].ensure(table.size()));
    CUDA_TRY(cudaMemcpyAsync(mTableC_[1].p, table.data(), table.size() * sizeof(MicroChunkDesc), cudaMemcpyHostToDevice, st));
    MicroParams mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.one = table[0];
    mp.table = mTableC_[1].p;
    mp.nChunks = (int) table.size();
    mp.total = total;
    mp.refill = microRefill_;
    mp.cfg = mc;
    tlMark(st, "all MicroB packets built, k_micro<MicroB> next; regions", total);
    microLaunch(mp, 1, st);
    CUDA_TRY(cudaGetLastError());
    tlMark(st, "k_micro<MicroB> done");
    for (const PendingB& pb : pend) {
      MicroJob& j = *pb.job;
      if (pb.scatter) scatterLaunch(j, pb.regs, pb.regIdx, st);
      else CUDA_TRY(cudaMemcpyAsync(j.results, j.dResults, (size_t) j.n * j.resBytes, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(j.evHandle), st));
      j.lane = 1;
    }
    ++stats.launches;
    return VCE_OK;
  }
  bool oldMid_ = true;    // mid tier (frontiers up to 2048) through the HBM-arena warp kernel: 24 resident regions per SM
  bool ctaBig_ = true;    // full-budget tier through the CTA kernel while the batch fits one wave
  bool ctaTma_ = true;    // the CTA kernel moves frontier blocks with bulk TMA copies
  static constexpr int kOrderMin = 256;   // smaller chunks are not worth the ordering launch
  bool classOrder_ = true;
  struct PendingB { MicroJob* job; cudaEvent_t built; const RegRange* regs; const int32_t* regIdx; bool scatter; };
  std::vector<PendingB> pendingB_;
  int microWait(MicroJob& j) override {
    CUDA_TRY(cudaSetDevice(device_));
    CUDA_TRY(cudaEventSynchronize(static_cast<cudaEvent_t>(j.evHandle)));
    return VCE_OK;
  }
  double lastMicroMs_ = 0;
  long long lastMicroRegions_ = 0;

  const char* lastError() const override { return err_.c_str(); }
  // worker threads report failures concurrently (microUpload / microBuildRun): the first message is kept
  void setErr(const char* msg) {
    std::lock_guard<std::mutex> lk(errMu_);
    if (err_.empty()) err_ = msg;
  }
  void clearErr() {
    std::lock_guard<std::mutex> lk(errMu_);
    err_.clear();
  }
  void timerStart() override {
    cudaSetDevice(device_);
    cudaEventRecord(evT0_, stream_);
  }
  double timerStopMs() override {
    cudaSetDevice(device_);
    cudaEventRecord(evT1_, stream_);
    cudaEventSynchronize(evT1_);
    float ms = 0;
    cudaEventElapsedTime(&ms, evT0_, evT1_);
    return ms;
  }
  cudaStream_t stream() const { return stream_; }
  int device() const { return device_; }
  double lastFastMs_ = 0;
  long long lastFastRegions_ = 0;
  long long lastFastVariants_ = 0;

 private:
  int device_;
  bool ready_ = false;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, evT0_ = nullptr, evT1_ = nullptr;
  int smCount_ = 148;
  int maxSmemOptin_ = 0;
  AlleleDev al_[2];
  SideDev pdev_[2][2];   // [pass][side]
  int curPass_ = 0;
  DevBuf<RegionDesc> regs_, sRegs_;         // transient / staged launch inputs
  DevBuf<int32_t> sSyncOff_;
  DevBuf<uint8_t> sWindows_;
  DevBuf<RegionResult> rres_;
  DevBuf<int32_t> syncOff_, syncBuf_;
  size_t syncBufUsed_ = 0;
  DevBuf<uint8_t> windows_;
  DevBuf<WarnRec> warns_, warnScratch_;
  DevBuf<int> ints_;
  DevBuf<int32_t> arena_;
  DevBuf<uint8_t> scanTmp_;
  // thread-per-region path: device bump arena, copy/compute lanes, per-job events
  struct DevBlock { uint8_t* p; size_t cap, used; };
  static constexpr int kMicroLanes = 4;
  std::mutex mMu_;
  std::vector<DevBlock> dBlocks_;
  std::vector<size_t> dMarks_;
  int mMarkEvent_ = 0;
  cudaStream_t mStream_[kMicroLanes] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> mEvents_;
  int mNextEvent_ = 0, mNextLane_ = 0;
  cudaEvent_t evM0_[2] = {nullptr, nullptr}, evM1_[2] = {nullptr, nullptr};
  int microBlocksPerSm_[2] = {1, 1};
  int microRefill_ = 32;
  DevBuf<MicroChunkDesc> mTableC_[2];
  std::atomic<bool> uploadsPending_{false};
  std::mutex errMu_;
  std::string err_;
};

// ---- caller memory page-locked through vce_host_register
namespace {
struct PinnedRange { const uint8_t* p; size_t bytes; };
std::mutex g_pinMu;
std::vector<PinnedRange> g_pinned;
}  // namespace
bool pinnedRegistryCovers(const void* p, size_t bytes) {
  if (bytes == 0) return true;
  if (p == nullptr) return false;
  const uint8_t* q = static_cast<const uint8_t*>(p);
  std::lock_guard<std::mutex> lk(g_pinMu);
  for (const PinnedRange& r : g_pinned) if (q >= r.p && q + bytes <= r.p + r.bytes) return true;
  return false;
}
int pinnedRegistryAdd(const void* p, size_t bytes, std::string& err) {
  if (p == nullptr || bytes == 0) return VCE_OK;
  if (pinnedRegistryCovers(p, bytes)) return VCE_OK;
  const cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); err = std::string("cudaHostRegister failed: ") + cudaGetErrorString(e); return VCE_ERR_CUDA; }
  std::lock_guard<std::mutex> lk(g_pinMu);
  g_pinned.push_back(PinnedRange{static_cast<const uint8_t*>(p), bytes});
  return VCE_OK;
}
void pinnedRegistryRemove(const void* p) {
  std::lock_guard<std::mutex> lk(g_pinMu);
  for (size_t i = 0; i < g_pinned.size();) {
    if (p == nullptr || g_pinned[i].p == p) {
      cudaHostUnregister(const_cast<uint8_t*>(g_pinned[i].p));
      g_pinned.erase(g_pinned.begin() + (long) i);
    } else {
      ++i;
    }
  }
}

// ---- registered templates: device-resident 2-bit copies, per device, keyed by the caller's pointer + length
namespace {
struct TemplateEntry { const uint8_t* bases; int32_t len; int device; uint32_t* t2; uint32_t* nmask; };
std::mutex g_tmplMu;
std::vector<TemplateEntry> g_templates;
}  // namespace

bool templateRegistryLookup(int device, const uint8_t* bases, int32_t len, const uint32_t** t2, const uint32_t** nmask) {
  std::lock_guard<std::mutex> lk(g_tmplMu);
  for (const TemplateEntry& e : g_templates) {
    // (a shorter length is a prefix of the registered sequence: the pieces of a cut sequence end before the next piece's
    // first variant, vce_split.h)
    if (e.bases == bases && len <= e.len && e.device == device) {
      *t2 = e.t2;
      *nmask = e.nmask;
      return true;
    }
  }
  return false;
}

int templateRegistryAdd(int device, const uint8_t* bases, int32_t len, const uint32_t* packed, size_t words, const uint32_t* nmask,
                        size_t mwords, std::string& err) {
  templateRegistryRemove(device, bases);
  if (cudaSetDevice(device) != cudaSuccess) { err = "cudaSetDevice failed"; return VCE_ERR_CUDA; }
  uint32_t* d2 = nullptr;
  uint32_t* dn = nullptr;
  // (a few words of slack: the packet builder may look one granule beyond the last base)
  if (cudaMalloc(&d2, (words + 8) * 4) != cudaSuccess || cudaMalloc(&dn, (mwords + 8) * 4) != cudaSuccess) {
    if (d2) cudaFree(d2);
    err = "out of device memory for the resident template";
    return VCE_ERR_CUDA;
  }
  cudaMemset(d2 + words, 0, 32);
  cudaMemset(dn + mwords, 0, 32);
  if (cudaMemcpy(d2, packed, words * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(dn, nmask, mwords * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d2);
    cudaFree(dn);
    err = "template upload failed";
    return VCE_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_tmplMu);
  g_templates.push_back(TemplateEntry{bases, len, device, d2, dn});
  return VCE_OK;
}

// The arrays were made on the device (SDF decode, vce_vcf.cu) and now belong to the registry.
int templateRegistryAdopt(int device, const uint8_t* bases, int32_t len, uint32_t* d2, uint32_t* dn, std::string&) {
  templateRegistryRemove(device, bases);
  std::lock_guard<std::mutex> lk(g_tmplMu);
  g_templates.push_back(TemplateEntry{bases, len, device, d2, dn});
  return VCE_OK;
}

void templateRegistryRemove(int device, const uint8_t* bases) {
  std::lock_guard<std::mutex> lk(g_tmplMu);
  for (size_t i = 0; i < g_templates.size();) {
    if ((device < 0 || g_templates[i].device == device) && (bases == nullptr || g_templates[i].bases == bases)) {
      cudaSetDevice(g_templates[i].device);
      cudaFree(g_templates[i].t2);
      cudaFree(g_templates[i].nmask);
      g_templates.erase(g_templates.begin() + (long) i);
    } else {
      ++i;
    }
  }
}

Backend* createCudaBackend(int device, std::string& err) {
  CudaBackend* b = new CudaBackend(device);
  if (b->init() != VCE_OK) {
    err = b->lastError();
    delete b;
    return nullptr;
  }
  return b;
}

int cudaDeviceCount() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void cudaBackendFastStats(Backend* b, double* ms, long long* regions, long long* variants) {
  CudaBackend* c = static_cast<CudaBackend*>(b);
  *ms = c->lastFastMs_;
  *regions = c->lastFastRegions_;
  if (variants) *variants = c->lastFastVariants_;
}

void cudaBackendMicroStats(Backend* b, double* ms, long long* regions) {
  CudaBackend* c = static_cast<CudaBackend*>(b);
  c->microCollect();
  *ms = c->lastMicroMs_;
  *regions = c->lastMicroRegions_;
}

}  // namespace vce
