// vce_final.h -- what follows the search, on the device: the step right after PathFinder.bestPath in the reference.
//
//   * per-variant results of the thread-per-region kernel scattered straight from its result records into pass-wide
//     arrays (no host decode), the regions the host settled (warp / CTA kernels, merges, prefix re-runs) patched in;
//   * Path.getSyncPoints() of every sequence: the regions' sync points flattened in region order, de-duplicated at the
//     region joints (E/Path.java:293);
//   * SequenceEvaluator.mergeVariants / insertSkipped statuses (E/SequenceEvaluator.java:170-209), the orientation row,
//     OrientedVariant.isOriginal, call weights, and the sync-region id of every variant
//     (InterleavingEvalSynchronizer.floorSyncPos, E/InterleavingEvalSynchronizer.java:71-83);
//   * the TP / FP / FN classes of the split writers and their counters (E/SplitEvalSynchronizer.java:112-156) and the
//     (tp, fp, tpRaw) tuple of every call as WithRocsEvalSynchronizer.addToROCContainer feeds it to the ROC
//     (E/WithRocsEvalSynchronizer.java:132-141) -- K4 can consume them without a host round trip;
//   * PhasingEvaluator.countMisphasings (E/PhasingEvaluator.java:55-142) as data-parallel passes: baseline sections per sync
//     region by binary search, the call -> sync-region assignment of its do/while loops as a prefix minimum
//     (region(k) = k + min(1, min_{j<=k}(natural(j) - j)), natural = first sync point beyond the call's start), and the
//     4-bit state machine (baseIsPhased, basePhase, callIsPhased, callPhase) as a 16-state transducer evaluated per block
//     of sync regions and composed per sequence.
//
// Every pass is ONE source: an element-wise body run through `Ops::parFor` plus a few scans (`Ops`), instantiated with
// CUDA kernels + CUB in the product (vce_final.cu) and with plain loops in the test-only host emulation.
#ifndef VCE_FINAL_H_
#define VCE_FINAL_H_

#include <stdint.h>

#include "../../include/vce.h"
#include "vce_device.h"
#include "vce_micro.h"
#include "vce_packet.h"

#if defined(__CUDACC__)
#define VCE_LAMBDA [=] __host__ __device__
#else
#define VCE_LAMBDA [=]
#endif

namespace vce {

// ---- region table (one entry per region of the pass, global region index = sequence's first region + j)
enum : uint32_t { RI_NONE = 0, RI_CLEAN = 1, RI_FALLBACK = 2, RI_HOST = 3, RI_DEAD = 4 };
VCE_HD uint32_t riPack(uint32_t status, uint32_t flags, int lastId) { return status | (flags << 8) | (((uint32_t) lastId & 0xffu) << 16); }
VCE_HD uint32_t riStatus(uint32_t v) { return v & 0xffu; }
VCE_HD uint32_t riFlags(uint32_t v) { return (v >> 8) & 0xffu; }
VCE_HD int riLastId(uint32_t v) { return (int) (int8_t) ((v >> 16) & 0xffu); }

constexpr int kFinalRel = 8;          // relative sync points kept per thread-per-region result
constexpr int64_t kSyncBias = 16;     // sync positions are >= -1: biased so that the sequence id can sit above them
constexpr int kPhaseBlock = 32;       // sync regions per transducer block
constexpr int kPhaseGroup = 64;       // blocks per super-block of the composition

struct FinalSeq {   // one sequence of the pass; side 0 = calls, 1 = baseline
  const int32_t* vStart[2];     // [n] locus start, sequence-local index
  const uint8_t* phased[2];     // [n] or nullptr
  const uint8_t* statusIn[2];   // [n] or nullptr
  int32_t first[2], n[2];       // slice of the pass-wide per-variant arrays
  int32_t regFirst, nReg;       // slice of the region table
};

struct FinalArrays {
  int nSeq;
  int H, kind[2];               // haplotypes and orientor per side (for isOriginal)
  const FinalSeq* seq;          // [nSeq]
  int32_t nVar[2];              // pass-wide variants per side
  int32_t nReg;                 // regions of the pass
  // per-variant search results, pass-wide
  uint8_t* dec[2];              // 0 untouched, 1 excluded, 2 + row included
  double* weight;               // calls
  // region table
  uint32_t* regInfo;            // riPack
  int32_t* regCnt;              // sync points of the region
  int32_t* regBase;             // RI_CLEAN: position of relative sync point 0; RI_HOST: offset into extSync
  uint8_t* regRel;              // [nReg * kFinalRel] RI_CLEAN: relative sync points
  const int32_t* extSync;       // sync points of the regions the host settled
  // scratch / outputs (all sized by the caller)
  int32_t* regOff;              // [nReg + 1]
  int64_t* syncKeys;            // [nVar[0] + nVar[1] + 2 * nReg + 2] flattened (sequence << 32 | position + bias)
  int64_t* syncUniq;            // same size: de-duplicated
  int32_t* syncPos;             // positions of syncUniq
  int32_t* seqSyncFirst;        // [nSeq + 1] slice of syncPos per sequence
  uint8_t* outCalls;            // [nVar[0] * 16] status, row, original, pad, syncId (i32), weight (f64)
  uint8_t* outBase;             // [nVar[1] * 8]  status, row, original, pad, syncId (i32)
  int64_t* counters;            // [nSeq * 8] baseline TP, FN, call TP, FP, skipped, mis-, correct, unphaseable
  double* rocTuples;            // optional [nVar[0] * 3] (tp weight, fp weight, tp raw), all 0 when the call feeds nothing
  // phasing scratch
  int32_t* keptScan;            // [nVar[0] + 1]
  int32_t* mStart;              // [nVar[0]] merged-order call starts
  uint8_t* mFlags;              // [nVar[0]] bit0 phased, bit1 included, bit2 phase
  int32_t* mSeq;                // [nVar[0]] sequence of the merged call
  int32_t* mVal;                // [nVar[0]] natural region - local index, then its running minimum
  int32_t* mMin;                // [nVar[0]]
  uint8_t* regBaseSum;          // [sync points] bit0 inPhase, bit1 phase, bit2 has baseline variants
  int32_t* regCall0;            // [sync points + nSeq] first merged call of the sync region (sequence-local)
  uint8_t* blockEnd;            // [blocks * 16] end state per start state
  int32_t* blockCnt;            // [blocks * 16 * 3] counts per start state
  int32_t* seqBlockFirst;       // [nSeq + 1]
  uint8_t* superEnd;            // [super-blocks * 16]
  int32_t* superCnt;            // [super-blocks * 16 * 3]
  int32_t* seqSuperFirst;       // [nSeq + 1]
};

VCE_HD int finalSeqOfRegion(const FinalSeq* seq, int nSeq, int r) {
  int lo = 0, hi = nSeq - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seq[mid].regFirst <= r) lo = mid; else hi = mid - 1;
  }
  return lo;
}
VCE_HD int finalSeqOfVariant(const FinalSeq* seq, int nSeq, int side, int v) {
  int lo = 0, hi = nSeq - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seq[mid].first[side] <= v) lo = mid; else hi = mid - 1;
  }
  return lo;
}
// OrientedVariant.isOriginal() of orientation `row` (Orientor.java: only the diploid unphased expansion's first row and
// the phased orientor as-is are "original")
VCE_HD bool finalIsOriginal(int kind, int H, bool phased, int row) {
  if (kind == VCE_ORIENT_PHASED && phased) return true;
  if (kind == VCE_ORIENT_PHASED_INVERT && phased) return false;
  if (kind == VCE_ORIENT_UNPHASED || kind == VCE_ORIENT_PHASED || kind == VCE_ORIENT_PHASED_INVERT) return H == 2 && row == 0;
  return false;
}
template <class T>
VCE_HD int finalUpperBound(const T* a, int n, T x) {  // first i with a[i] > x
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
template <class T>
VCE_HD int finalLowerBound(const T* a, int n, T x) {  // first i with a[i] >= x
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- scatter: one thread per packet of a chunk, straight from the search kernel's result record
struct ScatterArgs {
  const uint8_t* results;   // [n * resBytes]
  const RegRange* regs;     // [n] pass-wide variant ranges
  const int32_t* regIdx;    // [n] global region index
  const uint8_t* packets;   // device packets (winStart = position of relative 0) ...
  const uint32_t* offsets;  // ... through their word offsets (kNoPacket-style sentinel 0xffffffff = not built)
  int n, resBytes, kv, smax;
  uint8_t* dec[2];
  double* weight;
  uint32_t* regInfo;
  int32_t* regCnt;
  int32_t* regBase;
  uint8_t* regRel;
  unsigned long long* iterations;   // summed over the chunk
};
VCE_HD void scatterOne(const ScatterArgs& a, int q) {
  const uint8_t* rec = a.results + (size_t) q * a.resBytes;
  const int r = a.regIdx[q];
  const int it = rec[MR_ITERS] | (rec[MR_ITERS + 1] << 8);
#if defined(__CUDA_ARCH__)
  {  // one atomic per warp: millions of them on one address would serialise
    const unsigned active = __activemask();
    const unsigned sum = __reduce_add_sync(active, (unsigned) it);
    if ((int) (threadIdx.x & 31) == __ffs((int) active) - 1) atomicAdd(a.iterations, (unsigned long long) sum);
  }
#else
  *a.iterations += (unsigned long long) it;
#endif
  if (rec[MR_STATUS] != kMicroClean) {
    a.regInfo[r] = riPack(RI_FALLBACK, 0, kMicroNoLastId);
    a.regCnt[r] = 0;
    return;
  }
  const RegRange g = a.regs[q];
  const uint8_t* dec = rec + MR_SYNC + a.smax;
  const uint8_t* wn = dec + 2 * a.kv;
  const uint8_t* wd = wn + a.kv;
  for (int i = 0; i < g.cCount; ++i) {
    a.dec[0][g.cFirst + i] = dec[i];
    a.weight[g.cFirst + i] = dec[i] >= 2 ? (double) wn[i] / (double) wd[i] : 0.0;   // Path.java:450
  }
  for (int i = 0; i < g.bCount; ++i) a.dec[1][g.bFirst + i] = dec[a.kv + i];
  const int ns = rec[MR_NSYNC];
  a.regInfo[r] = riPack(RI_CLEAN, rec[MR_FLAGS] & 3u, (int) (int8_t) rec[MR_LASTID]);
  a.regCnt[r] = ns;
  const uint8_t* pk = a.packets + (size_t) a.offsets[q] * 4;
  a.regBase[r] = (int32_t) ((uint32_t) pk[MP_WINSTART] | ((uint32_t) pk[MP_WINSTART + 1] << 8) | ((uint32_t) pk[MP_WINSTART + 2] << 16) |
                            ((uint32_t) pk[MP_WINSTART + 3] << 24));
  for (int i = 0; i < kFinalRel; ++i) a.regRel[(size_t) r * kFinalRel + i] = i < ns && i < a.smax ? rec[MR_SYNC + i] : 0;
}

// ---- patch: the regions the host settled (one thread per patched region)
struct PatchArgs {
  int n;
  const int32_t* regIdx;     // [n]
  const uint32_t* info;      // [n] riPack(RI_HOST | RI_DEAD, flags, lastId)
  const int32_t* syncOff;    // [n] offset into extSync
  const int32_t* syncCnt;    // [n]
  const RegRange* range;     // [n] variants the region owns now (empty for dead regions)
  const int32_t* payOff;     // [n] offset of the region's decisions in decPay (calls first, then baseline) ...
  const uint8_t* decPay;
  const int32_t* wOff;       // [n] ... and of its call weights in wPay
  const double* wPay;
  uint8_t* dec[2];
  double* weight;
  uint32_t* regInfo;
  int32_t* regCnt;
  int32_t* regBase;
};
VCE_HD void patchOne(const PatchArgs& a, int q) {
  const int r = a.regIdx[q];
  a.regInfo[r] = a.info[q];
  a.regCnt[r] = riStatus(a.info[q]) == RI_HOST ? a.syncCnt[q] : 0;
  a.regBase[r] = a.syncOff[q];
  const RegRange g = a.range[q];
  const uint8_t* d = a.decPay + a.payOff[q];
  for (int i = 0; i < g.cCount; ++i) {
    a.dec[0][g.cFirst + i] = d[i];
    a.weight[g.cFirst + i] = a.wPay[a.wOff[q] + i];
  }
  for (int i = 0; i < g.bCount; ++i) a.dec[1][g.bFirst + i] = d[g.cCount + i];
}

// One counter update per warp instead of one per thread when the lanes of a warp work on the same sequence (they almost
// always do): millions of atomics on a handful of addresses would otherwise serialise in L2.
VCE_HD void finalCount(int64_t* cn, int idx, bool pred, int k) {
#if defined(__CUDA_ARCH__)
  const unsigned active = __activemask();
  const int leader = __ffs((int) active) - 1;
  const int k0 = __shfl_sync(active, k, leader);
  const bool same = __all_sync(active, k == k0);
  if (same) {
    const unsigned m = __ballot_sync(active, pred);
    if ((int) (threadIdx.x & 31) == leader && m != 0) atomicAdd(reinterpret_cast<unsigned long long*>(cn + idx), (unsigned long long) __popc(m));
  } else if (pred) {
    atomicAdd(reinterpret_cast<unsigned long long*>(cn + idx), 1ull);
  }
#else
  (void) k;
  if (pred) ++cn[idx];
#endif
}

// ---- the final passes.  Ops provides: parFor(n, body), exclusiveSum(in, out, n), unique64(in, out, n) -> count (blocking),
// segMinScan(keys, vals, out, n), fetchInt(ptr) (blocking read of one device int), zero(ptr, bytes).
template <class Ops>
int finalPasses(Ops& ops, const FinalArrays& A) {
  const FinalArrays a = A;
  const int nSeq = a.nSeq;
  // 1. sync points: counts -> offsets -> flattened keys (sequence, position) in region order -> unique
  ops.parFor(a.nReg, VCE_LAMBDA(int r) {
    const uint32_t st = riStatus(a.regInfo[r]);
    a.regOff[r] = (st == RI_CLEAN || st == RI_HOST) ? a.regCnt[r] : 0;
  });
  ops.exclusiveSum(a.regOff, a.regOff, a.nReg + 1);
  const int nFlat = ops.fetchInt(a.regOff + a.nReg);
  ops.parFor(a.nReg, VCE_LAMBDA(int r) {
    const uint32_t st = riStatus(a.regInfo[r]);
    if (st != RI_CLEAN && st != RI_HOST) return;
    const int k = finalSeqOfRegion(a.seq, nSeq, r);
    const int n = a.regCnt[r];
    int64_t* out = a.syncKeys + a.regOff[r];
    for (int i = 0; i < n; ++i) {
      const int pos = st == RI_CLEAN ? a.regBase[r] + (int) a.regRel[(size_t) r * kFinalRel + i] : a.extSync[a.regBase[r] + i];
      out[i] = ((int64_t) k << 32) | (int64_t) (uint32_t) (pos + kSyncBias);
    }
  });
  const int nSync = ops.unique64(a.syncKeys, a.syncUniq, nFlat);
  ops.parFor(nSync, VCE_LAMBDA(int i) { a.syncPos[i] = (int32_t) (uint32_t) (a.syncUniq[i] & 0xffffffffLL) - (int32_t) kSyncBias; });
  ops.parFor(nSeq + 1, VCE_LAMBDA(int k) { a.seqSyncFirst[k] = finalLowerBound<int64_t>(a.syncUniq, nSync, (int64_t) k << 32); });

  // 2. per-variant outputs, classes, counters, ROC tuples
  ops.zero(a.counters, (size_t) nSeq * 8 * sizeof(int64_t));
  for (int side = 0; side < 2; ++side) {
    ops.parFor(a.nVar[side], VCE_LAMBDA(int v) {
      const int k = finalSeqOfVariant(a.seq, nSeq, side, v);
      const FinalSeq& s = a.seq[k];
      const int li = v - s.first[side];
      const int d0 = a.dec[side][v];
      const int stIn = s.statusIn[side] ? s.statusIn[side][li] : 0;
      const bool ph = s.phased[side] != nullptr && s.phased[side][li] != 0;
      int st = stIn;
      int row = 0xff, orig = 0;
      double wgt = 0.0;
      if (d0 >= 2) {
        st |= VCE_STATUS_GT_MATCH;
        row = d0 - 2;
        orig = finalIsOriginal(a.kind[side], a.H, ph, row) ? 1 : 0;
        if (side == 0) wgt = a.weight[v];
      } else {
        st |= d0 == 1 ? VCE_STATUS_NO_MATCH : VCE_STATUS_SKIPPED;
      }
      // InterleavingEvalSynchronizer.floorSyncPos :71-83 over the sequence's sync points
      int syncId = 0;
      const int s0 = a.seqSyncFirst[k], sn = a.seqSyncFirst[k + 1] - s0;
      if (!(st & VCE_STATUS_SKIPPED) && sn > 0) {
        const int u = finalUpperBound<int32_t>(a.syncPos + s0, sn, s.vStart[side][li] - 1);
        syncId = a.syncPos[s0 + (u == 0 ? 0 : u - 1)] + 2;
      } else if (!(st & VCE_STATUS_SKIPPED)) {
        syncId = 1;
      }
      int64_t* cn = a.counters + (size_t) k * 8;
      if (side == 0) {
        uint8_t* o = a.outCalls + (size_t) v * 16;
        o[0] = (uint8_t) st; o[1] = (uint8_t) row; o[2] = (uint8_t) orig; o[3] = 0;
        *reinterpret_cast<int32_t*>(o + 4) = syncId;
        *reinterpret_cast<double*>(o + 8) = wgt;
        // SplitEvalSynchronizer.handleKnownCall :112-134 / WithRocsEvalSynchronizer.addToROCContainer :132-141
        double tp = 0.0, fp = 0.0, raw = 0.0;
        if (st & VCE_STATUS_GT_MATCH) {
          if (wgt > 0) { tp = wgt; raw = 1.0; }
        } else if (st & VCE_STATUS_NO_MATCH) {
          if (!(st & VCE_STATUS_OUTSIDE_EVAL)) fp = 1.0;
        }
        if (a.rocTuples) { a.rocTuples[(size_t) v * 3] = tp; a.rocTuples[(size_t) v * 3 + 1] = fp; a.rocTuples[(size_t) v * 3 + 2] = raw; }
        finalCount(cn, 2, raw > 0, k);
        finalCount(cn, 3, fp > 0, k);
        finalCount(cn, 4, (st & VCE_STATUS_SKIPPED) != 0, k);
      } else {
        uint8_t* o = a.outBase + (size_t) v * 8;
        o[0] = (uint8_t) st; o[1] = (uint8_t) row; o[2] = (uint8_t) orig; o[3] = 0;
        *reinterpret_cast<int32_t*>(o + 4) = syncId;
        // SplitEvalSynchronizer.handleKnownBaseline :136-150
        const bool inside = !(st & VCE_STATUS_OUTSIDE_EVAL);
        const bool tpb = inside && (st & VCE_STATUS_GT_MATCH);
        const bool fnb = inside && !(st & VCE_STATUS_GT_MATCH) && (st & VCE_STATUS_NO_MATCH);
        finalCount(cn, 0, tpb, k);
        finalCount(cn, 1, fnb, k);
        finalCount(cn, 4, (st & VCE_STATUS_SKIPPED) != 0, k);
      }
    });
  }

  // 3. phasing counts
  // 3a. calls in the order PhasingEvaluator.CallIterator delivers them (:160-195): by start, included before excluded on ties
  ops.parFor(a.nVar[0] + 1, VCE_LAMBDA(int v) { a.keptScan[v] = v < a.nVar[0] && a.dec[0][v] != 0 ? 1 : 0; });
  ops.exclusiveSum(a.keptScan, a.keptScan, a.nVar[0] + 1);
  const int nKept = ops.fetchInt(a.keptScan + a.nVar[0]);
  ops.parFor(a.nVar[0], VCE_LAMBDA(int v) {
    const int d0 = a.dec[0][v];
    if (d0 == 0) return;
    const int k = finalSeqOfVariant(a.seq, nSeq, 0, v);
    const FinalSeq& s = a.seq[k];
    const int li = v - s.first[0];
    const int32_t* vs = s.vStart[0];
    const int start = vs[li];
    // tie run [ra, rb) of equal starts within the sequence
    int ra = li, rb = li + 1;
    while (ra > 0 && vs[ra - 1] == start) --ra;
    while (rb < s.n[0] && vs[rb] == start) ++rb;
    int rank = a.keptScan[s.first[0] + ra];
    const uint8_t* dc = a.dec[0] + s.first[0];
    if (rb - ra > 1) {
      int inclBefore = 0, inclTotal = 0, exclBefore = 0;
      for (int i = ra; i < rb; ++i) {
        if (dc[i] >= 2) { ++inclTotal; if (i < li) ++inclBefore; }
        else if (dc[i] == 1 && i < li) ++exclBefore;
      }
      rank += d0 >= 2 ? inclBefore : inclTotal + exclBefore;
    }
    const bool ph = s.phased[0] != nullptr && s.phased[0][li] != 0;
    const bool incl = d0 >= 2;
    const bool phase = incl && finalIsOriginal(a.kind[0], a.H, ph, d0 - 2);
    a.mStart[rank] = start;
    a.mFlags[rank] = (uint8_t) ((ph ? 1 : 0) | (incl ? 2 : 0) | (phase ? 4 : 0));
    a.mSeq[rank] = k;
    // natural sync region of the call (first sync point beyond its start) minus its index within the sequence
    const int s0 = a.seqSyncFirst[k], sn = a.seqSyncFirst[k + 1] - s0;
    const int nat = finalUpperBound<int32_t>(a.syncPos + s0, sn, start);
    a.mVal[rank] = nat - (rank - a.keptScan[s.first[0]]);
  });
  ops.segMinScan(a.mSeq, a.mVal, a.mMin, nKept);
  // 3b. baseline section of every sync region (:143-155): all phased the same way?
  ops.parFor(nSync, VCE_LAMBDA(int i) {
    const int k = (int) (a.syncUniq[i] >> 32);
    const FinalSeq& s = a.seq[k];
    const int s0 = a.seqSyncFirst[k];
    const int32_t* vs = s.vStart[1];
    const int lo = i == s0 ? 0 : finalLowerBound<int32_t>(vs, s.n[1], a.syncPos[i - 1]);
    const int hi = finalLowerBound<int32_t>(vs, s.n[1], a.syncPos[i]);
    int nb = 0;
    bool inPhase = true, phase = false;
    const uint8_t* db = a.dec[1] + s.first[1];
    for (int b = lo; b < hi; ++b) {
      const int d0 = db[b];
      if (d0 == 0) continue;
      const bool ph = s.phased[1] != nullptr && s.phased[1][b] != 0;
      const bool p = d0 >= 2 && finalIsOriginal(a.kind[1], a.H, ph, d0 - 2);
      if (nb == 0) { inPhase = ph; phase = p; }
      else if (!ph || p != phase) inPhase = false;
      ++nb;
    }
    a.regBaseSum[i] = (uint8_t) ((inPhase ? 1 : 0) | (phase ? 2 : 0) | (nb > 0 ? 4 : 0));
    // first merged call of this sync region: region(k) = k + min(1, running minimum), non-decreasing in k
    const int c0 = a.keptScan[s.first[0]], cn = a.keptScan[s.first[0] + s.n[0]] - c0;
    const int il = i - s0;
    int a0 = 0, b0 = cn;
    while (a0 < b0) {
      const int mid = (a0 + b0) >> 1;
      const int mm = a.mMin[c0 + mid] < 1 ? a.mMin[c0 + mid] : 1;
      if (mid + mm < il) a0 = mid + 1; else b0 = mid;
    }
    a.regCall0[i + k] = a0;
    if (i + 1 == a.seqSyncFirst[k + 1]) {   // sentinel: one past the last region's calls
      int a1 = a0, b1 = cn;
      while (a1 < b1) {
        const int mid = (a1 + b1) >> 1;
        const int mm = a.mMin[c0 + mid] < 1 ? a.mMin[c0 + mid] : 1;
        if (mid + mm < il + 1) a1 = mid + 1; else b1 = mid;
      }
      a.regCall0[i + k + 1] = a1;
    }
  });
  // 3c. 16-state transducer per block of sync regions, 3d. composition per sequence
  ops.parFor(nSeq + 1, VCE_LAMBDA(int k) {
    int acc = 0;
    for (int q = 0; q < k; ++q) acc += (a.seqSyncFirst[q + 1] - a.seqSyncFirst[q] + kPhaseBlock - 1) / kPhaseBlock;
    a.seqBlockFirst[k] = acc;
  });
  int nBlocks = 0;
  {
    // (host copy of the per-sequence sync counts is not needed: the block count is bounded by nSync / kPhaseBlock + nSeq)
    nBlocks = nSync / kPhaseBlock + nSeq + 1;
  }
  ops.parFor(nBlocks, VCE_LAMBDA(int blk) {
    if (blk >= a.seqBlockFirst[nSeq]) return;
    int k = 0;
    {
      int lo = 0, hi = nSeq - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.seqBlockFirst[mid] <= blk) lo = mid; else hi = mid - 1;
      }
      k = lo;
      while (k + 1 < nSeq && a.seqBlockFirst[k + 1] <= blk) ++k;   // sequences without sync points share a first block
    }
    const FinalSeq& s = a.seq[k];
    const int s0 = a.seqSyncFirst[k], sn = a.seqSyncFirst[k + 1] - s0;
    const int r0 = (blk - a.seqBlockFirst[k]) * kPhaseBlock;
    const int r1 = r0 + kPhaseBlock < sn ? r0 + kPhaseBlock : sn;
    const int c0 = a.keptScan[s.first[0]];
    // state bits: 0 baseIsPhased, 1 basePhase, 2 callIsPhased, 3 callPhase
    uint8_t st[16];
    int mis[16], cor[16], unp[16];
    for (int t = 0; t < 16; ++t) { st[t] = (uint8_t) t; mis[t] = cor[t] = unp[t] = 0; }
    for (int r = r0; r < r1; ++r) {
      const int gi = s0 + r;
      const uint8_t bsum = a.regBaseSum[gi];
      const bool inPhase = (bsum & 1) != 0, phase = (bsum & 2) != 0, hasBase = (bsum & 4) != 0;
      const int cA = a.regCall0[gi + k], cB = a.regCall0[gi + k + 1];
      for (int t = 0; t < 16; ++t) {
        bool bIs = (st[t] & 1) != 0, bPh = (st[t] & 2) != 0, cIs = (st[t] & 4) != 0, cPh = (st[t] & 8) != 0;
        bool transition = false;
        if (inPhase) {
          if (!bIs) cIs = false;
          if (hasBase) { transition = bPh != phase; bIs = true; bPh = phase; }
        }
        for (int c = cA; c < cB; ++c) {
          const uint8_t f = a.mFlags[c0 + c];
          const bool cph = (f & 1) != 0, cinc = (f & 2) != 0, cphase = (f & 4) != 0;
          if (!inPhase) {
            if (cph) ++unp[t];
          } else {
            if (!cph) {
              cIs = false;
            } else if (!cIs) {
              cIs = true;
              cPh = cphase;
            } else {
              const bool ct = cphase != cPh;
              if (cinc && ct != transition) { ++mis[t]; cPh = cphase; }
              else if (cinc) { ++cor[t]; cPh = cphase; }
            }
            transition = false;
          }
        }
        if (!inPhase) { bIs = false; cIs = false; }
        st[t] = (uint8_t) ((bIs ? 1 : 0) | (bPh ? 2 : 0) | (cIs ? 4 : 0) | (cPh ? 8 : 0));
      }
    }
    for (int t = 0; t < 16; ++t) {
      a.blockEnd[(size_t) blk * 16 + t] = st[t];
      a.blockCnt[((size_t) blk * 16 + t) * 3] = mis[t];
      a.blockCnt[((size_t) blk * 16 + t) * 3 + 1] = cor[t];
      a.blockCnt[((size_t) blk * 16 + t) * 3 + 2] = unp[t];
    }
  });
  // 3d. composition, two levels: kPhaseGroup blocks -> one super-block per start state (in parallel), then one short
  // sequential walk per sequence (a chromosome has tens of thousands of blocks: walking them one by one is a chain of
  // dependent loads that takes longer than every other pass together)
  ops.parFor(nSeq + 1, VCE_LAMBDA(int k) {
    int acc = 0;
    for (int q = 0; q < k; ++q) {
      const int nbk = a.seqBlockFirst[q + 1] - a.seqBlockFirst[q];
      acc += (nbk + kPhaseGroup - 1) / kPhaseGroup;
    }
    a.seqSuperFirst[k] = acc;
  });
  const int nSuperMax = nBlocks / kPhaseGroup + nSeq + 1;
  ops.parFor(nSuperMax * 16, VCE_LAMBDA(int i) {
    const int sup = i >> 4, t = i & 15;
    if (sup >= a.seqSuperFirst[nSeq]) return;
    int k = 0;
    {
      int lo = 0, hi = nSeq - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.seqSuperFirst[mid] <= sup) lo = mid; else hi = mid - 1;
      }
      k = lo;
    }
    const int b0 = a.seqBlockFirst[k] + (sup - a.seqSuperFirst[k]) * kPhaseGroup;
    const int b1 = b0 + kPhaseGroup < a.seqBlockFirst[k + 1] ? b0 + kPhaseGroup : a.seqBlockFirst[k + 1];
    int state = t, mis = 0, cor = 0, unp = 0;
    for (int blk = b0; blk < b1; ++blk) {
      const size_t e = (size_t) blk * 16 + (size_t) state;
      mis += a.blockCnt[e * 3]; cor += a.blockCnt[e * 3 + 1]; unp += a.blockCnt[e * 3 + 2];
      state = a.blockEnd[e];
    }
    a.superEnd[(size_t) sup * 16 + t] = (uint8_t) state;
    a.superCnt[((size_t) sup * 16 + t) * 3] = mis;
    a.superCnt[((size_t) sup * 16 + t) * 3 + 1] = cor;
    a.superCnt[((size_t) sup * 16 + t) * 3 + 2] = unp;
  });
  ops.parFor(nSeq, VCE_LAMBDA(int k) {
    int state = 0;
    int64_t mis = 0, cor = 0, unp = 0;
    for (int sup = a.seqSuperFirst[k]; sup < a.seqSuperFirst[k + 1]; ++sup) {
      const size_t e = (size_t) sup * 16 + (size_t) state;
      mis += a.superCnt[e * 3]; cor += a.superCnt[e * 3 + 1]; unp += a.superCnt[e * 3 + 2];
      state = a.superEnd[e];
    }
    a.counters[(size_t) k * 8 + 5] = mis;
    a.counters[(size_t) k * 8 + 6] = cor;
    a.counters[(size_t) k * 8 + 7] = unp;
  });
  return nSync;
}

}  // namespace vce
#endif
