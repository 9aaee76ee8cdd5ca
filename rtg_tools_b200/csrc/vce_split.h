// vce_split.h -- SURVEY.md 8(e): "if max-load / mean > 1.1 split the largest sequences at verified region boundaries".
//
// Sequences are independent (VcfEvalTask.java:131-135), so several devices evaluate disjoint sets of them; one long
// chromosome, however, bounds the balance (chr1 holds ~8 % of a genome's variants).  Here a long sequence is cut into PIECES
// that are evaluated as sequences of their own -- on any device, by the unchanged engine -- and stitched back:
//
//   * cuts lie in wide gaps: no variant of either side within `gap` bases before the next variant's start (the engine's own
//     region rule, SURVEY A.13, with a much wider gap), so that a fresh in-sync path right before the cut is what the whole
//     sequence's search has there, too;
//   * neighbouring pieces OVERLAP by `overlap` variants either side of the cut; each variant is OWNED by exactly one piece;
//   * what a region's search still takes from its past are the two carries of MaxSumBoth (was a baseline variant included
//     before, and the allele id of the last one: MaxSumBoth.java:55,76).  The right piece sees the true carry at the cut when
//     the overlap before the cut holds an included baseline variant and its results there equal the left piece's;
//   * VERIFIED: the right piece must have included a baseline variant before the cut, and from that variant's start to the
//     end of the overlap the two pieces must agree on every variant (status, orientation, weight bits, sync id) and on the
//     sync points, at least one of which must lie before the cut.  Otherwise the sequence is evaluated again in one piece --
//     exactness never depends on where the cuts fell;
//   * sync points are joined in the middle of the gap before each cut (sync ids are position based: unchanged), warnings taken from the
//     owning piece, and the phasing counts (PhasingEvaluator.java:55-142: a state machine over the whole sequence) and the
//     split writers' counters are recomputed from the stitched per-variant results.
//
// Used by vce_submit's several-devices path (vce_capi.cpp) and, with one emulated engine per "device", by the test-only
// emulation (tests/emul/vce_emul.cpp).  Single-pass configurations with inputs sorted by start only; anything else is
// simply not split.
#ifndef VCE_SPLIT_H_
#define VCE_SPLIT_H_

#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <climits>
#include <memory>
#include <vector>

#include "../../include/vce.h"

namespace vce {

struct SplitOptions {
  int64_t minVariants = 100000;   // sequences with fewer variants (both sides) are never split
  int overlap = 2048;             // variants (both sides together) either side of a cut
  int gap = 64;                   // bases without any variant before a cut
  double imbalance = 1.1;         // split only when the longest-processing-time assignment of whole sequences exceeds this
  int piecesPerDevice = 4;        // target piece size = mean device load / this
  // Cutting costs host time (the scan for gaps, result buffers, verification, the phasing recount: measured ~5 ns per variant of
  // a cut sequence on the B200 box's host, profiles/README.md) and saves device time (~10 ns per variant the longest device
  // queue gets shorter).  Unless `force` is set, sequences are cut only when the saving is the larger of the two -- which
  // is the case when a few long sequences dominate a batch, and is NOT the case for 24 chromosomes on 8 devices.
  bool force = false;
  double hostNsPerVariant = 5.0, deviceNsPerVariant = 10.0;
};

template <class T>
struct SplitRaw {               // an uninitialised array (no page is touched before the engine's workers write it)
  std::unique_ptr<T[]> p;
  void alloc(size_t n) { p.reset(new T[n > 0 ? n : 1]); }
  T* data() { return p.get(); }
  const T* data() const { return p.get(); }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

// One piece of one sequence, with its own input view and result buffers.
struct SplitPiece {
  int seq = -1;                   // index of the sequence in the caller's batch
  int lo[2] = {0, 0}, hi[2] = {0, 0};      // variants [lo, hi) per side (0 = calls, 1 = baseline), overlap included
  int own0[2] = {0, 0}, own1[2] = {0, 0};  // owned variants [own0, own1)
  int32_t cutPos = INT_MIN;       // start of the first owned variant (INT_MIN for the first piece)
  int32_t joinPos = INT_MIN;      // sync points below this belong to the previous piece
  vce_sequence view;
  vce_sequence_result res;
  std::vector<int32_t> alleleOff[2], ntOff[2];
  // result buffers: NOT initialised -- the engine writes every entry of every per-variant array (checked with poisoned buffers)
  SplitRaw<uint8_t> status[2], original[2];
  SplitRaw<int8_t> alleleIds[2];
  SplitRaw<double> weight[2];
  SplitRaw<int32_t> syncId[2];
  SplitRaw<int32_t> sync[2];
  std::vector<vce_warning> warn;
};

struct SplitSeq {                 // a sequence that was cut
  int seq = -1;
  std::vector<int> pieces;        // indices into SplitPlan::pieces, in order
  std::vector<int32_t> start[2];  // Allele.getAlleleBounds start per variant (kept for the phasing recount)
};

struct SplitPlan {
  std::vector<std::unique_ptr<SplitPiece>> pieces;
  std::vector<SplitSeq> split;
  std::vector<int32_t> whole;     // sequences evaluated in one piece
};

// ---- bounds (Allele.getAlleleBounds, Allele.java:195-206) and the wide gaps of one sequence
struct SplitBoundary { int32_t c, b, pos; };   // first call / baseline variant after the gap, start of the earlier of the two

// Allele bounds of the variants [lo, hi) of one side; false when a variant has no playable allele or the starts are not ascending
// (such a sequence is not cut).  Chunks of one sequence run as separate tasks.
inline bool splitBoundsChunk(const vce_variants& v, int lo, int hi, int32_t* start, int32_t* end) {
  auto bounds = [&](int i, int32_t& a, int32_t& e) {
    a = INT_MAX; e = INT_MIN;
    for (int q = v.allele_off[i]; q < v.allele_off[i + 1]; ++q) {
      if (!v.allele_null[q]) { a = std::min(a, v.allele_start[q]); e = std::max(e, v.allele_end[q]); }
    }
  };
  int32_t prev = INT_MIN, pe;
  if (lo > 0) bounds(lo - 1, prev, pe);
  for (int i = lo; i < hi; ++i) {
    int32_t a, e;
    bounds(i, a, e);
    if (a == INT_MAX || a < prev) return false;
    prev = a;
    start[i] = a;
    end[i] = e;
  }
  return true;
}

// The wide gaps of one sequence from its variants' bounds.
inline void splitScan(const vce_sequence& s, int gap, const std::vector<int32_t> start[2], const std::vector<int32_t> end[2],
                      std::vector<SplitBoundary>& bounds) {
  const int nc = s.calls.n, nb = s.baseline.n;
  int c = 0, b = 0;
  int64_t maxEnd = INT_MIN;
  bounds.reserve((size_t) (nc + nb) / 2 + 16);
  while (c < nc || b < nb) {
    const bool takeC = b >= nb || (c < nc && start[0][(size_t) c] <= start[1][(size_t) b]);
    const int32_t st = takeC ? start[0][(size_t) c] : start[1][(size_t) b];
    if ((c > 0 || b > 0) && (int64_t) st >= maxEnd + gap) bounds.push_back(SplitBoundary{c, b, st});
    maxEnd = std::max<int64_t>(maxEnd, takeC ? end[0][(size_t) c] : end[1][(size_t) b]);
    maxEnd = std::max<int64_t>(maxEnd, st);
    if (takeC) ++c; else ++b;
  }
}

// Cuts sequence `k` into about `nPieces` pieces; appends them to the plan.  Returns the number of pieces made (0 or 1 = not split).
inline int splitSequence(const vce_sequence* seqs, const vce_sequence_result* results, int k, int nPieces, const SplitOptions& o,
                         std::vector<int32_t> start[2], const std::vector<int32_t> end[2], SplitPlan& plan) {
  const vce_sequence& s = seqs[k];
  const int64_t total = (int64_t) s.calls.n + s.baseline.n;
  if (nPieces < 2 || s.calls.n == 0 || s.baseline.n == 0) return 0;
  SplitSeq ss;
  ss.seq = k;
  std::vector<SplitBoundary> bd;
  splitScan(s, o.gap, start, end, bd);
  if (bd.empty()) return 0;
  ss.start[0].swap(start[0]);
  ss.start[1].swap(start[1]);
  auto count = [&](const SplitBoundary& x) { return (int64_t) x.c + x.b; };
  auto nearest = [&](int64_t target) {   // boundary whose variant count is closest to target
    size_t lo = 0, hi = bd.size();
    while (lo < hi) { const size_t m = (lo + hi) / 2; if (count(bd[m]) < target) lo = m + 1; else hi = m; }
    if (lo == bd.size()) return bd.size() - 1;
    if (lo > 0 && target - count(bd[lo - 1]) < count(bd[lo]) - target) return lo - 1;
    return lo;
  };
  auto atMost = [&](int64_t target, size_t below) -> long {   // last boundary before index `below` with count <= target
    long i = (long) below - 1;
    while (i >= 0 && count(bd[(size_t) i]) > target) --i;
    return i;
  };
  auto atLeast = [&](int64_t target, size_t above) -> long {  // first boundary after index `above` with count >= target
    size_t i = above + 1;
    while (i < bd.size() && count(bd[i]) < target) ++i;
    return i < bd.size() ? (long) i : -1;
  };
  struct Cut { size_t at; long lo, hi; };
  std::vector<Cut> cuts;
  for (int j = 1; j < nPieces; ++j) {
    const size_t at = nearest(total * j / nPieces);
    if (!cuts.empty() && at <= cuts.back().at) continue;
    const long lo = atMost(count(bd[at]) - o.overlap, at), hi = atLeast(count(bd[at]) + o.overlap, at);
    if (lo < 0 || hi < 0) continue;                                   // too close to an end of the sequence
    if (!cuts.empty() && (size_t) lo < (size_t) cuts.back().hi) continue;   // overlaps of neighbouring cuts must not meet
    // a baseline variant must lie in the overlap before the cut (it has to be an INCLUDED one; that is checked afterwards)
    if (bd[at].b - bd[(size_t) lo].b < 1) continue;
    cuts.push_back(Cut{at, lo, hi});
  }
  if (cuts.empty()) return 0;
  const int m = (int) cuts.size() + 1;
  for (int j = 0; j < m; ++j) {
    std::unique_ptr<SplitPiece> p(new SplitPiece());
    p->seq = k;
    const SplitBoundary first = j == 0 ? SplitBoundary{0, 0, INT_MIN} : bd[(size_t) cuts[(size_t) j - 1].lo];
    const SplitBoundary own = j == 0 ? SplitBoundary{0, 0, INT_MIN} : bd[cuts[(size_t) j - 1].at];
    const SplitBoundary ownEnd = j == m - 1 ? SplitBoundary{s.calls.n, s.baseline.n, INT_MAX} : bd[cuts[(size_t) j].at];
    const SplitBoundary last = j == m - 1 ? SplitBoundary{s.calls.n, s.baseline.n, INT_MAX} : bd[(size_t) cuts[(size_t) j].hi];
    p->lo[0] = first.c; p->lo[1] = first.b;
    p->hi[0] = last.c; p->hi[1] = last.b;
    p->own0[0] = own.c; p->own0[1] = own.b;
    p->own1[0] = ownEnd.c; p->own1[1] = ownEnd.b;
    p->cutPos = own.pos;
    p->joinPos = j == 0 ? INT_MIN : own.pos - o.gap / 2;
    // the view: the caller's arrays from variant lo on, slot and base offsets rebased to the piece
    p->view = s;
    // The piece's sequence ends right before the first variant it does not hold (a gap of `gap` bases or more lies before
    // that variant): the search of its last region then finishes there instead of asking for the far end of the chromosome
    // (Path.finished: the last base of the template is read, HaplotypePlayback.java:170-214).  A prefix of a registered
    // template is recognised as registered.
    if (j < m - 1) p->view.template_len = std::min(s.template_len, last.pos);
    memset(&p->res, 0, sizeof(p->res));
    for (int side = 0; side < 2; ++side) {
      const vce_variants& v = side == 0 ? s.calls : s.baseline;
      vce_variants& w = side == 0 ? p->view.calls : p->view.baseline;
      const int lo = p->lo[side], n = p->hi[side] - lo;
      const int32_t slot0 = v.allele_off[lo], slots = v.allele_off[lo + n] - slot0;
      const int32_t nt0 = v.allele_nt_off[slot0];
      p->alleleOff[side].resize((size_t) n + 1);
      for (int i = 0; i <= n; ++i) p->alleleOff[side][(size_t) i] = v.allele_off[lo + i] - slot0;
      p->ntOff[side].resize((size_t) slots + 1);
      for (int q = 0; q <= slots; ++q) p->ntOff[side][(size_t) q] = v.allele_nt_off[slot0 + q] - nt0;
      w.n = n;
      w.id = v.id + lo;
      w.phased = v.phased ? v.phased + lo : nullptr;
      w.status_in = v.status_in ? v.status_in + lo : nullptr;
      w.gt = v.gt ? v.gt + (size_t) lo * (size_t) v.gt_ploidy : nullptr;
      w.allele_off = p->alleleOff[side].data();
      w.n_slots = slots;
      w.allele_start = v.allele_start + slot0;
      w.allele_end = v.allele_end + slot0;
      w.allele_null = v.allele_null + slot0;
      w.allele_nt_off = p->ntOff[side].data();
      w.nt = v.nt + nt0;
      vce_side_result& r = side == 0 ? p->res.calls : p->res.baseline;
      p->status[side].alloc((size_t) n);
      p->original[side].alloc((size_t) n);
      p->alleleIds[side].alloc((size_t) n * VCE_MAX_PLOIDY);
      p->weight[side].alloc((size_t) n);
      p->syncId[side].alloc((size_t) n);
      r.status = p->status[side].data();
      r.original = p->original[side].data();
      r.allele_ids = p->alleleIds[side].data();
      r.weight = p->weight[side].data();
      r.sync_id = p->syncId[side].data();
    }
    const int cap = p->view.calls.n + p->view.baseline.n + 2;
    for (int pass = 0; pass < 2; ++pass) {   // (single-pass configurations only: the second list stays empty)
      const int c = pass == 0 ? cap : 2;
      p->sync[pass].alloc((size_t) c);
      p->res.sync_points[pass] = p->sync[pass].data();
      p->res.sync_cap[pass] = c;
    }
    p->warn.resize((size_t) std::max(1, results[k].warn_cap));
    p->res.warnings = p->warn.data();
    p->res.warn_cap = results[k].warn_cap;
    ss.pieces.push_back((int) plan.pieces.size());
    plan.pieces.push_back(std::move(p));
  }
  plan.split.push_back(std::move(ss));
  return m;
}

// Longest-processing-time greedy: item weights -> device of every item; returns max load / mean load.
inline double splitAssign(const std::vector<int64_t>& weight, int nDev, std::vector<int>& device, int64_t* maxLoad = nullptr) {
  std::vector<int> order(weight.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int) i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight[(size_t) a] > weight[(size_t) b]; });
  std::vector<int64_t> load((size_t) nDev, 0);
  device.assign(weight.size(), 0);
  for (int i : order) {
    size_t best = 0;
    for (size_t d = 1; d < (size_t) nDev; ++d) if (load[d] < load[best]) best = d;
    device[(size_t) i] = (int) best;
    load[best] += weight[(size_t) i] + 1;
  }
  int64_t sum = 0, mx = 0;
  for (int64_t l : load) { sum += l; mx = std::max(mx, l); }
  if (maxLoad) *maxLoad = mx;
  return sum > 0 ? (double) mx * nDev / (double) sum : 1.0;
}

// Runs fn(0) .. fn(n - 1), possibly concurrently (the caller's worker pool; the emulation runs them in turn).
struct SplitSerial {
  template <class F> void operator()(int n, F fn) const { for (int i = 0; i < n; ++i) fn(i); }
};

// Decides which sequences to cut for `nDev` devices and cuts them (every sequence's scan is a task of `par`).
template <class Par>
inline void splitPlan(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs, const vce_sequence_result* results, int nDev,
                      const SplitOptions& o, SplitPlan& plan, const Par& par) {
  std::vector<int64_t> w((size_t) nSeq);
  int64_t total = 0;
  for (int32_t k = 0; k < nSeq; ++k) { w[(size_t) k] = (int64_t) seqs[k].calls.n + seqs[k].baseline.n; total += w[(size_t) k]; }
  std::vector<int> dev;
  int64_t maxWhole = 0;
  const bool wanted = nDev > 1 && cfg.n_passes == 1 && cfg.flag_alternates == 0 && splitAssign(w, nDev, dev, &maxWhole) > o.imbalance;
  const int64_t target = std::max<int64_t>(o.minVariants / 2, total / std::max(1, nDev) / std::max(1, o.piecesPerDevice));
  std::vector<int32_t> cand;
  int64_t cutVariants = 0;
  for (int32_t k = 0; k < nSeq; ++k) {
    if (wanted && w[(size_t) k] >= o.minVariants && w[(size_t) k] > target + target / 2) { cand.push_back(k); cutVariants += w[(size_t) k]; }
  }
  if (!o.force && !cand.empty()) {
    const double saved = ((double) maxWhole - 1.03 * (double) total / nDev) * o.deviceNsPerVariant;
    if (saved <= (double) cutVariants * o.hostNsPerVariant) cand.clear();
  }
  std::vector<SplitPlan> local(cand.size());
  std::vector<int> made(cand.size(), 0);
  // 1. allele bounds, in chunks of 64 Ki variants (a long chromosome is many tasks)
  struct Bounds { std::vector<int32_t> start[2], end[2]; };
  std::vector<Bounds> bounds(cand.size());
  struct Chunk { int i, side, lo, hi; };
  std::vector<Chunk> chunks;
  for (size_t i = 0; i < cand.size(); ++i) {
    for (int side = 0; side < 2; ++side) {
      const int n = side == 0 ? seqs[cand[i]].calls.n : seqs[cand[i]].baseline.n;
      bounds[i].start[side].resize((size_t) n);
      bounds[i].end[side].resize((size_t) n);
      for (int lo = 0; lo < n; lo += 1 << 16) chunks.push_back(Chunk{(int) i, side, lo, std::min(n, lo + (1 << 16))});
    }
  }
  std::vector<uint8_t> bad(cand.size(), 0);
  par((int) chunks.size(), [&](int t) {
    const Chunk& ch = chunks[(size_t) t];
    const vce_sequence& sq = seqs[cand[(size_t) ch.i]];
    Bounds& bo = bounds[(size_t) ch.i];
    if (!splitBoundsChunk(ch.side == 0 ? sq.calls : sq.baseline, ch.lo, ch.hi, bo.start[ch.side].data(), bo.end[ch.side].data())) bad[(size_t) ch.i] = 1;
  });
  // 2. gaps, cuts, views and result buffers, one task per sequence
  par((int) cand.size(), [&](int i) {
    const int32_t k = cand[(size_t) i];
    if (bad[(size_t) i]) return;
    made[(size_t) i] = splitSequence(seqs, results, k, (int) ((w[(size_t) k] + target - 1) / target), o, bounds[(size_t) i].start,
                                     bounds[(size_t) i].end, local[(size_t) i]);
  });
  std::vector<uint8_t> cut((size_t) nSeq, 0);
  for (size_t i = 0; i < cand.size(); ++i) {
    if (made[i] < 2) continue;
    cut[(size_t) cand[i]] = 1;
    const int base = (int) plan.pieces.size();
    for (auto& p : local[i].pieces) plan.pieces.push_back(std::move(p));
    for (SplitSeq& ss : local[i].split) {
      for (int& q : ss.pieces) q += base;
      plan.split.push_back(std::move(ss));
    }
  }
  for (int32_t k = 0; k < nSeq; ++k) if (!cut[(size_t) k]) plan.whole.push_back(k);
}
inline void splitPlan(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs, const vce_sequence_result* results, int nDev,
                      const SplitOptions& o, SplitPlan& plan) {
  splitPlan(cfg, nSeq, seqs, results, nDev, o, plan, SplitSerial());
}

// ---- stitching
struct SplitVSum { int32_t start; bool phased, included, phase; };

// PhasingEvaluator.countMisphasings (PhasingEvaluator.java:55-142) over per-variant RESULTS of a single-pass evaluation:
// included = GT_MATCH (orientation's isOriginal = its phase), excluded = evaluated and not matched, skipped variants take no
// part (CallIterator :160-195 merges included and excluded by start, included first on ties).
struct SplitCallIter {
  const uint8_t* status; const uint8_t* original; const uint8_t* phased; const int32_t* start;
  int n, i = 0, e = 0;
  static bool inc(uint8_t st) { return (st & VCE_STATUS_GT_MATCH) != 0; }
  static bool exc(uint8_t st) { return !(st & VCE_STATUS_GT_MATCH) && !(st & VCE_STATUS_SKIPPED) && (st & (VCE_STATUS_NO_MATCH | VCE_STATUS_ALLELE_MATCH)); }
  SplitCallIter(const uint8_t* st, const uint8_t* orig, const uint8_t* ph, const int32_t* s, int n_) : status(st), original(orig), phased(ph), start(s), n(n_) {
    while (i < n && !inc(status[i])) ++i;
    while (e < n && !exc(status[e])) ++e;
  }
  bool hasNext() const { return i < n || e < n; }
  SplitVSum next() {
    const bool haveI = i < n, haveE = e < n;
    if (!haveI || (haveE && start[e] < start[i])) {
      const int li = e;
      do { ++e; } while (e < n && !exc(status[e]));
      return SplitVSum{start[li], phased && phased[li], false, false};
    }
    const int li = i;
    do { ++i; } while (i < n && !inc(status[i]));
    return SplitVSum{start[li], phased && phased[li], true, original[li] != 0};
  }
};

inline void splitPhasing(const vce_sequence& s, const vce_sequence_result& r, const std::vector<int32_t> start[2], int32_t out[3]) {
  SplitCallIter baseline(r.baseline.status, r.baseline.original, s.baseline.phased, start[1].data(), s.baseline.n);
  SplitCallIter calls(r.calls.status, r.calls.original, s.calls.phased, start[0].data(), s.calls.n);
  int mis = 0, unph = 0, correct = 0;
  bool baseIsPhased = false, basePhase = false, callIsPhased = false, callPhase = false;
  bool haveCall = false, haveBase = false;
  SplitVSum call{0, false, false, false}, base{0, false, false, false};
  for (int q = 0; q < r.n_sync[0]; ++q) {
    const int point = r.sync_points[0][q];
    int nb = 0;
    bool inPhase = true, phase = false;
    do {
      if (haveBase && base.start < point) {
        if (nb == 0) { inPhase = base.phased; phase = base.phase; }
        else if (!base.phased || base.phase != phase) inPhase = false;
        ++nb;
        haveBase = baseline.hasNext();
        if (haveBase) base = baseline.next();
      }
      if (!haveBase) {
        haveBase = baseline.hasNext();
        if (haveBase) base = baseline.next();
      }
    } while (haveBase && base.start < point);
    bool transition = false;
    if (inPhase) {
      if (!baseIsPhased) callIsPhased = false;
      if (nb > 0) {
        transition = basePhase != phase;
        baseIsPhased = true;
        basePhase = phase;
      }
    }
    do {
      if (haveCall) {
        const SplitVSum& c = call;
        if (!inPhase) {
          if (c.phased) ++unph;
        } else {
          if (!c.phased) {
            callIsPhased = false;
          } else if (!callIsPhased) {
            callIsPhased = true;
            callPhase = c.phase;
          } else {
            const bool callTransition = c.phase != callPhase;
            if (c.included && !(callTransition == transition)) { ++mis; callPhase = c.phase; }
            else if (c.included) { ++correct; callPhase = c.phase; }
          }
          transition = false;
        }
      }
      haveCall = calls.hasNext();
      if (haveCall) call = calls.next();
    } while (haveCall && call.start < point);
    if (!inPhase) {
      baseIsPhased = false;
      callIsPhased = false;
    }
  }
  out[0] = mis; out[1] = correct; out[2] = unph;
}

// The split writers' counters of one sequence (SplitEvalSynchronizer.java:112-156), as Engine::summaryFromResults counts them.
inline void splitSummary(const vce_sequence& sq, const vce_sequence_result& r, int64_t cn[8]) {
  for (int q = 0; q < 8; ++q) cn[q] = 0;
  if (r.error != VCE_OK) return;
  for (int i = 0; i < sq.calls.n; ++i) {
    const int st = r.calls.status[i];
    if (st & VCE_STATUS_SKIPPED) ++cn[4];
    if (st & VCE_STATUS_GT_MATCH) { if (r.calls.weight[i] > 0) ++cn[2]; }
    else if (st & VCE_STATUS_ALLELE_MATCH) { if (r.calls.weight[i] > 0) ++cn[3]; }
    else if ((st & VCE_STATUS_NO_MATCH) && !(st & VCE_STATUS_OUTSIDE_EVAL)) ++cn[3];
  }
  for (int i = 0; i < sq.baseline.n; ++i) {
    const int st = r.baseline.status[i];
    if (st & VCE_STATUS_SKIPPED) ++cn[4];
    if (st & VCE_STATUS_OUTSIDE_EVAL) continue;
    if (st & VCE_STATUS_GT_MATCH) ++cn[0];
    else if (st & (VCE_STATUS_ALLELE_MATCH | VCE_STATUS_NO_MATCH)) ++cn[1];
  }
  cn[5] = r.phasing[0]; cn[6] = r.phasing[1]; cn[7] = r.phasing[2];
}

// Stitches the pieces of one sequence into the caller's result.  Returns false when the verification failed (or a piece
// reported an error): the caller evaluates the sequence again in one piece.
inline bool splitMerge(const SplitPlan& plan, const SplitSeq& ss, const vce_sequence* seqs, vce_sequence_result* results) {
  const vce_sequence& s = seqs[ss.seq];
  vce_sequence_result& out = results[ss.seq];
  const int m = (int) ss.pieces.size();
  for (int j = 0; j < m; ++j) if (plan.pieces[(size_t) ss.pieces[(size_t) j]]->res.error != VCE_OK) return false;
  // 1. verification of every overlap
  for (int j = 1; j < m; ++j) {
    const SplitPiece& a = *plan.pieces[(size_t) ss.pieces[(size_t) j - 1]];
    const SplitPiece& b = *plan.pieces[(size_t) ss.pieces[(size_t) j]];
    // The right piece starts with an empty past; from its first included baseline variant on its carries are its own.  That
    // variant must lie before the cut, and from its start on the two pieces must agree on everything.
    int f = -1;
    for (int v = b.lo[1]; v < b.own0[1] && f < 0; ++v) if (b.status[1][(size_t) (v - b.lo[1])] & VCE_STATUS_GT_MATCH) f = v;
    if (f < 0) return false;
    const int32_t fromPos = ss.start[1][(size_t) f];
    for (int side = 0; side < 2; ++side) {
      for (int v = b.lo[side]; v < a.hi[side]; ++v) {
        if (ss.start[side][(size_t) v] < fromPos) continue;
        const size_t ia = (size_t) (v - a.lo[side]), ib = (size_t) (v - b.lo[side]);
        const uint8_t st = a.status[side][ia];
        if (st != b.status[side][ib]) return false;
        if (st & (VCE_STATUS_GT_MATCH | VCE_STATUS_ALLELE_MATCH)) {
          if (a.original[side][ia] != b.original[side][ib]) return false;
          if (memcmp(&a.alleleIds[side][ia * VCE_MAX_PLOIDY], &b.alleleIds[side][ib * VCE_MAX_PLOIDY], VCE_MAX_PLOIDY) != 0) return false;
        }
        if (memcmp(&a.weight[side][ia], &b.weight[side][ib], sizeof(double)) != 0) return false;
        if (a.syncId[side][ia] != b.syncId[side][ib]) return false;   // the same floor sync point in both pieces
      }
    }
    // The sync points from that variant on, up to the last variant both pieces hold, must be the same -- and one of them must
    // lie before the cut: there both searches stand in sync at the same position with the same carries, whatever the right
    // piece assumed at its start (in a long tandem repeat the whole sequence's search may NOT be in sync where the piece began).
    const int32_t from = fromPos;
    int32_t to = INT_MAX;
    for (int side = 0; side < 2; ++side) if (a.hi[side] > b.lo[side]) to = std::min(to, ss.start[side][(size_t) a.hi[side] - 1]);
    size_t qa = 0, qb = 0;
    const size_t na = (size_t) a.res.n_sync[0], nb = (size_t) b.res.n_sync[0];
    while (qa < na && a.sync[0][qa] < from) ++qa;
    while (qb < nb && b.sync[0][qb] < from) ++qb;
    bool beforeCut = false;
    while (qa < na && qb < nb && a.sync[0][qa] < to && b.sync[0][qb] < to) {
      if (a.sync[0][qa] != b.sync[0][qb]) return false;
      if (a.sync[0][qa] < b.joinPos) beforeCut = true;
      ++qa; ++qb;
    }
    if ((qa < na && a.sync[0][qa] < to) != (qb < nb && b.sync[0][qb] < to)) return false;
    if (!beforeCut) return false;
  }
  // 2. sync points: piece j contributes the points in [joinPos(j), joinPos(j + 1))
  int nSync = 0;
  for (int j = 0; j < m; ++j) {
    const SplitPiece& p = *plan.pieces[(size_t) ss.pieces[(size_t) j]];
    const int32_t from = p.joinPos;
    const int32_t to = j + 1 < m ? plan.pieces[(size_t) ss.pieces[(size_t) j + 1]]->joinPos : INT_MAX;
    int skip = 0;
    const int n = p.res.n_sync[0];
    while (skip < n && p.sync[0][(size_t) skip] < from) ++skip;
    for (int q = skip; q < n && (p.sync[0][(size_t) q] < to || j + 1 == m); ++q) {
      if (nSync < out.sync_cap[0]) out.sync_points[0][nSync] = p.sync[0][(size_t) q];
      ++nSync;
    }
  }
  if (nSync > out.sync_cap[0]) return false;
  out.n_sync[0] = nSync;
  out.n_sync[1] = 0;
  // 3. per-variant results of the owned ranges
  out.n_regions = 0;
  out.n_iterations = 0;
  out.n_warnings = 0;
  out.error = VCE_OK;
  for (int j = 0; j < m; ++j) {
    const SplitPiece& p = *plan.pieces[(size_t) ss.pieces[(size_t) j]];
    for (int side = 0; side < 2; ++side) {
      vce_side_result& r = side == 0 ? out.calls : out.baseline;
      const int v0 = p.own0[side], n = p.own1[side] - v0;
      if (n <= 0) continue;
      const size_t off = (size_t) (v0 - p.lo[side]);
      memcpy(r.status + v0, p.status[side].data() + off, (size_t) n);
      memcpy(r.original + v0, p.original[side].data() + off, (size_t) n);
      memcpy(r.allele_ids + (size_t) v0 * VCE_MAX_PLOIDY, p.alleleIds[side].data() + off * VCE_MAX_PLOIDY, (size_t) n * VCE_MAX_PLOIDY);
      memcpy(r.weight + v0, p.weight[side].data() + off, (size_t) n * sizeof(double));
      // (the SYNC value is the floor sync point's POSITION + 2, InterleavingEvalSynchronizer.floorSyncPos :71-83: nothing to shift)
      if (r.sync_id) memcpy(r.sync_id + v0, p.syncId[side].data() + off, (size_t) n * sizeof(int32_t));
    }
    out.n_regions += p.res.n_regions;
    out.n_iterations += p.res.n_iterations;
    // warnings of the regions this piece owns (PathFinder.java:162-169: start1 = lastSyncPos + 1)
    const int32_t wFrom = p.joinPos;
    const int32_t wTo = j + 1 < m ? plan.pieces[(size_t) ss.pieces[(size_t) j + 1]]->joinPos : INT_MAX;
    const int nw = std::min(p.res.n_warnings, p.res.warn_cap);
    if (p.res.n_warnings > p.res.warn_cap) return false;   // more warnings than the caller has room for: the one-piece run reports them
    for (int q = 0; q < nw; ++q) {
      const vce_warning& w = p.warn[(size_t) q];
      if (w.start1 - 1 < wFrom || w.start1 - 1 >= wTo) continue;
      if (out.n_warnings < out.warn_cap) out.warnings[out.n_warnings] = w;
      ++out.n_warnings;
    }
  }
  // 4. the whole-sequence state machine over the stitched results
  splitPhasing(s, out, ss.start, out.phasing);
  return true;
}

// Stitches every cut sequence (a task of `par` each); `summary` ([nSeq * 8]) receives the writers' counters of the stitched
// ones, `redo` the sequences whose cuts could not be verified.
template <class Par>
inline void splitMergeAll(const SplitPlan& plan, const vce_sequence* seqs, vce_sequence_result* results, int64_t* summary,
                          std::vector<int32_t>& redo, const Par& par) {
  std::vector<uint8_t> ok(plan.split.size(), 0);
  par((int) plan.split.size(), [&](int i) {
    const SplitSeq& ss = plan.split[(size_t) i];
    if (splitMerge(plan, ss, seqs, results)) {
      splitSummary(seqs[ss.seq], results[ss.seq], summary + (size_t) ss.seq * 8);
      ok[(size_t) i] = 1;
    }
  });
  for (size_t i = 0; i < plan.split.size(); ++i) if (!ok[i]) redo.push_back(plan.split[i].seq);
}

}  // namespace vce
#endif
