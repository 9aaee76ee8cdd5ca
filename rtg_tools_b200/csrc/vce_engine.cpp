// vce_engine.cpp -- see vce_engine.h
#include "vce_engine.h"

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "vce_orient.h"

#include <chrono>

namespace vce {

namespace {
// VCE_TIMING=1: stage durations on stderr
struct StageTimer {
  bool on;
  std::chrono::steady_clock::time_point t0;
  StageTimer() : on(std::getenv("VCE_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[vce] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

const int kDummyPrefix = 0;  // "some baseline variant was included earlier, allele id unknown": the common case

// Upper bound on the number of orientations any variant of a pass can have (Orientor.java), used to size child
// slots and decision fields before K1 has run.
int orientBound(int kind, int H, int gtPloidy, int maxSlots) {
  const int nA = std::max(1, maxSlots - 1);  // numAlleles()
  static const int fact[5] = {1, 1, 2, 6, 24};
  switch (kind) {
    case VCE_ORIENT_UNPHASED:
    case VCE_ORIENT_PHASED:
    case VCE_ORIENT_PHASED_INVERT: return H == 2 ? 2 : fact[std::min(4, std::max(gtPloidy, H))];
    case VCE_ORIENT_SQUASH_GT: return std::max(1, gtPloidy);
    case VCE_ORIENT_ALLELE_GT: return 6;
    case VCE_ORIENT_HAPLOID_POP: return std::max(1, nA - 1);
    case VCE_ORIENT_DIPLOID_POP: return std::max(1, (nA - 1) * (2 * nA + 1));
    case VCE_ORIENT_ALT: {
      long b = nA - 1;
      for (int h = 1; h < H; ++h) b *= (nA + 1);
      b *= fact[std::min(4, H)];
      return (int) std::min<long>(std::max<long>(1, b), 1 << 20);
    }
  }
  return 1;
}

struct RowPick {  // selects one orientation row of a variant (host side of K1, for the output arrays)
  int want, at = 0;
  int8_t ids[4] = {-2, -2, -2, -2};
  bool original = false;
  void operator()(const int* p, int n, bool orig) {
    if (at == want) {
      for (int h = 0; h < 4; ++h) ids[h] = h < n ? (int8_t) p[h] : (int8_t) -2;
      original = orig;
    }
    ++at;
  }
};
}  // namespace

void packTemplate(const uint8_t* bases, int32_t len, WorkerPool* pool, std::vector<uint32_t>& packed, std::vector<uint32_t>& nmask) {
  const size_t words = ((size_t) len + 15) / 16;
  const size_t mwords = (words + 31) / 32;
  packed.assign(words, 0);
  nmask.assign(mwords, 0);
  const size_t block = 4096;   // mask words per task (2 Mi bases)
  auto work = [&](int t, int) {
    const size_t m0 = (size_t) t * block, m1 = std::min(mwords, m0 + block);
    for (size_t m = m0; m < m1; ++m) {
      uint32_t mask = 0;
      for (int gi = 0; gi < 32; ++gi) {
        const size_t w = m * 32 + (size_t) gi;
        if (w >= words) break;
        const size_t p0 = w * 16;
        uint32_t v = 0;
        bool other = false;
        for (int q = 0; q < 16 && p0 + q < (size_t) len; ++q) {
          const unsigned b = bases[p0 + q];
          if (b < 1 || b > 4) other = true;
          v |= ((b - 1u) & 3u) << (2 * q);
        }
        packed[w] = v;
        if (other) mask |= 1u << gi;
      }
      nmask[m] = mask;
    }
  };
  const int nTasks = (int) ((mwords + block - 1) / block);
  if (pool) pool->parallelFor(nTasks, work);
  else for (int t = 0; t < nTasks; ++t) work(t, 0);
}

Engine::Engine(Backend* b, WorkerPool* pool, const EngineOptions& o) : be_(b), pool_(pool), opt_(o) {
  classCount_.resize(pool_->threads());
  classKeys_.resize(pool_->threads());
  classOrder_.resize(pool_->threads());
  scratch_ = new std::vector<uint8_t>[pool_->threads()][2];
  scratchOff_ = new std::vector<uint32_t>[pool_->threads()][2];
}

Engine::~Engine() {
  delete[] scratch_;
  delete[] scratchOff_;
  for (HostBlock& hb : arena_) be_->hostFree(hb.p);
}

uint8_t* Engine::arenaAlloc(size_t bytes) {
  bytes = (bytes + 63) & ~(size_t) 63;
  std::lock_guard<std::mutex> lk(arenaMu_);
  for (HostBlock& hb : arena_) {
    if (hb.cap - hb.used >= bytes) {
      uint8_t* p = hb.p + hb.used;
      hb.used += bytes;
      return p;
    }
  }
  const size_t cap = std::max(bytes, (size_t) 64 << 20);
  uint8_t* p = static_cast<uint8_t*>(be_->hostAlloc(cap));
  if (!p) return nullptr;
  arena_.push_back(HostBlock{p, cap, bytes});
  return p;
}

void Engine::arenaReset() {
  std::lock_guard<std::mutex> lk(arenaMu_);
  for (HostBlock& hb : arena_) hb.used = 0;
  arenaMarks_.clear();
}
void Engine::arenaMark() {
  std::lock_guard<std::mutex> lk(arenaMu_);
  arenaMarks_.clear();
  for (HostBlock& hb : arena_) arenaMarks_.push_back(hb.used);
}
void Engine::arenaRewind() {
  std::lock_guard<std::mutex> lk(arenaMu_);
  for (size_t i = 0; i < arena_.size(); ++i) arena_[i].used = i < arenaMarks_.size() ? arenaMarks_[i] : 0;
}

// ------------------------------------------------------------------------------------------- pass set-up
namespace {
template <class V> void ensureSize(V& v, size_t n) { if (v.size() < n) v.resize(n); }
}  // namespace

void Engine::resetPass(PassCtx& pc) {  // forget the contents, keep the memory
  pc.chunks.clear();
  pc.wideSync.clear();
  pc.warns.clear();
  pc.rare.clear();
  pc.undo.clear();
  pc.launched = false;
  pc.gtPloidy[0] = pc.gtPloidy[1] = 0;
  for (int side = 0; side < 2; ++side) {
    pc.ss[side].clear();
    for (auto& v : pc.idxStore[side]) v.clear();
  }
}

int Engine::setupPass0() {
  PassCtx& pc = pass_[0];
  resetPass(pc);
  pc.pass = 0;
  pc.H = cfg_.ploidy[0];
  pc.kind[0] = cfg_.calls_orientor[0];
  pc.kind[1] = cfg_.baseline_orientor[0];
  pc.shortCircuit.assign(nSeq_, 0);
  for (int side = 0; side < 2; ++side) {
    pc.ss[side].assign(nSeq_, SideSeq());
    ensureSize(pc.idxStore[side], (size_t) nSeq_);
  }
  int32_t tot[2] = {0, 0};
  bool havePloidy[2] = {false, false};
  for (int k = 0; k < nSeq_; ++k) {
    const bool sc = seqs_[k].baseline.n == 0 || seqs_[k].calls.n == 0;  // SequenceEvaluator.java:94-97
    pc.shortCircuit[k] = sc ? 1 : 0;
    for (int side = 0; side < 2; ++side) {
      const vce_variants& v = side == 0 ? seqs_[k].calls : seqs_[k].baseline;
      if (v.n > 0) {
        // one genotype width per side and batch (the orientation tables and the GT padding are per pass): a mixed
        // batch, e.g. haploid chrX records beside diploid autosomes, is the caller's to split (VariantFactory.java:170-174
        // already widens haploid GTs in diploid mode)
        if (v.gt_ploidy < 0 || v.gt_ploidy > VCE_MAX_PLOIDY) { err_ = "gt_ploidy out of range"; return VCE_ERR_BAD_ARGUMENT; }
        if (havePloidy[side] && pc.gtPloidy[side] != (int) v.gt_ploidy) {
          err_ = "gt_ploidy differs between the sequences of one batch";
          return VCE_ERR_BAD_ARGUMENT;
        }
        havePloidy[side] = true;
        pc.gtPloidy[side] = (int) v.gt_ploidy;
      }
      SideSeq& s = pc.ss[side][k];
      s.in = &v;
      s.first = tot[side];
      s.n = sc ? 0 : v.n;
      s.idx = nullptr;
      tot[side] += s.n;
    }
  }
  for (int side = 0; side < 2; ++side) {
    // every variant belongs to exactly one settled region, whose results overwrite these: no clearing needed
    ensureSize(pc.vStart[side], (size_t) tot[side]);
    ensureSize(pc.vEnd[side], (size_t) tot[side]);
    ensureSize(pc.decision[side], (size_t) tot[side]);
  }
  ensureSize(pc.weight, (size_t) tot[0]);
  computeBounds(pc);
  return VCE_OK;
}

// Allele.getAlleleBounds (Allele.java:195-206) for every variant, and the stable (start, end) order PathFinder
// establishes (PathFinder.java:122-126) when the input is not already in that order.
void Engine::computeBounds(PassCtx& pc) {
  struct Task { int side, seq, lo, hi; };
  std::vector<Task> tasks;
  const int block = 1 << 16;
  for (int side = 0; side < 2; ++side)
    for (int k = 0; k < nSeq_; ++k)
      for (int lo = 0; lo < pc.ss[side][k].n; lo += block) tasks.push_back(Task{side, k, lo, std::min(pc.ss[side][k].n, lo + block)});
  std::vector<std::atomic<int>> unsorted(2 * (size_t) std::max(1, nSeq_));
  for (auto& u : unsorted) u.store(0);
  pool_->parallelFor((int) tasks.size(), [&](int t, int) {
    const Task& tk = tasks[t];
    const SideSeq& s = pc.ss[tk.side][tk.seq];
    const vce_variants& v = *s.in;
    int32_t* vs = pc.vStart[tk.side].data() + s.first;
    int32_t* ve = pc.vEnd[tk.side].data() + s.first;
    auto bounds = [&](int i, int& st, int& en) {
      int a = INT_MAX, b = INT_MIN;
      for (int q = v.allele_off[i]; q < v.allele_off[i + 1]; ++q) {
        if (!v.allele_null[q]) {
          a = std::min(a, v.allele_start[q]);
          b = std::max(b, v.allele_end[q]);
        }
      }
      st = a;
      en = b;
    };
    int ps = INT_MIN, pe = INT_MIN;
    if (tk.lo > 0) bounds(tk.lo - 1, ps, pe);
    bool bad = false;
    for (int i = tk.lo; i < tk.hi; ++i) {
      int st, en;
      bounds(i, st, en);
      vs[i] = st;
      ve[i] = en;
      if (st < ps || (st == ps && en < pe)) bad = true;
      ps = st;
      pe = en;
    }
    if (bad) unsorted[(size_t) tk.side * nSeq_ + tk.seq].store(1);
  });
  for (int side = 0; side < 2; ++side) {
    for (int k = 0; k < nSeq_; ++k) {
      if (!unsorted[(size_t) side * nSeq_ + k].load()) continue;
      SideSeq& s = pc.ss[side][k];
      std::vector<int32_t>& order = pc.idxStore[side][k];
      order.resize(s.n);
      for (int i = 0; i < s.n; ++i) order[i] = i;
      int32_t* vs = pc.vStart[side].data() + s.first;
      int32_t* ve = pc.vEnd[side].data() + s.first;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (vs[a] != vs[b]) return vs[a] < vs[b];
        return ve[a] < ve[b];
      });
      std::vector<int32_t> ts(s.n), te(s.n);
      for (int i = 0; i < s.n; ++i) { ts[i] = vs[order[i]]; te[i] = ve[order[i]]; }
      std::copy(ts.begin(), ts.end(), vs);
      std::copy(te.begin(), te.end(), ve);
      s.idx = order.data();
    }
  }
}

int Engine::setupPass1() {
  const PassCtx& p0 = pass_[0];
  PassCtx& pc = pass_[1];
  resetPass(pc);
  pc.pass = 1;
  pc.H = cfg_.ploidy[1];
  pc.kind[0] = cfg_.calls_orientor[1];
  pc.kind[1] = cfg_.baseline_orientor[1];
  pc.shortCircuit = p0.shortCircuit;
  pc.gtPloidy[0] = p0.gtPloidy[0];
  pc.gtPloidy[1] = p0.gtPloidy[1];
  for (int side = 0; side < 2; ++side) {
    pc.ss[side].assign(nSeq_, SideSeq());
    ensureSize(pc.idxStore[side], (size_t) nSeq_);
    // Path.getCalledExcluded / getBaselineExcluded of the first pass, per sequence
    pool_->parallelFor(nSeq_, [&](int k, int) {
      const SideSeq& a = p0.ss[side][k];
      std::vector<int32_t>& ix = pc.idxStore[side][k];
      ix.clear();
      const uint8_t* dec = p0.decision[side].data() + a.first;
      for (int i = 0; i < a.n; ++i) if (dec[i] == 1) ix.push_back(i);  // position in pass 0 for now
    });
    int32_t tot = 0;
    for (int k = 0; k < nSeq_; ++k) {
      SideSeq& s = pc.ss[side][k];
      s.in = p0.ss[side][k].in;
      s.first = tot;
      s.n = (int32_t) pc.idxStore[side][k].size();
      tot += s.n;
    }
    ensureSize(pc.vStart[side], (size_t) tot);
    ensureSize(pc.vEnd[side], (size_t) tot);
    ensureSize(pc.decision[side], (size_t) tot);
    if (side == 0) ensureSize(pc.weight, (size_t) tot);
    pool_->parallelFor(nSeq_, [&](int k, int) {
      const SideSeq& a = p0.ss[side][k];
      SideSeq& s = pc.ss[side][k];
      std::vector<int32_t>& ix = pc.idxStore[side][k];
      const int32_t* as = p0.vStart[side].data() + a.first;
      const int32_t* ae = p0.vEnd[side].data() + a.first;
      int32_t* vs = pc.vStart[side].data() + s.first;
      int32_t* ve = pc.vEnd[side].data() + s.first;
      for (int i = 0; i < s.n; ++i) {
        const int p = ix[i];
        vs[i] = as[p];
        ve[i] = ae[p];
        ix[i] = a.idx ? a.idx[p] : p;  // now the input index
      }
      s.idx = ix.data();
    });
  }
  return VCE_OK;
}

// ------------------------------------------------------------------------------------------- regions
// Splits calls [i, ie) and baseline [j, je) (pass-wide indices, merged by start, calls first on ties) into runs
// separated by a gap; `maxEnd` carries the largest variant end seen before the slice (LONG_MIN = nothing).
// *continues is set when the first run of the slice is not separated from what came before.
// `near` (optional): one flag per region, set when the next region of the slice starts less than denseGap bases after this
// region's largest variant end (the last region of a slice has none).
void Engine::splitSlice(const PassCtx& pc, int i, int ie, int j, int je, long maxEnd, std::vector<RegRange>& out, bool* continues,
                        std::vector<uint8_t>* near) const {
  const int32_t* cs = pc.vStart[0].data();
  const int32_t* ce = pc.vEnd[0].data();
  const int32_t* bs = pc.vStart[1].data();
  const int32_t* bee = pc.vEnd[1].data();
  const long gap = std::max(1, opt_.gap);
  bool firstRun = true;
  do {
    RegRange g{i, 0, j, 0};
    bool any = false;
    uint8_t nearNext = 0;
    if (!firstRun) maxEnd = LONG_MIN;  // a new run starts after a gap
    while (i < ie || j < je) {
      const bool takeC = j >= je || (i < ie && cs[i] <= bs[j]);
      const int s = takeC ? cs[i] : bs[j];
      const int e = takeC ? ce[i] : bee[j];
      if (maxEnd != LONG_MIN && (long) s >= maxEnd + gap) {
        if (any) {
          nearNext = (long) s - maxEnd < (long) opt_.denseGap ? 1 : 0;
          break;
        }
        maxEnd = LONG_MIN;  // the slice itself starts with a gap: its first run is a fresh region
      } else if (!any && firstRun && maxEnd != LONG_MIN && continues) {
        *continues = true;
      }
      any = true;
      maxEnd = std::max(maxEnd, (long) e);
      if (takeC) ++i; else ++j;
    }
    g.cCount = i - g.cFirst;
    g.bCount = j - g.bFirst;
    out.push_back(g);
    if (near) near->push_back(nearNext);
    firstRun = false;
  } while (i < ie || j < je);
}

void Engine::splitAll(PassCtx& pc) {
  StageTimer ts;
  ensureSize(pc.rr, (size_t) nSeq_);
  ensureSize(pc.rs, (size_t) nSeq_);
  ensureSize(pc.rcls, (size_t) nSeq_);
  // (the segments' output vectors persist across calls and keep their capacity: no allocation, no page faults)
  struct Seg { int seq, i, ie, j, je; long maxEndLocal, carry; bool continues; std::vector<RegRange>* outp; };
  std::vector<Seg> segs;
  std::vector<int> segFirst(nSeq_ + 1, 0);
  const int per = std::max(1, opt_.splitSegment);
  for (int k = 0; k < nSeq_; ++k) {
    segFirst[k] = (int) segs.size();
    if (pc.shortCircuit[k]) continue;
    const SideSeq& C = pc.ss[0][k];
    const SideSeq& B = pc.ss[1][k];
    const int total = C.n + B.n;
    const int nSeg = std::max(1, std::min(4096, total / per));
    // cut by template position: quantiles of the larger side
    const int side = C.n >= B.n ? 0 : 1;
    const SideSeq& big = side == 0 ? C : B;
    const int32_t* cs = pc.vStart[0].data() + C.first;
    const int32_t* bs = pc.vStart[1].data() + B.first;
    int pi = 0, pj = 0;
    for (int sgi = 1; sgi <= nSeg; ++sgi) {
      int ni = C.n, nj = B.n;
      if (sgi < nSeg) {
        const int32_t cut = pc.vStart[side][big.first + (int) ((int64_t) big.n * sgi / nSeg)];
        ni = (int) (std::lower_bound(cs, cs + C.n, cut) - cs);
        nj = (int) (std::lower_bound(bs, bs + B.n, cut) - bs);
        ni = std::max(ni, pi);
        nj = std::max(nj, pj);
      }
      if (sgi == nSeg || ni > pi || nj > pj) {
        Seg sg{k, C.first + pi, C.first + ni, B.first + pj, B.first + nj, LONG_MIN, LONG_MIN, false, nullptr};
        segs.push_back(sg);
        pi = ni;
        pj = nj;
      }
    }
  }
  segFirst[nSeq_] = (int) segs.size();
  if (splitOut_.size() < segs.size()) splitOut_.resize(segs.size());
  for (size_t t = 0; t < segs.size(); ++t) { segs[t].outp = &splitOut_[t].v; splitOut_[t].v.clear(); }
  ts.lap("  split: segment table");
  // phase 1: largest variant end per segment
  pool_->parallelFor((int) segs.size(), [&](int t, int) {
    Seg& sg = segs[t];
    long m = LONG_MIN;
    const int32_t* ce = pc.vEnd[0].data();
    const int32_t* be2 = pc.vEnd[1].data();
    for (int v = sg.i; v < sg.ie; ++v) m = std::max(m, (long) ce[v]);
    for (int v = sg.j; v < sg.je; ++v) m = std::max(m, (long) be2[v]);
    sg.maxEndLocal = m;
  });
  for (int k = 0; k < nSeq_; ++k) {
    long carry = LONG_MIN;
    for (int t = segFirst[k]; t < segFirst[k + 1]; ++t) {
      segs[t].carry = carry;
      carry = std::max(carry, segs[t].maxEndLocal);
    }
  }
  ts.lap("  split: largest ends");
  // phase 2: split every segment with the carried maximum end
  // Dense clusters (several variants per side a few bases apart, typically in a tandem repeat) rarely come back in sync
  // at a gap of a few bases: their pieces would SPILL into each other one merge and one full re-run at a time.
  // Neighbours closer than denseGap are joined up front when the chain they form is crowded.  Any coarser split is exact
  // (SURVEY.md A.13); sparse genomes have almost no such regions, so this costs them nothing.  Chains are looked for
  // within a segment (a chain across a segment boundary simply stays apart, as it would with a smaller denseGap).
  pool_->parallelFor((int) segs.size(), [&](int t, int) {
    Seg& sg = segs[t];
    sg.outp->reserve((size_t) std::max(sg.ie - sg.i, sg.je - sg.j) + 16);
    const bool firstOfSeq = t == segFirst[sg.seq];
    std::vector<uint8_t>& near = splitOut_[(size_t) t].near;
    near.clear();
    const bool dense = opt_.denseGap > opt_.gap;
    splitSlice(pc, sg.i, sg.ie, sg.j, sg.je, sg.carry, *sg.outp, firstOfSeq ? nullptr : &sg.continues, dense ? &near : nullptr);
    std::vector<RegRange>& out = *sg.outp;
    if (dense && out.size() > 1) {
      // chains of regions whose gaps stay below denseGap (the splitter noted them: a region ends at its largest variant
      // end, the next one starts at its first variant); a chain becomes ONE region when it is crowded as a whole
      size_t j = 0, w = 0;   // in place: the write index never passes the read index
      const size_t nOut = out.size();
      while (j < nOut) {
        if (!near[j]) { out[w++] = out[j++]; continue; }
        size_t e = j + 1;
        long nc = out[j].cCount, nb = out[j].bCount;
        while (e < nOut && near[e - 1]) {
          nc += out[e].cCount;
          nb += out[e].bCount;
          ++e;
        }
        if (e - j > 1 && (nc >= opt_.denseCount || nb >= opt_.denseCount)) {
          RegRange g = out[j];
          g.cCount = (int32_t) nc;
          g.bCount = (int32_t) nb;
          out[w++] = g;
        } else {
          for (size_t q = j; q < e; ++q) out[w++] = out[q];
        }
        j = e;
      }
      out.resize(w);
    }
  });
  ts.lap("  split: segments");
  // phase 3: stitch.  Where every segment's regions land in its sequence's table (a segment whose first region continues
  // the previous segment's last one gives that region's variants to it), then all segments copy side by side.
  struct Join { int seq; size_t at; int32_t c, b; };
  std::vector<size_t> segAt(segs.size(), 0), segFrom(segs.size(), 0);
  std::vector<Join> joins;
  std::vector<size_t> seqTotal((size_t) nSeq_, 0);
  for (int k = 0; k < nSeq_; ++k) {
    size_t n = 0;
    for (int t = segFirst[k]; t < segFirst[k + 1]; ++t) {
      const std::vector<RegRange>& so = *segs[t].outp;
      size_t from = 0;
      if (segs[t].continues && n > 0 && !so.empty()) {
        joins.push_back(Join{k, n - 1, so[0].cCount, so[0].bCount});
        from = 1;
      }
      segAt[(size_t) t] = n;
      segFrom[(size_t) t] = from;
      n += so.size() - from;
    }
    seqTotal[(size_t) k] = n;
  }
  pool_->parallelFor(nSeq_, [&](int k, int) {   // (the tables keep their capacity: growing is the rare case)
    const size_t n = pc.shortCircuit[k] ? 0 : seqTotal[(size_t) k];
    pc.rr[k].resize(n);
    pc.rs[k].resize(n);
    pc.rcls[k].resize(n);
  });
  pool_->parallelFor((int) segs.size(), [&](int t, int) {
    const Seg& sg = segs[t];
    const std::vector<RegRange>& so = *sg.outp;
    const size_t from = segFrom[(size_t) t], cnt = so.size() - from, at = segAt[(size_t) t];
    if (cnt == 0) return;
    std::memcpy(pc.rr[sg.seq].data() + at, so.data() + from, cnt * sizeof(RegRange));
    std::memset(static_cast<void*>(pc.rs[sg.seq].data() + at), 0, cnt * sizeof(RegRes));
    std::memset(pc.rcls[sg.seq].data() + at, 0, cnt);
  });
  for (const Join& jn : joins) {
    RegRange& g = pc.rr[jn.seq][jn.at];
    g.cCount += jn.c;
    g.bCount += jn.b;
  }
  ts.lap("  split: stitch + dense chains");
}

int Engine::regionStartPos(const PassCtx& pc, int k, int j) const {
  if (j == 0) return -1;
  const RegRange& g = pc.rr[k][j];
  int first = INT_MAX;
  for (int v = g.cFirst; v < g.cFirst + g.cCount; ++v) first = std::min(first, pc.vStart[0][v]);
  for (int v = g.bFirst; v < g.bFirst + g.bCount; ++v) first = std::min(first, pc.vStart[1][v]);
  return first - 1;
}

void Engine::planChunks(PassCtx& pc) {
  // (chunk objects persist across calls: their packet -> region lists keep their capacity)
  size_t nChunks = 0;
  size_t total = 0;
  for (int k = 0; k < nSeq_; ++k) total += pc.rr[k].size();
  int per = opt_.chunkRegions;
  if (per <= 0) {
    per = (int) (total / ((size_t) pool_->threads() * 6 + 1));
    per = std::max(8192, std::min(131072, per));
    // resident inputs + device-resident results: a chunk costs the host a list of regions, nothing else -- large chunks give
    // the search kernel full waves (one launch per chunk) and the host few launches
    if (devRes_ && pc.pass == 0) per = std::max(per, 1 << 18);
  }
  for (int k = 0; k < nSeq_; ++k) {
    const int n = (int) pc.rr[k].size();
    for (int r0 = 0; r0 < n; r0 += per) {
      if (pc.chunks.size() <= nChunks) pc.chunks.emplace_back();
      Chunk& ch = pc.chunks[nChunks++];
      ch.seq = k;
      ch.r0 = r0;
      ch.r1 = std::min(n, r0 + per);
      ch.list.clear();
      ch.transient = false;
      ch.variants = ch.variantsA = 0;
      for (int c = 0; c < 2; ++c) { ch.job[c] = MicroJob(); ch.pktRegion[c].clear(); }
    }
  }
  pc.chunks.resize(nChunks);
}

// ------------------------------------------------------------------------------------------- packets
// Writes the packet of region (k, j) in the layout of vce_micro.h; returns its length in bytes (a multiple of 4) or
// 0 when the region does not fit the thread-per-region kernel.
template <class T>
int Engine::buildPacket(const PassCtx&, const PacketSrc& ps, int j, uint8_t* out) const {
  if (j == 0) return 0;  // the first path of a sequence starts at -1 and is not in sync (HalfPath.java:50-57)
  const RegRange& g = ps.rr[j];
  if (g.cCount > T::KV || g.bCount > T::KV) return 0;
  int32_t winStart = 0, winLen = 0, nextStart[2];
  if (!microWindow(ps.pv, g, &winStart, &winLen, nextStart)) return 0;
  return microBuildPacket<T>(ps.pv, g, winStart, winLen, nextStart, ps.tmpl + winStart, out);
}

// The caller's arrays of sequence k as the packet builder reads them.
PacketView Engine::packetView(const PassCtx& pc, int k) const {
  PacketView pv;
  for (int side = 0; side < 2; ++side) {
    const SideSeq& ss = pc.ss[side][k];
    const vce_variants& in = *ss.in;
    PacketSideView& s = pv.s[side];
    s.vStart = pc.vStart[side].data();
    s.vEnd = pc.vEnd[side].data();
    s.sideEnd = ss.first + ss.n;
    s.first = ss.first;
    s.idx = ss.idx;
    s.gtPloidy = in.gt_ploidy;
    s.gt = in.gt;
    s.phased = in.phased;
    s.statusIn = in.status_in;
    s.alleleOff = in.allele_off;
    s.aStart = in.allele_start;
    s.aEnd = in.allele_end;
    s.aNull = in.allele_null;
    s.aNtOff = in.allele_nt_off;
    s.nt = in.nt;
    pv.expectPloidy[side] = pc.gtPloidy[side];
  }
  pv.tl = seqs_[k].template_len;
  pv.H = pc.H;
  pv.margin = opt_.microMargin;
  return pv;
}

static std::atomic<long long> g_tBuild(0), g_tCopy(0), g_tUpload(0);
static inline long long nowNs() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Builds the packets of a chunk -- the regions [r0, r1) of its sequence, or the explicit list ch.list -- one job per
// packet class, copies them to page-locked memory and starts the upload (and the kernel when `launch`).
int Engine::buildChunk(PassCtx& pc, Chunk& ch, int worker, bool launch) {
  const long long tA = nowNs();
  const int k = ch.seq;
  const bool listed = !ch.list.empty();
  const int nr = listed ? (int) ch.list.size() : ch.r1 - ch.r0;
  std::vector<uint8_t>* sc = scratch_[worker];
  std::vector<uint32_t>* so = scratchOff_[worker];
  // MicroA packets (the bulk) are built straight into page-locked memory, sized for the worst case; the few MicroB
  // packets go through a scratch buffer
  static const bool directPinned = std::getenv("VCE_DIRECT_PINNED") ? std::atoi(std::getenv("VCE_DIRECT_PINNED")) != 0 : true;
  static const bool inputPrefetch = std::getenv("VCE_PREFETCH") ? std::atoi(std::getenv("VCE_PREFETCH")) != 0 : true;
  uint8_t* pkA = nullptr;
  if (pc.microOk && directPinned) pkA = arenaAlloc((size_t) nr * MicroA::PKT_MAX + 16);
  else if (pc.microOk) { sc[0].resize((size_t) nr * MicroA::PKT_MAX + 16); pkA = sc[0].data(); }
  if (pc.microOk && !pkA) {
    std::lock_guard<std::mutex> lk(errMu_);
    err_ = "out of page-locked host memory";
    return VCE_ERR_CUDA;
  }
  so[0].resize(nr);
  size_t at[2] = {0, 0};
  int np[2] = {0, 0};
  for (int c = 0; c < 2; ++c) {
    ch.pktRegion[c].clear();
    ch.job[c] = MicroJob();
    ch.job[c].cls = c;
  }
  ch.pktRegion[0].reserve(nr);
  ch.variants = 0;
  ch.variantsA = 0;
  uint8_t* cls = pc.rcls[k].data();
  RegRes* rs = pc.rs[k].data();
  const RegRange* rrk = pc.rr[k].data();
  const uint8_t* tmpl = seqs_[k].template_bases;
  const int nReg = (int) pc.rr[k].size();
  const int32_t* cs = pc.vStart[0].data();
  const int32_t* bs = pc.vStart[1].data();
  PacketSrc ps;
  ps.rr = rrk;
  ps.tmpl = tmpl;
  ps.pv = packetView(pc, k);
  for (int q = 0; q < nr; ++q) {
    const int j = listed ? ch.list[q] : ch.r0 + q;
    if (!listed && cls[j] == kClsEarly) continue;  // already running in the warp-per-region kernels
    if (!listed && j + 12 < nReg) {  // the reference window of a region is a cache miss: request it well ahead
      const RegRange& f = rrk[j + 12];
      const int32_t pos = f.cCount > 0 ? cs[f.cFirst] : (f.bCount > 0 ? bs[f.bFirst] : 0);
      __builtin_prefetch(tmpl + (pos > 0 ? pos - 1 : 0));
      // the caller's arrays are two dozen sequential streams, more than the hardware prefetchers follow
      if (inputPrefetch && (j & 3) == 0) {
        for (int side = 0; side < 2; ++side) {
          const SideSeq& ssd = pc.ss[side][k];
          if (ssd.idx != nullptr) continue;
          const vce_variants& in = *ssd.in;
          const int li = (side == 0 ? f.cFirst : f.bFirst) - ssd.first + 8;
          if (li + 1 >= ssd.n) continue;
          const int sl = in.allele_off[li];
          __builtin_prefetch(in.allele_off + li + 16);
          __builtin_prefetch(in.allele_start + sl + 16);
          __builtin_prefetch(in.allele_end + sl + 16);
          __builtin_prefetch(in.allele_nt_off + sl + 16);
          __builtin_prefetch(in.allele_null + sl + 64);
          __builtin_prefetch(in.gt + (size_t) (li + 32) * in.gt_ploidy);
          __builtin_prefetch(in.nt + in.allele_nt_off[sl] + 64);
        }
      }
    }
    int bytes = 0, c = 0;
    if (pc.microOk) {
      const RegRange& g = rrk[j];
      if (g.cCount <= MicroA::KV && g.bCount <= MicroA::KV) {
        bytes = buildPacket<MicroA>(pc, ps, j, pkA + at[0]);
      } else if (g.cCount <= MicroB::KV && g.bCount <= MicroB::KV) {
        c = 1;
        if (sc[1].size() < at[1] + MicroB::PKT_MAX + 16) sc[1].resize(std::max(sc[1].size() * 2, at[1] + (size_t) 64 * MicroB::PKT_MAX));
        bytes = buildPacket<MicroB>(pc, ps, j, sc[1].data() + at[1]);
      }
    }
    if (bytes > 0) {
      if (c == 1 && (int) so[1].size() <= np[1]) so[1].resize(std::max<size_t>(so[1].size() * 2, 64));
      so[c][np[c]++] = (uint32_t) (at[c] >> 2);
      ch.pktRegion[c].push_back(j);
      at[c] += (size_t) bytes;
      if (!ch.transient) cls[j] = (uint8_t) (1 + c);
      rs[j].state = RS_PENDING;
      ch.variants += rrk[j].cCount + rrk[j].bCount;
      if (c == 0) ch.variantsA += rrk[j].cCount + rrk[j].bCount;
    } else {
      if (!ch.transient) cls[j] = 0;
      rs[j].state = RS_RESIDUAL;
    }
  }
  const long long tB = nowNs();
  g_tBuild += tB - tA;
  int rc = VCE_OK;
  for (int c = 0; c < 2 && rc == VCE_OK; ++c) {
    MicroJob& job = ch.job[c];
    job.n = np[c];
    job.resBytes = c == 0 ? MicroA::RES_BYTES : MicroB::RES_BYTES;
    if (np[c] == 0) continue;
    uint8_t* pk = (c == 0 && directPinned) ? pkA : arenaAlloc(at[c] + 16);
    uint8_t* off = arenaAlloc((size_t) np[c] * 4);
    uint8_t* res = arenaAlloc((size_t) np[c] * job.resBytes);
    if (!pk || !off || !res) {
      std::lock_guard<std::mutex> lk(errMu_);
      err_ = "out of page-locked host memory";
      return VCE_ERR_CUDA;
    }
    if (c != 0 || !directPinned) std::memcpy(pk, sc[c].data(), at[c]);
    if (opt_.sortPackets && np[c] > 64) {
      // counting sort by shape class of the OFFSETS only: the kernel reaches every packet through its offset, so the
      // packet bytes stay where they were built
      std::vector<uint32_t>& cnt = classCount_[worker];
      std::vector<uint32_t>& keys = classKeys_[worker];
      cnt.assign(0x8000 + 1, 0);
      keys.resize((size_t) np[c]);
      for (int q = 0; q < np[c]; ++q) {
        keys[q] = microPacketClass(pk + (size_t) so[c][q] * 4);
        ++cnt[keys[q] + 1];
      }
      for (int kx = 0; kx < 0x8000; ++kx) cnt[kx + 1] += cnt[kx];   // first output index of every class
      std::vector<int32_t>& regs = classOrder_[worker];
      regs.assign(ch.pktRegion[c].begin(), ch.pktRegion[c].end());
      uint32_t* offOut = reinterpret_cast<uint32_t*>(off);
      for (int q = 0; q < np[c]; ++q) {
        const uint32_t o = cnt[keys[q]]++;
        offOut[o] = so[c][q];
        ch.pktRegion[c][o] = regs[q];
      }
    } else {
      std::memcpy(off, so[c].data(), (size_t) np[c] * 4);
    }
    job.packets = pk;
    job.packetBytes = at[c];
    job.offsets = reinterpret_cast<const uint32_t*>(off);
    job.results = res;
    const long long tC = nowNs();
    rc = be_->microUpload(job);
    if (rc == VCE_OK && launch) rc = be_->microRun(job, pc.mc);
    g_tUpload += nowNs() - tC;
  }
  g_tCopy += nowNs() - tB;
  if (rc) {
    std::lock_guard<std::mutex> lk(errMu_);
    err_ = be_->lastError();
  }
  return rc;
}

template <class T>
void Engine::decodeJob(PassCtx& pc, int k, const MicroJob& job, const std::vector<int32_t>& pktRegion, int64_t& iterations) {
  RegRes* rs = pc.rs[k].data();
  const RegRange* rr = pc.rr[k].data();
  uint8_t* dc = pc.decision[0].data();
  uint8_t* db = pc.decision[1].data();
  double* wt = pc.weight.data();
  int64_t it = 0;
  for (int q = 0; q < job.n; ++q) {
    const uint8_t* rec = job.results + (size_t) q * T::RES_BYTES;
    const int j = pktRegion[q];
    RegRes& s = rs[j];
    it += rec[MR_ITERS] | (rec[MR_ITERS + 1] << 8);
    if (rec[MR_STATUS] != kMicroClean) {
      s.state = RS_RESIDUAL;
      if (std::getenv("VCE_DEBUG_FALLBACK") && rr[j].cCount + rr[j].bCount == 1) {
        static std::atomic<int> shown(0);
        if (shown.fetch_add(1) < 6) {
          const RegRange& g = rr[j];
          const int side = g.cCount ? 0 : 1;
          const int v = side == 0 ? g.cFirst : g.bFirst;
          const int nxC = g.cFirst + g.cCount < pc.ss[0][k].first + pc.ss[0][k].n ? pc.vStart[0][g.cFirst + g.cCount] : -1;
          const int nxB = g.bFirst + g.bCount < pc.ss[1][k].first + pc.ss[1][k].n ? pc.vStart[1][g.bFirst + g.bCount] : -1;
          std::fprintf(stderr, "[vce] single-variant fallback: seq %d region %d side %d start %d end %d nextC %d nextB %d iters %d maxF %d\n", k, j, side,
                       pc.vStart[side][v], pc.vEnd[side][v], nxC, nxB, rec[MR_ITERS] | (rec[MR_ITERS + 1] << 8), rec[MR_MAXF]);
        }
      }
      continue;
    }
    const RegRange& g = rr[j];
    s.state = RS_MICRO_DONE;
    s.flags = (uint8_t) ((rec[MR_FLAGS] & 1) | ((rec[MR_FLAGS] & 2) ? 0x20 : 0));
    s.lastId = (int8_t) rec[MR_LASTID];
    s.nSync = rec[MR_NSYNC];
    for (int i = 0; i < T::SMAX; ++i) s.rel[i] = rec[MR_SYNC + i];
    const uint8_t* dec = rec + MR_SYNC + T::SMAX;
    const uint8_t* wn = dec + 2 * T::KV;
    const uint8_t* wd = wn + T::KV;
    for (int i = 0; i < g.cCount; ++i) {
      dc[g.cFirst + i] = dec[i];
      wt[g.cFirst + i] = dec[i] >= 2 ? (double) wn[i] / (double) wd[i] : 0.0;  // Path.java:450
    }
    for (int i = 0; i < g.bCount; ++i) db[g.bFirst + i] = dec[T::KV + i];
    if (liveResults_ != nullptr) assembleRegionEarly(pc, k, j);
  }
  iterations += it;
}

static inline void pickRow(int kind, int H, int gtPloidy, const vce_variants& in, int ii, int row, int8_t* ids, uint8_t* original);

// Single-pass runs through vce_submit: the per-variant outputs of a region the packet kernel settled are final as
// soon as its record is decoded (unless the region is later absorbed by a spilling neighbour or re-run for a prefix
// tie-break: settle() clears the marks then).  Writing them here, while the GPU is still busy with the warp-per-region
// batch, takes the bulk of the assembly off the critical path.  Same rules as assembleBlock().
void Engine::assembleRegionEarly(PassCtx& pc, int k, int j) {
  const RegRange& g = pc.rr[k][j];
  const RegRes& rsj = pc.rs[k][j];
  const int base = regionStartPos(pc, k, j);
  vce_sequence_result& res = liveResults_[k];
  const vce_sequence& sq = seqs_[k];
  const int first[2] = {g.cFirst, g.bFirst}, count[2] = {g.cCount, g.bCount};
  for (int side = 0; side < 2; ++side) {
    vce_side_result& o = side == 0 ? res.calls : res.baseline;
    const vce_variants& in = side == 0 ? sq.calls : sq.baseline;
    const SideSeq& a = pc.ss[side][k];
    for (int v = first[side]; v < first[side] + count[side]; ++v) {
      const int li = v - a.first;
      const int ii = a.idx ? a.idx[li] : li;
      int st = in.status_in ? in.status_in[ii] : 0;
      const int d0 = pc.decision[side][v];
      int8_t* ids = o.allele_ids + (size_t) ii * VCE_MAX_PLOIDY;
      double wgt = 0.0;
      if (d0 >= 2) {
        st |= VCE_STATUS_GT_MATCH;
        pickRow(pc.kind[side], pc.H, pc.gtPloidy[side], in, ii, d0 - 2, ids, &o.original[ii]);
        wgt = side == 0 ? pc.weight[v] : 0.0;
      } else {
        st |= d0 == 1 ? VCE_STATUS_NO_MATCH : VCE_STATUS_SKIPPED;
        ids[0] = ids[1] = ids[2] = ids[3] = -2;
        o.original[ii] = 0;
      }
      o.status[ii] = (uint8_t) st;
      o.weight[ii] = wgt;
      if (o.sync_id) {
        // floorSyncPos over the region's own sync points: the first one stands right before the region's first variant
        int val = 0;
        if (!(st & VCE_STATUS_SKIPPED)) {
          const int vv = pc.vStart[side][v] - 1;
          int fl = base + rsj.rel[0];
          for (int q = 1; q < rsj.nSync && base + rsj.rel[q] <= vv; ++q) fl = base + rsj.rel[q];
          val = fl + 2;
        }
        o.sync_id[ii] = val;
      }
      earlyDone_[side][v] = 1;
    }
  }
}

void Engine::decodeChunk(PassCtx& pc, Chunk& ch, int64_t& iterations) {
  decodeJob<MicroA>(pc, ch.seq, ch.job[0], ch.pktRegion[0], iterations);
  decodeJob<MicroB>(pc, ch.seq, ch.job[1], ch.pktRegion[1], iterations);
}

// ------------------------------------------------------------------------------------------- life cycle
int Engine::prepare(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs, bool staged) {
  staged_ = false;
  cfg_ = cfg;
  nSeq_ = nSeq;
  seqs_ = seqs;
  err_.clear();
  executed_ = false;
  if (cfg.n_passes < 1 || cfg.n_passes > 2) { err_ = "n_passes must be 1 or 2"; return VCE_ERR_BAD_ARGUMENT; }
  if (cfg.flag_alternates) { err_ = "flag_alternates is not supported"; return VCE_ERR_UNSUPPORTED; }
  for (int p = 0; p < cfg.n_passes; ++p) {
    if (cfg.ploidy[p] < 1 || cfg.ploidy[p] > VCE_MAX_PLOIDY) { err_ = "ploidy out of range"; return VCE_ERR_BAD_ARGUMENT; }
  }
  StageTimer tm;
  arenaReset();
  int rc = be_->microReset();
  if (rc) { err_ = be_->lastError(); return rc; }
  if ((rc = setupPass0())) return rc;
  tm.lap("bounds + order check");
  if (staged) {  // digest + upload now, so that execute() starts with everything resident in HBM
    PassCtx& pc = pass_[0];
    if ((rc = runChunks(pc))) return rc;
    tm.lap("split into regions");
    std::atomic<int> failed(0);
    pool_->parallelFor((int) pc.chunks.size(), [&](int t, int worker) {
      if (failed.load()) return;
      const int r = buildChunk(pc, pc.chunks[t], worker, false);
      if (r) failed.store(r);
    });
    if (failed.load()) return failed.load();
    arenaMark();
    be_->microMark();
    tm.lap("packets + upload");
    if (tm.on) {
      std::fprintf(stderr, "[vce]   summed over workers: build %.1f ms, copy to page-locked %.1f ms, upload calls %.1f ms\n",
                   g_tBuild.exchange(0) / 1e6, g_tCopy.exchange(0) / 1e6, g_tUpload.exchange(0) / 1e6);
    }
    staged_ = true;
  }
  return VCE_OK;
}

// Device-side digest of one chunk (vce_submit on sorted inputs): instead of chasing variant -> alleles -> bases per region,
// the host copies the slices of the caller's arrays that the chunk touches wholesale into one page-locked slab, adds the
// regions and their reference windows (the template itself stays on the host), and k_build packs the packets on the
// device (vce_packet.h is the one source of the packing for both sides).
int Engine::buildChunkDevice(PassCtx& pc, Chunk& ch, int worker) {
  const int k = ch.seq;
  const SideSeq* SS[2] = {&pc.ss[0][k], &pc.ss[1][k]};
  if (!pc.microOk || SS[0]->idx != nullptr || SS[1]->idx != nullptr || !ch.list.empty()) return buildChunk(pc, ch, worker, true);
  const long long tA = nowNs();
  uint8_t* cls = pc.rcls[k].data();
  RegRes* rs = pc.rs[k].data();
  const RegRange* rrk = pc.rr[k].data();
  const PacketView hv = packetView(pc, k);
  for (int c = 0; c < 2; ++c) {
    ch.pktRegion[c].clear();
    ch.job[c] = MicroJob();
    ch.job[c].cls = c;
  }
  ch.variants = 0;
  ch.variantsA = 0;
  thread_local std::vector<int32_t> wsv[2], wlv[2];
  for (int c = 0; c < 2; ++c) { wsv[c].clear(); wlv[c].clear(); }
  size_t winBytes = 0;
  int lo[2] = {INT_MAX, INT_MAX}, hi[2] = {-1, -1};
  for (int j = ch.r0; j < ch.r1; ++j) {
    if (cls[j] == kClsEarly) continue;  // already running in the warp-per-region kernels
    const RegRange& g = rrk[j];
    int c = -1;
    if (j != 0) {  // the first path of a sequence starts at -1 and is not in sync (HalfPath.java:50-57)
      if (g.cCount <= MicroA::KV && g.bCount <= MicroA::KV) c = 0;
      else if (g.cCount <= MicroB::KV && g.bCount <= MicroB::KV) c = 1;
    }
    int32_t ws = 0, wl = 0, ns[2];
    if (c >= 0 && !microWindow(hv, g, &ws, &wl, ns)) c = -1;
    if (c < 0) {
      if (!ch.transient) cls[j] = 0;
      rs[j].state = RS_RESIDUAL;
      continue;
    }
    ch.pktRegion[c].push_back(j);
    wsv[c].push_back(ws);
    wlv[c].push_back(wl);
    winBytes += (size_t) wl;
    cls[j] = (uint8_t) (1 + c);
    rs[j].state = RS_PENDING;
    ch.variants += g.cCount + g.bCount;
    if (c == 0) ch.variantsA += g.cCount + g.bCount;
    lo[0] = std::min(lo[0], g.cFirst); hi[0] = std::max(hi[0], g.cFirst + g.cCount);
    lo[1] = std::min(lo[1], g.bFirst); hi[1] = std::max(hi[1], g.bFirst + g.bCount);
  }
  const int nA = (int) ch.pktRegion[0].size(), nB = (int) ch.pktRegion[1].size();
  if (nA + nB == 0) { g_tBuild += nowNs() - tA; return VCE_OK; }
  // ---- slab layout
  MicroBuildJob bj;
  size_t at = 0;
  auto place = [&](size_t bytes) { const size_t o = at; at = (at + bytes + 15) & ~(size_t) 15; return (int64_t) o; };
  bj.regions = place((size_t) (nA + nB) * sizeof(RegRange));
  bj.scatter = devRes_;
  if (devRes_) bj.regIdx = place((size_t) (nA + nB) * 4);
  bj.winOff = place((size_t) (nA + nB) * 4);
  bj.windows = place(winBytes + 8);
  int nV[2], nVs[2], nSl[2], nNt[2];
  for (int side = 0; side < 2; ++side) {
    const SideSeq& ss = *SS[side];
    const vce_variants& in = *ss.in;
    MicroBuildSide& m = bj.side[side];
    nV[side] = hi[side] - lo[side];
    nVs[side] = std::min(hi[side] + 1, ss.first + ss.n) - lo[side];   // one more start: the next variant after the chunk
    m.v0 = lo[side];
    m.i0 = lo[side] - ss.first;
    m.slot0 = in.allele_off[m.i0];
    const int slotN = in.allele_off[m.i0 + nV[side]];
    nSl[side] = slotN - m.slot0;
    m.nt0 = in.allele_nt_off[m.slot0];
    nNt[side] = in.allele_nt_off[slotN] - m.nt0;
    m.sideEnd = ss.first + ss.n;
    m.first = ss.first;
    m.gtPloidy = in.gt_ploidy;
    m.vStart = place((size_t) nVs[side] * 4 + 4);
    m.vEnd = place((size_t) nVs[side] * 4 + 4);
    m.gt = place((size_t) nV[side] * (size_t) std::max(1, in.gt_ploidy) + 4);
    m.phased = in.phased ? place((size_t) nV[side] + 4) : -1;
    m.statusIn = in.status_in ? place((size_t) nV[side] + 4) : -1;
    m.alleleOff = place((size_t) (nV[side] + 1) * 4);
    m.aStart = place((size_t) nSl[side] * 4 + 4);
    m.aEnd = place((size_t) nSl[side] * 4 + 4);
    m.aNull = place((size_t) nSl[side] + 4);
    m.aNtOff = place((size_t) (nSl[side] + 1) * 4);
    m.nt = place((size_t) nNt[side] + 4);
  }
  uint8_t* slab = arenaAlloc(at + 16);
  if (!slab) {
    std::lock_guard<std::mutex> lk(errMu_);
    err_ = "out of page-locked host memory";
    return VCE_ERR_CUDA;
  }
  // ---- fill
  RegRange* regs = reinterpret_cast<RegRange*>(slab + bj.regions);
  uint32_t* winOff = reinterpret_cast<uint32_t*>(slab + bj.winOff);
  uint8_t* wins = slab + bj.windows;
  const uint8_t* tmpl = seqs_[k].template_bases;
  size_t wo = 0;
  int q = 0;
  for (int c = 0; c < 2; ++c) {
    const std::vector<int32_t>& list = ch.pktRegion[c];
    for (size_t i = 0; i < list.size(); ++i, ++q) {
      if (i + 24 < list.size()) __builtin_prefetch(tmpl + wsv[c][i + 24]);
      regs[q] = rrk[list[i]];
      if (devRes_) reinterpret_cast<int32_t*>(slab + bj.regIdx)[q] = regBase_[k] + list[i];
      winOff[q] = (uint32_t) wo;
      std::memcpy(wins + wo, tmpl + wsv[c][i], (size_t) wlv[c][i]);
      wo += (size_t) wlv[c][i];
    }
  }
  for (int side = 0; side < 2; ++side) {
    const vce_variants& in = *SS[side]->in;
    const MicroBuildSide& m = bj.side[side];
    const int P = in.gt_ploidy;
    std::memcpy(slab + m.vStart, pc.vStart[side].data() + m.v0, (size_t) nVs[side] * 4);
    std::memcpy(slab + m.vEnd, pc.vEnd[side].data() + m.v0, (size_t) nVs[side] * 4);
    if (P > 0) std::memcpy(slab + m.gt, in.gt + (size_t) m.i0 * P, (size_t) nV[side] * P);
    if (m.phased >= 0) std::memcpy(slab + m.phased, in.phased + m.i0, (size_t) nV[side]);
    if (m.statusIn >= 0) std::memcpy(slab + m.statusIn, in.status_in + m.i0, (size_t) nV[side]);
    std::memcpy(slab + m.alleleOff, in.allele_off + m.i0, (size_t) (nV[side] + 1) * 4);
    std::memcpy(slab + m.aStart, in.allele_start + m.slot0, (size_t) nSl[side] * 4);
    std::memcpy(slab + m.aEnd, in.allele_end + m.slot0, (size_t) nSl[side] * 4);
    std::memcpy(slab + m.aNull, in.allele_null + m.slot0, (size_t) nSl[side]);
    std::memcpy(slab + m.aNtOff, in.allele_nt_off + m.slot0, (size_t) (nSl[side] + 1) * 4);
    std::memcpy(slab + m.nt, in.nt + m.nt0, (size_t) nNt[side]);
  }
  bj.slab = slab;
  bj.slabBytes = at;
  bj.tl = hv.tl;
  bj.H = hv.H;
  bj.expectPloidy[0] = hv.expectPloidy[0];
  bj.expectPloidy[1] = hv.expectPloidy[1];
  bj.margin = hv.margin;
  bj.nA = nA;
  bj.nB = nB;
  for (int c = 0; c < 2; ++c) {
    MicroJob& job = ch.job[c];
    job.n = c == 0 ? nA : nB;
    job.resBytes = c == 0 ? MicroA::RES_BYTES : MicroB::RES_BYTES;
    if (job.n == 0) continue;
    job.results = arenaAlloc((size_t) job.n * job.resBytes);
    if (!job.results) {
      std::lock_guard<std::mutex> lk(errMu_);
      err_ = "out of page-locked host memory";
      return VCE_ERR_CUDA;
    }
  }
  const long long tB = nowNs();
  g_tBuild += tB - tA;
  const int rc = be_->microBuildRun(bj, ch.job[0], ch.job[1], pc.mc);
  g_tUpload += nowNs() - tB;
  g_tCopy += nowNs() - tB;
  if (rc) {
    std::lock_guard<std::mutex> lk(errMu_);
    err_ = be_->lastError();
  }
  return rc;
}

// Resident inputs.  For every sequence of the pass whose template is registered on the backend's device
// (vce_template_register) and whose sides are already in (start, end) order, the caller's arrays are uploaded WHOLE, once:
// one page-locked block per side (parallel memcpys of the arrays as they are, no per-region work), one asynchronous copy.
// The packet builder then reads the sequence through a PacketView of device pointers and cuts every reference window out of
// the device copy of the template: the host digest of such a sequence is the region split plus a list of regions per chunk.
int Engine::residentSetup(PassCtx& pc) {
  resident_.assign((size_t) nSeq_, ResidentSeq());
  if (!opt_.deviceBuild || !pc.microOk || pc.pass != 0) return VCE_OK;
  struct Task { int k, side; };
  std::vector<Task> tasks;
  for (int k = 0; k < nSeq_; ++k) {
    if (pc.shortCircuit[k] || pc.ss[0][k].idx != nullptr || pc.ss[1][k].idx != nullptr) continue;
    ResidentSeq& rsq = resident_[k];
    rsq.on = opt_.residentInputs && be_->templateLookup(seqs_[k].template_bases, seqs_[k].template_len, &rsq.t2, &rsq.nmask);
    // without a registered template only what the final passes read goes up whole (the packets come from chunk slabs)
    if (!rsq.on && !devRes_) continue;
    rsq.have = true;
    tasks.push_back(Task{k, 0});
    tasks.push_back(Task{k, 1});
  }
  if (tasks.empty()) return VCE_OK;
  std::sort(tasks.begin(), tasks.end(), [&](const Task& a, const Task& b) { return pc.ss[a.side][a.k].n > pc.ss[b.side][b.k].n; });
  std::atomic<int> failed(0);
  pool_->parallelFor((int) tasks.size(), [&](int t, int) {
    if (failed.load()) return;
    const int k = tasks[t].k, side = tasks[t].side;
    const SideSeq& ss = pc.ss[side][k];
    const vce_variants& in = *ss.in;
    const size_t n = (size_t) ss.n, nSl = (size_t) in.allele_off[ss.n], nNt = (size_t) in.allele_nt_off[nSl];
    const int P = in.gt_ploidy;
    const bool full = resident_[k].on;
    size_t at = 0;
    auto place = [&](size_t bytes) { const size_t o = at; at = (at + bytes + 15) & ~(size_t) 15; return o; };
    const size_t oVs = place(n * 4 + 4), oPh = place(n + 4), oSt = place(n + 4);
    size_t oVe = 0, oGt = 0, oAo = 0, oAs = 0, oAe = 0, oAn = 0, oNo = 0, oNt = 0;
    if (full) {
      oVe = place(n * 4 + 4); oGt = place(n * (size_t) std::max(1, P) + 4); oAo = place((n + 1) * 4);
      oAs = place(nSl * 4 + 4); oAe = place(nSl * 4 + 4); oAn = place(nSl + 4); oNo = place((nSl + 1) * 4); oNt = place(nNt + 4);
    }
    void* dev = nullptr;
    int ev = -1;
    // page-locked caller arrays (vce_host_register) go up from where they are; only the locus bounds (computed here) are staged
    const bool direct = (!in.phased || be_->hostPinned(in.phased, n)) && (!in.status_in || be_->hostPinned(in.status_in, n)) &&
                        (!full || ((P == 0 || be_->hostPinned(in.gt, n * (size_t) P)) && be_->hostPinned(in.allele_off, (n + 1) * 4) &&
                                   be_->hostPinned(in.allele_start, nSl * 4) && be_->hostPinned(in.allele_end, nSl * 4) &&
                                   be_->hostPinned(in.allele_null, nSl) && be_->hostPinned(in.allele_nt_off, (nSl + 1) * 4) &&
                                   be_->hostPinned(in.nt, nNt)));
    if (direct) {
      int lane = 0;
      int r = be_->residentAlloc(at, &dev, &lane);
      if (r) { failed.store(r); return; }
      uint8_t* dd = static_cast<uint8_t*>(dev);
      const size_t bBytes = n * 4;
      uint8_t* bounds = arenaAlloc(2 * bBytes + 32);
      if (!bounds) { failed.store(VCE_ERR_CUDA); return; }
      std::memcpy(bounds, pc.vStart[side].data() + ss.first, bBytes);
      r = be_->residentCopy(dd + oVs, bounds, bBytes, lane);
      if (!r && in.phased) r = be_->residentCopy(dd + oPh, in.phased, n, lane);
      if (!r && in.status_in) r = be_->residentCopy(dd + oSt, in.status_in, n, lane);
      if (!r && full) {
        std::memcpy(bounds + bBytes, pc.vEnd[side].data() + ss.first, bBytes);
        r = be_->residentCopy(dd + oVe, bounds + bBytes, bBytes, lane);
        if (!r && P > 0) r = be_->residentCopy(dd + oGt, in.gt, n * (size_t) P, lane);
        if (!r) r = be_->residentCopy(dd + oAo, in.allele_off, (n + 1) * 4, lane);
        if (!r) r = be_->residentCopy(dd + oAs, in.allele_start, nSl * 4, lane);
        if (!r) r = be_->residentCopy(dd + oAe, in.allele_end, nSl * 4, lane);
        if (!r) r = be_->residentCopy(dd + oAn, in.allele_null, nSl, lane);
        if (!r) r = be_->residentCopy(dd + oNo, in.allele_nt_off, (nSl + 1) * 4, lane);
        if (!r) r = be_->residentCopy(dd + oNt, in.nt, nNt, lane);
      }
      if (!r) r = be_->residentMark(lane, &ev);
      if (r) { failed.store(r); return; }
    } else {
      uint8_t* blk = arenaAlloc(at + 16);
      if (!blk) { failed.store(VCE_ERR_CUDA); return; }
      std::memcpy(blk + oVs, pc.vStart[side].data() + ss.first, n * 4);
      if (in.phased) std::memcpy(blk + oPh, in.phased, n);
      if (in.status_in) std::memcpy(blk + oSt, in.status_in, n);
      if (full) {
        std::memcpy(blk + oVe, pc.vEnd[side].data() + ss.first, n * 4);
        if (P > 0) std::memcpy(blk + oGt, in.gt, n * (size_t) P);
        std::memcpy(blk + oAo, in.allele_off, (n + 1) * 4);
        std::memcpy(blk + oAs, in.allele_start, nSl * 4);
        std::memcpy(blk + oAe, in.allele_end, nSl * 4);
        std::memcpy(blk + oAn, in.allele_null, nSl);
        std::memcpy(blk + oNo, in.allele_nt_off, (nSl + 1) * 4);
        std::memcpy(blk + oNt, in.nt, nNt);
      }
      const int r = be_->residentUpload(blk, at, &dev, &ev);
      if (r) { failed.store(r); return; }
    }
    const uint8_t* d = static_cast<const uint8_t*>(dev);
    ResidentSeq& rsq = resident_[k];
    rsq.fVStart[side] = reinterpret_cast<const int32_t*>(d + oVs);
    rsq.fPhased[side] = in.phased ? d + oPh : nullptr;
    rsq.fStatus[side] = in.status_in ? d + oSt : nullptr;
    rsq.events[side] = ev;
    if (!full) return;
    PacketSideView& v = rsq.dv.s[side];
    v.vStart = reinterpret_cast<const int32_t*>(d + oVs) - ss.first;   // indexed with pass-wide variant indices
    v.vEnd = reinterpret_cast<const int32_t*>(d + oVe) - ss.first;
    v.sideEnd = ss.first + ss.n;
    v.first = ss.first;
    v.idx = nullptr;
    v.gtPloidy = P;
    v.gt = reinterpret_cast<const int8_t*>(d + oGt);
    v.phased = in.phased ? d + oPh : nullptr;
    v.statusIn = in.status_in ? d + oSt : nullptr;
    v.alleleOff = reinterpret_cast<const int32_t*>(d + oAo);
    v.aStart = reinterpret_cast<const int32_t*>(d + oAs);
    v.aEnd = reinterpret_cast<const int32_t*>(d + oAe);
    v.aNull = d + oAn;
    v.aNtOff = reinterpret_cast<const int32_t*>(d + oNo);
    v.nt = d + oNt;
  });
  if (failed.load()) {
    err_ = be_->lastError();
    if (err_.empty()) err_ = "out of page-locked host memory";
    return failed.load();
  }
  for (int k = 0; k < nSeq_; ++k) {
    ResidentSeq& rsq = resident_[k];
    if (!rsq.on) continue;
    rsq.dv.tl = seqs_[k].template_len;
    rsq.dv.H = pc.H;
    rsq.dv.margin = opt_.microMargin;
    rsq.dv.expectPloidy[0] = pc.gtPloidy[0];
    rsq.dv.expectPloidy[1] = pc.gtPloidy[1];
  }
  return VCE_OK;
}

// One chunk of a resident sequence: the packet-shaped regions go to the device as they are (16 bytes each); windows,
// alleles and genotypes are read there.  A region the builder cannot pack comes back as a fallback record.
int Engine::buildChunkResident(PassCtx& pc, Chunk& ch, int worker) {
  const int k = ch.seq;
  if (!resident_[k].on || !ch.list.empty()) return opt_.deviceBuild ? buildChunkDevice(pc, ch, worker) : buildChunk(pc, ch, worker, true);
  const long long tA = nowNs();
  uint8_t* cls = pc.rcls[k].data();
  RegRes* rs = pc.rs[k].data();
  const RegRange* rrk = pc.rr[k].data();
  for (int c = 0; c < 2; ++c) {
    ch.pktRegion[c].clear();
    ch.job[c] = MicroJob();
    ch.job[c].cls = c;
  }
  ch.variants = 0;
  ch.variantsA = 0;
  for (int j = ch.r0; j < ch.r1; ++j) {
    if (cls[j] == kClsEarly) continue;  // already running in the warp-per-region kernels
    const RegRange& g = rrk[j];
    int c = -1;
    if (j != 0) {  // the first path of a sequence starts at -1 and is not in sync (HalfPath.java:50-57)
      if (g.cCount <= MicroA::KV && g.bCount <= MicroA::KV) c = 0;
      else if (g.cCount <= MicroB::KV && g.bCount <= MicroB::KV) c = 1;
    }
    if (c < 0) {
      if (!ch.transient) cls[j] = 0;
      rs[j].state = RS_RESIDUAL;
      continue;
    }
    ch.pktRegion[c].push_back(j);
    cls[j] = (uint8_t) (1 + c);
    rs[j].state = RS_PENDING;
    ch.variants += g.cCount + g.bCount;
    if (c == 0) ch.variantsA += g.cCount + g.bCount;
  }
  const int nA = (int) ch.pktRegion[0].size(), nB = (int) ch.pktRegion[1].size();
  if (nA + nB == 0) { g_tBuild += nowNs() - tA; return VCE_OK; }
  MicroBuildJob bj;
  const size_t regBytes = ((size_t) (nA + nB) * sizeof(RegRange) + 15) & ~(size_t) 15;
  const size_t bytes = regBytes + (devRes_ ? (((size_t) (nA + nB) * 4 + 15) & ~(size_t) 15) : 0);
  uint8_t* slab = arenaAlloc(bytes + 16);
  if (!slab) {
    std::lock_guard<std::mutex> lk(errMu_);
    err_ = "out of page-locked host memory";
    return VCE_ERR_CUDA;
  }
  RegRange* regs = reinterpret_cast<RegRange*>(slab);
  int32_t* regIdx = reinterpret_cast<int32_t*>(slab + regBytes);
  int q = 0;
  for (int c = 0; c < 2; ++c) {
    for (int32_t j : ch.pktRegion[c]) {
      if (devRes_) regIdx[q] = regBase_[k] + j;
      regs[q++] = rrk[j];
    }
  }
  bj.scatter = devRes_;
  bj.regIdx = (int64_t) regBytes;
  const ResidentSeq& rsq = resident_[k];
  bj.slab = slab;
  bj.slabBytes = bytes;
  bj.regions = 0;
  bj.resident = true;
  bj.rv = rsq.dv;
  bj.t2 = rsq.t2;
  bj.nmask = rsq.nmask;
  bj.waitEvents[0] = rsq.events[0];
  bj.waitEvents[1] = rsq.events[1];
  bj.tl = rsq.dv.tl;
  bj.H = rsq.dv.H;
  bj.margin = rsq.dv.margin;
  bj.expectPloidy[0] = rsq.dv.expectPloidy[0];
  bj.expectPloidy[1] = rsq.dv.expectPloidy[1];
  bj.nA = nA;
  bj.nB = nB;
  for (int c = 0; c < 2; ++c) {
    MicroJob& job = ch.job[c];
    job.n = c == 0 ? nA : nB;
    job.resBytes = c == 0 ? MicroA::RES_BYTES : MicroB::RES_BYTES;
    if (job.n == 0) continue;
    job.results = arenaAlloc((size_t) job.n * job.resBytes);
    if (!job.results) {
      std::lock_guard<std::mutex> lk(errMu_);
      err_ = "out of page-locked host memory";
      return VCE_ERR_CUDA;
    }
  }
  const long long tB = nowNs();
  g_tBuild += tB - tA;
  const int rc = be_->microBuildRun(bj, ch.job[0], ch.job[1], pc.mc);
  g_tUpload += nowNs() - tB;
  if (rc) {
    std::lock_guard<std::mutex> lk(errMu_);
    err_ = be_->lastError();
  }
  return rc;
}

// Splits the pass into regions, builds + uploads (+ starts, when `launch`) every chunk.
static inline bool orientorPairMicro(int kc, int kb, int H) { return microOrientorSupported(kc, H) && microOrientorSupported(kb, H); }

void Engine::resetPassRun(PassCtx& pc) {
  for (size_t i = pc.undo.size(); i-- > 0;) pc.rr[pc.undo[i].first.seq][pc.undo[i].first.j] = pc.undo[i].second;
  pc.undo.clear();
  pc.rare.clear();
  pc.warns.clear();
  pc.wideSync.clear();
  pool_->parallelFor(nSeq_, [&](int k, int) {
    RegRes* rs = pc.rs[k].data();
    const uint8_t* cls = pc.rcls[k].data();
    const size_t n = pc.rs[k].size();
    for (size_t j = 0; j < n; ++j) {
      std::memset(&rs[j], 0, sizeof(RegRes));
      rs[j].state = cls[j] == kClsEarly ? RS_EARLY : (cls[j] ? RS_PENDING : RS_RESIDUAL);
    }
  });
}

void Engine::configurePass(PassCtx& pc) {
  pc.microOk = !opt_.disableMicro && !opt_.forceGeneral && orientorPairMicro(pc.kind[0], pc.kind[1], pc.H);
  pc.mc.H = pc.H;
  pc.mc.kind[0] = pc.kind[0];
  pc.mc.kind[1] = pc.kind[1];
  pc.mc.maxPaths = cfg_.max_paths;
  pc.mc.maxIterations = cfg_.max_iterations;
  pc.mc.pruneNoOps = cfg_.prune_no_ops;
  pc.mc.preference = cfg_.preference;
}

int Engine::runChunks(PassCtx& pc) {
  // split (per sequence), plan, build + upload + start
  configurePass(pc);
  splitAll(pc);
  StageTimer ts;
  // regions that cannot go through the thread-per-region kernel are known from their shape alone: they start right
  // away in the warp-per-region kernels, concurrently with the packet pipeline
  {
    struct Blk { int k, j0, j1; };
    std::vector<Blk> blks;
    const int perBlk = 1 << 15;
    for (int k = 0; k < nSeq_; ++k) {
      const int n = (int) pc.rr[k].size();
      for (int j0 = 0; j0 < n; j0 += perBlk) blks.push_back(Blk{k, j0, std::min(n, j0 + perBlk)});
    }
    std::vector<std::vector<std::pair<RegKey, int>>> per(blks.size());
    pool_->parallelFor((int) blks.size(), [&](int t, int) {
      const Blk& bl = blks[(size_t) t];
      const int k = bl.k;
      const RegRange* rr = pc.rr[k].data();
      const int n = (int) pc.rr[k].size();
      for (int j = bl.j0; j < bl.j1; ++j) {
        const RegRange& g = rr[j];
        // four or more variants on one side mostly outgrow the packet kernel's frontier: start them in the warp tiers
        const bool shape = g.cCount >= MicroB::KV || g.bCount >= MicroB::KV || j == 0 || j == n - 1;
        if (shape || !pc.microOk) {
          per[(size_t) t].push_back(std::make_pair(RegKey{k, j}, startTier(pc, g)));
          pc.rcls[k][j] = kClsEarly;
          pc.rs[k][j].state = RS_EARLY;
        }
      }
    });
    pc.earlyItems.clear();
    pc.earlyBatch = WideBatch();
    for (size_t t = 0; t < blks.size(); ++t) pc.earlyItems.insert(pc.earlyItems.end(), per[t].begin(), per[t].end());
  }
  ts.lap("  split: early regions");
  planChunks(pc);
  ts.lap("  split: chunk plan");
  return VCE_OK;
}

int Engine::execute() {
  int rc;
  StageTimer tm;
  microMs = 0;
  microRegions = microVariants = 0;
  devRes_ = !staged_ && liveResults_ != nullptr && wantDeviceResults(pass_[0]);
  if (devRes_) liveResults_ = nullptr;   // nothing is assembled on the host while records are decoded: they never come back
  earlyUsed_ = liveResults_ != nullptr;
  if (earlyUsed_) {
    for (int side = 0; side < 2; ++side) {
      ensureSize(earlyDone_[side], pass_[0].vStart[side].size());
      // cleared in parallel below, per chunk of variants
    }
    const size_t blockE = (size_t) 1 << 20;
    for (int side = 0; side < 2; ++side) {
      const size_t nE = pass_[0].vStart[side].size();
      pool_->parallelFor((int) ((nE + blockE - 1) / blockE), [&](int t, int) {
        std::memset(earlyDone_[side].data() + (size_t) t * blockE, 0, std::min(blockE, nE - (size_t) t * blockE));
      });
    }
  }
  for (int p = 0; p < cfg_.n_passes; ++p) {
    PassCtx& pc = pass_[p];
    const bool rerun = p == 0 && staged_;
    if (p == 1) {
      if ((rc = setupPass1())) return rc;
      tm.lap("pass 2 set-up");
    }
    std::atomic<int> failed(0);
    WideBatch localEarly;
    // a staged job keeps the packed early batch: re-runs start from the same regions
    WideBatch& early = staged_ && p == 0 ? pc.earlyBatch : localEarly;
    if (!rerun) {
      // the whole-sequence uploads go first: their DMA runs while the host splits the sequences into regions
      configurePass(pc);
      if ((rc = residentSetup(pc))) return rc;
      tm.lap("resident inputs: staging + upload");
      if ((rc = runChunks(pc))) return rc;
      tm.lap("split into regions");
      // The packet kernels go first: the warp-per-region kernel of the mid tier keeps every SM's register file for itself until
      // its longest regions are done (measured with VCE_TIMELINE: k_build sat behind it for 11 ms), whereas the packet kernels'
      // blocks are short-lived -- and the early batch is packed while they run.
      if (devRes_ && p == 0 && (rc = resultsSetup(pc))) return rc;
      if (devRes_ && p == 0) {
        ResultsSetup rsu;
        rsu.H = pc.H;
        rsu.kind[0] = pc.kind[0];
        rsu.kind[1] = pc.kind[1];
        rsu.seq.resize((size_t) nSeq_);
        for (int k = 0; k < nSeq_; ++k) {
          ResultsSeq& q = rsu.seq[(size_t) k];
          for (int side = 0; side < 2; ++side) {
            q.first[side] = pc.ss[side][k].first;
            q.n[side] = pc.ss[side][k].n;
            q.vStart[side] = resident_[k].fVStart[side];
            q.phased[side] = resident_[k].fPhased[side];
            q.statusIn[side] = resident_[k].fStatus[side];
          }
          q.regFirst = regBase_[k];
          q.nReg = regBase_[k + 1] - regBase_[k];
        }
        for (int side = 0; side < 2; ++side) rsu.nVar[side] = nSeq_ > 0 ? pc.ss[side][nSeq_ - 1].first + pc.ss[side][nSeq_ - 1].n : 0;
        rsu.nReg = regBase_[nSeq_];
        if ((rc = be_->resultsBegin(rsu))) { err_ = be_->lastError(); return rc; }
        tm.lap("device results: set-up");
      }
      pool_->parallelFor((int) pc.chunks.size(), [&](int t, int worker) {
        if (failed.load()) return;
        const int r = opt_.deviceBuild ? buildChunkResident(pc, pc.chunks[t], worker) : buildChunk(pc, pc.chunks[t], worker, true);
        if (r) failed.store(r);
      });
      if (opt_.deviceBuild && !failed.load() && (rc = be_->microBuildFlush(pc.mc))) { err_ = be_->lastError(); return rc; }
      tm.lap("packets + upload + launch");
      if (tm.on) {
        std::fprintf(stderr, "[vce]   summed over workers: build %.1f ms, copy to page-locked %.1f ms, upload+launch calls %.1f ms\n",
                     g_tBuild.exchange(0) / 1e6, g_tCopy.exchange(0) / 1e6, g_tUpload.exchange(0) / 1e6);
      }
      if (!failed.load()) {
        if ((rc = wideStart(pc, pc.earlyItems, early))) return rc;
        tm.lap("early warp-per-region batch");
      }
    } else {
      if (executed_) {
        resetPassRun(pc);
        arenaRewind();
        be_->microRewind();
        tm.lap("reset for re-run");
      }
      // the packets are resident: their kernels start first and cover the packing of the warp-per-region batch
      for (int c = 0; c < 2; ++c) {
        std::vector<MicroJob*> jobs;
        for (Chunk& ch : pc.chunks) if (ch.job[c].n > 0) jobs.push_back(&ch.job[c]);
        if (!jobs.empty()) {
          const int r = be_->microRunAll(jobs, pc.mc);
          if (r) { err_ = be_->lastError(); return r; }
        }
      }
      tm.lap("packet kernel launches");
      if (!early.packed && (rc = widePack(pc, pc.earlyItems, early))) return rc;
      if ((rc = wideLaunch(early))) return rc;
      tm.lap("early warp-per-region batch");
    }
    if (failed.load()) return failed.load();
    std::vector<int64_t> iters(pool_->threads(), 0);
    if (devRes_ && p == 0) {
      if ((rc = resultsCollect(pc))) return rc;
    }
    // the MicroA records of every chunk first: the MicroB-shaped regions may still be in their one late launch
    for (int c = 0; c < 2 && !(devRes_ && p == 0); ++c) {
      pool_->parallelFor((int) pc.chunks.size(), [&](int t, int worker) {
        Chunk& ch = pc.chunks[t];
        if (ch.job[c].n == 0) return;
        const int r = be_->microWait(ch.job[c]);
        if (r) { failed.store(r); return; }
        if (c == 0) decodeJob<MicroA>(pc, ch.seq, ch.job[0], ch.pktRegion[0], iters[worker]);
        else decodeJob<MicroB>(pc, ch.seq, ch.job[1], ch.pktRegion[1], iters[worker]);
      });
    }
    if (failed.load()) { err_ = be_->lastError(); return failed.load(); }
    be_->microCollect();
    tm.lap(devRes_ && p == 0 ? "wait for the packet kernels + region table" : "wait + decode");
    for (int64_t v : iters) be_->stats.iterations += v;
    for (const Chunk& ch : pc.chunks) {
      be_->stats.regions += ch.r1 - ch.r0;
      microRegions += ch.job[0].n + ch.job[1].n;
      microVariants += ch.variantsA;
    }
    if ((rc = settle(pc, early))) return rc;
    tm.lap("settle residual");
  }
  executed_ = true;
  return VCE_OK;
}

int Engine::run(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs, vce_sequence_result* results) {
  StageTimer tm;
  int rc = prepare(cfg, nSeq, seqs, false);
  if (rc) return rc;
  liveResults_ = cfg.n_passes == 1 ? results : nullptr;
  rc = execute();
  liveResults_ = nullptr;
  if (rc) return rc;
  rc = finish(results);
  tm.lap("vce_submit total");
  return rc;
}

// ------------------------------------------------------------------------------------------- warp-per-region path
// Runs the regions `ids` through the warp-per-region kernels (vce_search.h).  The variants and alleles of just these
// regions are packed into compact arrays, so that the cost is proportional to the residual, not to the batch.
// Tier of the warp-per-region kernels a region starts in: small regions in the many-warps tier, regions that will
// grow a large frontier (many variants on one side) directly in the large shared-memory tier, the rest in HBM arenas.
int Engine::startTier(const PassCtx& pc, const RegRange& g) const {
  if (opt_.forceGeneral) return kTierGeneral;
  if (pc.H > 2) return kTierGeneral;   // the CTA kernel's keys hold two haplotypes per side
  // the frontier grows with the variants on one side (orientations multiply): pick the tier that will hold it
  const int m = std::max(g.cCount, g.bCount);
  if (opt_.forceCta) return m >= 10 ? kTierCta : kTierMid;
  return m <= 3 ? 0 : kTierMid;   // what outgrows the mid tier's 2048 paths moves on to the full-budget tiers
}

// Packs the regions `items` (region, tier) for the warp-per-region kernels and starts them: the variants and alleles of
// just these regions go into compact arrays (cost proportional to the residual, not to the batch), K1 expands the
// orientations on the device, one launch per tier.  Returns without waiting; wideFinish collects.
int Engine::wideStart(PassCtx& pc, const std::vector<std::pair<RegKey, int>>& items, WideBatch& wb) {
  const int rc = widePack(pc, items, wb);
  return rc ? rc : wideLaunch(wb);
}

int Engine::widePack(PassCtx& pc, const std::vector<std::pair<RegKey, int>>& itemsIn, WideBatch& wb) {
  wb = WideBatch();
  if (itemsIn.empty()) return VCE_OK;
  StageTimer tmw;
  // by tier; within the arena tiers the regions with the most variants first: the kernels pull regions from a work
  // queue, and the longest searches should not be the last ones to start
  struct Item { RegKey id; int32_t tier, weight; };
  std::vector<Item> items(itemsIn.size());
  pool_->parallelFor((int) ((itemsIn.size() + 1023) / 1024), [&](int t, int) {
    for (size_t q = (size_t) t * 1024; q < std::min(itemsIn.size(), (size_t) (t + 1) * 1024); ++q) {
      const RegKey id = itemsIn[q].first;
      const RegRange& g = pc.rr[id.seq][id.j];
      items[q] = Item{id, itemsIn[q].second, itemsIn[q].second >= kTierMid ? g.cCount + g.bCount : 0};
    }
  });
  std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
    return a.tier != b.tier ? a.tier < b.tier : a.weight > b.weight;
  });
  tmw.lap("  wide: order");
  const size_t n = items.size();
  wb.ids.resize(n);
  wb.tier.resize(n);
  for (size_t q = 0; q < n; ++q) { wb.ids[q] = items[q].id; wb.tier[q] = items[q].tier; }
  PassData& pd = wb.pd;
  pd.pass = pc.pass;
  pd.H = pc.H;
  pd.kind[0] = pc.kind[0];
  pd.kind[1] = pc.kind[1];
  pd.wantOrientTables = false;
  wb.descs.assign(n, RegionDesc());
  wb.syncOff.assign(n, 0);
  // pass A (parallel): sizes per region
  struct Sz { int32_t nv[2], nsl[2], nnt[2], lo, maxSl[2], nextStart[2]; long hi; };
  std::vector<Sz> sz(n);
  const int blockA = 256;
  pool_->parallelFor((int) ((n + blockA - 1) / blockA), [&](int t, int) {
    for (size_t q = (size_t) t * blockA; q < std::min(n, (size_t) (t + 1) * blockA); ++q) {
      const int k = wb.ids[q].seq, j = wb.ids[q].j;
      const RegRange& g = pc.rr[k][j];
      Sz& z = sz[q];
      z.lo = INT_MAX;
      z.hi = 0;
      const int first[2] = {g.cFirst, g.bFirst}, count[2] = {g.cCount, g.bCount};
      for (int side = 0; side < 2; ++side) {
        const SideSeq& s = pc.ss[side][k];
        const vce_variants& in = *s.in;
        z.nextStart[side] = first[side] + count[side] < s.first + s.n ? pc.vStart[side][first[side] + count[side]] : kNoNext;
        z.nv[side] = count[side];
        z.nsl[side] = 0;
        z.nnt[side] = 0;
        z.maxSl[side] = 2;
        for (int v = first[side]; v < first[side] + count[side]; ++v) {
          const int li = v - s.first;
          const int ii = s.idx ? s.idx[li] : li;
          const int off0 = in.allele_off[ii], nSl = in.allele_off[ii + 1] - off0;
          z.nsl[side] += nSl;
          z.nnt[side] += in.allele_nt_off[off0 + nSl] - in.allele_nt_off[off0];
          z.maxSl[side] = std::max(z.maxSl[side], nSl);
          z.lo = std::min(z.lo, pc.vStart[side][v]);
          z.hi = std::max(z.hi, (long) pc.vEnd[side][v]);
        }
      }
    }
  });
  tmw.lap("  wide: sizes");
  // serial prefix: offsets of every region in the compact arrays, windows, sync space
  std::vector<int32_t> vOff[2], sOff[2], nOff[2];
  int maxSl[2] = {2, 2};
  for (int side = 0; side < 2; ++side) { vOff[side].resize(n + 1); sOff[side].resize(n + 1); nOff[side].resize(n + 1); }
  size_t so = 0, winBytes = 0;
  {
    int32_t av[2] = {0, 0}, as[2] = {0, 0}, an[2] = {0, 0};
    for (size_t q = 0; q < n; ++q) {
      const int k = wb.ids[q].seq, j = wb.ids[q].j;
      RegionDesc& d = wb.descs[q];
      std::memset(&d, 0, sizeof(d));
      for (int side = 0; side < 2; ++side) {
        vOff[side][q] = av[side]; sOff[side][q] = as[side]; nOff[side][q] = an[side];
        av[side] += sz[q].nv[side]; as[side] += sz[q].nsl[side]; an[side] += sz[q].nnt[side];
        maxSl[side] = std::max(maxSl[side], sz[q].maxSl[side]);
      }
      d.seq = k;
      d.tmplLen = seqs_[k].template_len;
      d.cFirst = vOff[0][q]; d.cCount = sz[q].nv[0]; d.bFirst = vOff[1][q]; d.bCount = sz[q].nv[1];
      d.cSlot0 = sOff[0][q]; d.cSlotN = sz[q].nsl[0]; d.bSlot0 = sOff[1][q]; d.bSlotN = sz[q].nsl[1];
      d.cNextStart = sz[q].nextStart[0];
      d.bNextStart = sz[q].nextStart[1];
      int lo = sz[q].lo;
      d.startPos = j == 0 ? -1 : lo - 1;
      d.prefixAlleleId = j == 0 ? kPrefixEmpty : kDummyPrefix;
      int extra = 0;
      if (!pc.rare.empty()) {
        const auto rit = pc.rare.find(key(k, j));
        if (rit != pc.rare.end()) {
          if (rit->second.hasPrefix) d.prefixAlleleId = rit->second.assumedPrefix;
          extra = rit->second.extraMargin;
        }
      }
      if (lo == INT_MAX) lo = 0;
      // A sequence's first region starts its path at -1 and jumps to the base before its first variant without reading
      // anything in between (PathFinder.skipToNextVariant :271-282 -> moveForward): its window starts where any other
      // region's does.  (It used to start at 0: harmless for a chromosome, but a PIECE of a cut sequence, vce_split.h, starts
      // tens of megabases in.)
      const int ws = std::max(0, std::min(lo, d.startPos < 0 ? lo - 1 : d.startPos) - 1);
      // regions of the full-budget tiers sit in repeats more often than not: a wider window from the start (there are few of them)
      const int baseMargin = wb.tier[q] >= kTierCta ? std::max(opt_.margin, opt_.wideMargin) : opt_.margin;
      long we = extra < 0 ? (long) d.tmplLen : sz[q].hi + baseMargin + extra;
      we = std::min(we, (long) d.tmplLen);
      if (we < ws) we = ws;
      d.winStart = ws;
      d.winLen = (int32_t) (we - ws);
      d.winOff = (uint32_t) winBytes;
      winBytes += ((size_t) d.winLen + 15 + 16) & ~(size_t) 15;
      wb.syncOff[q] = (int32_t) so;
      so += (size_t) d.cCount + d.bCount + 2;
    }
    for (int side = 0; side < 2; ++side) { vOff[side][n] = av[side]; sOff[side][n] = as[side]; nOff[side][n] = an[side]; }
  }
  wb.so = so;
  if (tmw.on) {
    long mx = 0;
    size_t at = 0;
    for (size_t q = 0; q < n; ++q) if (wb.descs[q].winLen > mx) { mx = wb.descs[q].winLen; at = q; }
    if (n) std::fprintf(stderr, "[vce]   wide: %zu window bytes over %zu regions; longest window %ld (region %d of its sequence, start %d, tier %d)\n",
                        winBytes, n, mx, wb.ids[at].j, wb.descs[at].winStart, wb.tier[at]);
  }
  for (int side = 0; side < 2; ++side) {
    HostSide& hs = pd.side[side];
    hs.gtPloidy = pc.gtPloidy[side];
    const size_t nv = (size_t) vOff[side][n], ns = (size_t) sOff[side][n];
    hs.vStart.resize(nv); hs.vEnd.resize(nv); hs.slot0.resize(nv); hs.nSl.resize(nv); hs.vFlags.resize(nv); hs.gt.resize(nv * 4);
    wb.at[side].aStart.resize(ns); wb.at[side].aEnd.resize(ns); wb.at[side].aNull.resize(ns); wb.at[side].aNtOff.resize(ns + 1);
    wb.at[side].nt.resize((size_t) nOff[side][n]);
    wb.at[side].aNtOff[ns] = nOff[side][n];
    hs.seqFirst.assign(2, 0);
    hs.seqFirst[1] = (int32_t) nv;
    pd.maxOrient[side] = orientBound(pc.kind[side], pc.H, pc.gtPloidy[side], maxSl[side]);
    if (pd.maxOrient[side] + 2 > 255) { err_ = "too many orientations per variant for this engine"; return VCE_ERR_UNSUPPORTED; }
  }
  const int mo = std::max(pd.maxOrient[0], pd.maxOrient[1]);
  if (1 + mo > FastLimits::kMaxChildren) for (size_t q = 0; q < n; ++q) wb.tier[q] = kTierGeneral;   // more children than the fixed shapes hold
  for (int t = 0; t <= kTierGeneral + 1; ++t) wb.tierFirst[t] = 0;
  for (size_t q = 0; q < n; ++q) ++wb.tierFirst[wb.tier[q] + 1];
  for (int t = 1; t <= kTierGeneral + 1; ++t) wb.tierFirst[t] += wb.tierFirst[t - 1];
  wb.windows.assign(winBytes + 16, 0);
  tmw.lap("  wide: offsets + tables");
  // pass B (parallel): fill
  pool_->parallelFor((int) ((n + blockA - 1) / blockA), [&](int t, int) {
    for (size_t q = (size_t) t * blockA; q < std::min(n, (size_t) (t + 1) * blockA); ++q) {
      const int k = wb.ids[q].seq, j = wb.ids[q].j;
      const RegRange& g = pc.rr[k][j];
      RegionDesc& d = wb.descs[q];
      if (wb.tier[q] >= kTierMid) {   // (the CTA tiers derive their own record layout, vce_cta.h ctaLayout)
        d.qMax = std::max(1, std::max(d.cCount, d.bCount));
        d.decBits = (mo + 2 <= 15) ? 4 : 8;
        PathLayout PL;
        PL.init(pc.H, d.qMax, std::max(1, std::max(d.cCount, d.bCount)), d.cCount + d.bCount + 2, d.decBits);
        d.layoutW = PL.W;
      }
      if (d.winLen > 0) std::memcpy(wb.windows.data() + d.winOff, seqs_[k].template_bases + d.winStart, (size_t) d.winLen);
      const int first[2] = {g.cFirst, g.bFirst}, count[2] = {g.cCount, g.bCount};
      for (int side = 0; side < 2; ++side) {
        const SideSeq& s = pc.ss[side][k];
        const vce_variants& in = *s.in;
        HostSide& hs = pd.side[side];
        AlleleTable& tb = wb.at[side];
        int vo = vOff[side][q], slo = sOff[side][q], no = nOff[side][q];
        for (int v = first[side]; v < first[side] + count[side]; ++v, ++vo) {
          const int li = v - s.first;
          const int ii = s.idx ? s.idx[li] : li;
          hs.vStart[vo] = pc.vStart[side][v];
          hs.vEnd[vo] = pc.vEnd[side][v];
          hs.slot0[vo] = slo;
          const int off0 = in.allele_off[ii], nSl = in.allele_off[ii + 1] - off0;
          hs.nSl[vo] = nSl;
          uint8_t f = (in.phased && in.phased[ii]) ? 1 : 0;
          if (in.status_in) f |= in.status_in[ii] & VCE_STATUS_OUTSIDE_EVAL;
          hs.vFlags[vo] = f;
          for (int h = 0; h < 4; ++h) hs.gt[(size_t) vo * 4 + h] = h < in.gt_ploidy ? in.gt[(size_t) ii * in.gt_ploidy + h] : (int8_t) -2;
          const int nb0 = in.allele_nt_off[off0];
          for (int sl = off0; sl < off0 + nSl; ++sl, ++slo) {
            tb.aStart[slo] = in.allele_start[sl];
            tb.aEnd[slo] = in.allele_end[sl];
            tb.aNull[slo] = in.allele_null[sl];
            tb.aNtOff[slo] = no + (in.allele_nt_off[sl] - nb0);
          }
          const int nbytes = in.allele_nt_off[off0 + nSl] - nb0;
          if (nbytes > 0) std::memcpy(tb.nt.data() + no, in.nt + nb0, (size_t) nbytes);
          no += nbytes;
        }
      }
    }
  });
  tmw.lap("  wide: pack");
  wb.packed = true;
  return VCE_OK;
}

int Engine::wideLaunch(WideBatch& wb) {
  StageTimer tmw;
  PassData& pd = wb.pd;
  int rc;
  for (int side = 0; side < 2; ++side) {
    if ((rc = be_->setAlleles(side, wb.at[side]))) { err_ = be_->lastError(); return rc; }
  }
  if ((rc = be_->uploadPass(pd))) { err_ = be_->lastError(); return rc; }
  if ((rc = be_->orient(pd))) { err_ = be_->lastError(); return rc; }
  const SearchConfig sc{cfg_.max_paths, cfg_.max_iterations, cfg_.prune_no_ops, cfg_.preference};
  wb.searchBefore = be_->stats.searchMs;
  if ((rc = be_->searchStart(pd, sc, wb.descs, wb.windows, wb.tierFirst, wb.syncOff, wb.so))) { err_ = be_->lastError(); return rc; }
  wb.started = true;
  tmw.lap("  wide: upload + launch");
  return VCE_OK;
}

// Waits for the batch, extracts the settled regions' results; statusOut[q] is the kRegion* status of wb.ids[q].
int Engine::wideFinish(PassCtx& pc, WideBatch& wb, std::vector<int>& statusOut) {
  statusOut.assign(wb.ids.size(), kRegionPending);
  if (!wb.started) return VCE_OK;
  StageTimer tmw;
  int rc;
  PassData& pd = wb.pd;
  const std::vector<RegKey>& ids = wb.ids;
  std::vector<RegionResult> out;
  std::vector<WarnRec> warns;
  if ((rc = be_->searchFinish(out, warns))) { err_ = be_->lastError(); return rc; }
  if (tmw.on) {
    std::unordered_map<int, int> shapes;
    for (size_t q = 0; q < ids.size(); ++q) ++shapes[wb.descs[q].cCount * 100 + wb.descs[q].bCount + 10000 * wb.tier[q]];
    std::vector<std::pair<int, int>> sh(shapes.begin(), shapes.end());
    std::sort(sh.begin(), sh.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.second > b.second; });
    std::fprintf(stderr, "[vce]   shapes (tier: calls x baseline = regions):");
    for (size_t i = 0; i < sh.size() && i < 10; ++i) std::fprintf(stderr, " %d:%dx%d=%d", sh[i].first / 10000, (sh[i].first % 10000) / 100, sh[i].first % 100, sh[i].second);
    std::fprintf(stderr, "\n");
    for (int t = 0; t <= kTierGeneral; ++t) {
      long long it = 0, nv = 0;
      int maxIt = 0, maxF = 0;
      for (int q = wb.tierFirst[t]; q < wb.tierFirst[t + 1]; ++q) {
        it += out[q].iterations;
        nv += wb.descs[q].cCount + wb.descs[q].bCount;
        maxIt = std::max(maxIt, out[q].iterations);
        maxF = std::max(maxF, out[q].maxFrontier);
      }
      if (wb.tierFirst[t + 1] > wb.tierFirst[t]) {
        std::fprintf(stderr, "[vce]   wide tier %d: %d regions, %lld variants, %lld iterations (longest %d, largest frontier %d)\n", t,
                     wb.tierFirst[t + 1] - wb.tierFirst[t], nv, it, maxIt, maxF);
      }
    }
    std::fprintf(stderr, "[vce]   wide kernels: %.2f ms\n", be_->stats.searchMs - wb.searchBefore);
  }
  tmw.lap("  wide: wait");
  std::vector<int32_t> syncBuf(wb.so, 0);
  if ((rc = be_->fetchPass(pd, syncBuf))) { err_ = be_->lastError(); return rc; }
  tmw.lap("  wide: fetch");
  // a re-run region reports its warnings afresh
  if (!pc.warns.empty()) {
    std::vector<WarnEntry> keep;
    for (const WarnEntry& w : pc.warns) {
      bool redo = false;
      for (const RegKey& id : ids) if (id.seq == w.seq && id.j == w.j) { redo = true; break; }
      if (!redo) keep.push_back(w);
    }
    pc.warns.swap(keep);
  }
  // pass 1 (serial, sequential reads): statuses, the rare bookkeeping of unsettled regions, where each settled region's
  // sync points go; pass 2 (parallel): the scattered writes into the per-region / per-variant arrays
  std::vector<int32_t> syncIdx(ids.size(), -1);
  size_t syncAt = pc.wideSync.size();
  for (size_t q = 0; q < ids.size(); ++q) {
    const RegionResult& r = out[q];
    statusOut[q] = r.status;
    be_->stats.iterations += r.iterations;
    if (r.status == kRegionClean || r.status == kRegionFinished) {
      syncIdx[q] = (int32_t) syncAt;
      syncAt += (size_t) r.nSync;
    } else {
      Rare& ra = pc.rare[key(ids[q].seq, ids[q].j)];
      ra.tier = wb.tier[q];
      ra.lastStatus = r.status;
    }
  }
  pc.wideSync.resize(syncAt);
  const size_t blockX = 512;
  pool_->parallelFor((int) ((ids.size() + blockX - 1) / blockX), [&](int t, int) {
    for (size_t q = (size_t) t * blockX; q < std::min(ids.size(), (size_t) (t + 1) * blockX); ++q) {
      const int k = ids[q].seq, j = ids[q].j;
      const RegRange& g = pc.rr[k][j];
      RegRes& s = pc.rs[k][j];
      const RegionResult& r = out[q];
      const RegionDesc& d = wb.descs[q];
      const int tier = wb.tier[q];
      if (r.status == kRegionClean || r.status == kRegionFinished) {
        s.state = RS_WIDE_DONE;
        s.flags = (uint8_t) ((r.usedPrefix & 1) | ((r.usedPrefix & 2) ? 0x20 : 0) | (r.status == kRegionFinished ? 2 : 0) | (tier << 2));
        s.lastId = r.lastBaseAlleleId == kPrefixEmpty ? (int8_t) kMicroNoLastId : (int8_t) r.lastBaseAlleleId;
        s.nSync = 0;
        s.ext.idx = syncIdx[q];
        s.ext.n = r.nSync;
        std::copy(syncBuf.begin() + wb.syncOff[q], syncBuf.begin() + wb.syncOff[q] + r.nSync, pc.wideSync.begin() + syncIdx[q]);
        for (int i = 0; i < g.cCount; ++i) {
          pc.decision[0][g.cFirst + i] = pd.side[0].decision[d.cFirst + i];
          pc.weight[g.cFirst + i] = pd.side[0].weight[d.cFirst + i];
        }
        for (int i = 0; i < g.bCount; ++i) pc.decision[1][g.bFirst + i] = pd.side[1].decision[d.bFirst + i];
      } else if (r.status == kRegionErrNull) {
        s.state = RS_ERR_NULL;
      } else if (r.status == kRegionErrMove) {
        s.state = RS_ERR_MOVE;
      } else if (r.status == kRegionErrOrder) {
        s.state = RS_ERR_ORDER;
      } else {
        s.state = RS_RESIDUAL;
      }
    }
  });
  for (const WarnRec& w : warns) {
    if (w.region < 0 || (size_t) w.region >= ids.size()) continue;
    const int st = statusOut[w.region];
    if (st != kRegionClean && st != kRegionFinished) continue;
    pc.warns.push_back(WarnEntry{ids[w.region].seq, ids[w.region].j, w});
  }
  tmw.lap("  wide: extract");
  wb.started = false;
  return VCE_OK;
}

// Settles every region that is still open: the regions started early in the warp-per-region kernels (`early`), the
// ones the thread-per-region kernel handed back, then SPILL -> merge with the following region and re-run, capacity ->
// next tier, and finally the prefix-dependent tie-breaks (MaxSumBoth.java:55,76).
int Engine::settle(PassCtx& pc, WideBatch& early) {
  StageTimer tms;
  int rc;
  std::vector<std::pair<RegKey, int>> todo;   // (region, tier)
  {
    std::vector<std::vector<RegKey>> per(nSeq_);
    pool_->parallelFor(nSeq_, [&](int k, int) {
      const RegRes* rs = pc.rs[k].data();
      const size_t n = pc.rs[k].size();
      for (size_t j = 0; j < n; ++j) if (rs[j].state == RS_RESIDUAL) per[k].push_back(RegKey{k, (int) j});
    });
    for (int k = 0; k < nSeq_; ++k) {
      for (const RegKey& id : per[k]) todo.push_back(std::make_pair(id, startTier(pc, pc.rr[k][id.j])));
    }
  }
  be_->stats.escalated += (int64_t) (todo.size() + early.ids.size());
  auto nextAlive = [&](int k, int j) {
    const int n = (int) pc.rs[k].size();
    int nx = j + 1;
    while (nx < n && pc.rs[k][nx].state == RS_DEAD) ++nx;
    return nx;
  };
  auto unmark = [&](int k, int j) {  // the region's early-assembled outputs are no longer final
    if (!earlyUsed_ || pc.pass != 0) return;
    const RegRange& g = pc.rr[k][j];
    for (int v = g.cFirst; v < g.cFirst + g.cCount; ++v) earlyDone_[0][v] = 0;
    for (int v = g.bFirst; v < g.bFirst + g.bCount; ++v) earlyDone_[1][v] = 0;
  };
  std::vector<RegKey> retry;  // merged regions: their new shape may fit the thread-per-region kernel
  std::vector<std::pair<RegKey, int>> all;  // (region, status) to inspect
  {
    std::vector<int> st;
    if ((rc = wideFinish(pc, early, st))) return rc;
    for (size_t q = 0; q < early.ids.size(); ++q) all.push_back(std::make_pair(early.ids[q], st[q]));
  }
  tms.lap("  settle: early batch collected");
  bool settled = false;
  for (int round = 0; round < 100000; ++round) {
    if (!retry.empty()) {
      std::sort(retry.begin(), retry.end(), [](const RegKey& a, const RegKey& b) { return a.seq != b.seq ? a.seq < b.seq : a.j < b.j; });
      std::vector<Chunk> rch;
      for (const RegKey& id : retry) {
        if (rch.empty() || rch.back().seq != id.seq) {
          rch.emplace_back();
          rch.back().seq = id.seq;
          rch.back().transient = true;
        }
        rch.back().list.push_back(id.j);
      }
      // build every sequence's packets, then ONE launch per packet class over all of them
      for (Chunk& ch : rch) if ((rc = buildChunk(pc, ch, 0, false))) return rc;
      for (int c = 0; c < 2; ++c) {
        std::vector<MicroJob*> jobs;
        for (Chunk& ch : rch) if (ch.job[c].n > 0) jobs.push_back(&ch.job[c]);
        if (!jobs.empty() && (rc = be_->microRunAll(jobs, pc.mc))) { err_ = be_->lastError(); return rc; }
      }
      int64_t it = 0;
      for (Chunk& ch : rch) {
        for (int c = 0; c < 2; ++c) if (ch.job[c].n > 0 && (rc = be_->microWait(ch.job[c]))) { err_ = be_->lastError(); return rc; }
        decodeChunk(pc, ch, it);
      }
      be_->stats.iterations += it;
      for (const RegKey& id : retry) {
        if (pc.rs[id.seq][id.j].state == RS_RESIDUAL) todo.push_back(std::make_pair(id, startTier(pc, pc.rr[id.seq][id.j])));
      }
      retry.clear();
      tms.lap("  settle: merged regions through the packet kernel");
    }
    if (todo.empty() && all.empty()) {
      // prefix-dependent tie-breaks: re-run the first region per sequence whose assumption was wrong
      // Every region of a sequence whose assumption does not hold is collected in one round (the scan goes on with the
      // region's current last allele id; should the re-run change it, the next round's scan catches what follows).
      std::vector<std::vector<std::pair<int, int>>> fix(nSeq_);   // per sequence: (region, true prefix)
      pool_->parallelFor(nSeq_, [&](int k, int) {
        const RegRes* rs = pc.rs[k].data();
        const int n = (int) pc.rs[k].size();
        int prefix = kPrefixEmpty;
        for (int j = 0; j < n; ++j) {
          const RegRes& s = rs[j];
          if (s.state != RS_MICRO_DONE && s.state != RS_WIDE_DONE) continue;
          if (s.flags & 1) {
            int assumed = j == 0 ? kPrefixEmpty : kDummyPrefix;
            if (s.state == RS_WIDE_DONE) {
              const auto it = pc.rare.find(key(k, j));
              if (it != pc.rare.end() && it->second.hasPrefix) assumed = it->second.assumedPrefix;
            }
            // wrong only if the emptiness differs, or -- when the tie-break got as far as comparing allele ids -- the id does
            const bool emptyDiffers = (assumed == kPrefixEmpty) != (prefix == kPrefixEmpty);
            if (emptyDiffers || ((s.flags & 0x20) && assumed != prefix)) fix[k].push_back(std::make_pair(j, prefix));
          }
          if (s.lastId != (int8_t) kMicroNoLastId) prefix = s.lastId;
        }
      });
      for (int k = 0; k < nSeq_; ++k) {
        for (const std::pair<int, int>& fx : fix[k]) {
          const RegRes& s = pc.rs[k][fx.first];
          Rare& ra = pc.rare[key(k, fx.first)];
          ra.hasPrefix = true;
          ra.assumedPrefix = fx.second;
          unmark(k, fx.first);
          int tier = startTier(pc, pc.rr[k][fx.first]);
          if (s.state == RS_WIDE_DONE) tier = std::max(tier, (int) (s.flags >> 2) & 7);  // the tier that settled it before
          todo.push_back(std::make_pair(RegKey{k, fx.first}, tier));
          ++be_->stats.prefixReruns;
        }
      }
      tms.lap("  settle: prefix scan");
      if (todo.empty()) { settled = true; break; }
    }
    // one batch with every tier's regions, then inspect in sequence order so that chains of spilling regions merge
    // front to back
    if (!todo.empty()) {
      WideBatch wb;
      std::vector<int> st;
      if ((rc = wideStart(pc, todo, wb))) return rc;
      if ((rc = wideFinish(pc, wb, st))) return rc;
      for (size_t q = 0; q < wb.ids.size(); ++q) all.push_back(std::make_pair(wb.ids[q], st[q]));
      todo.clear();
    }
    std::sort(all.begin(), all.end(), [](const std::pair<RegKey, int>& a, const std::pair<RegKey, int>& b) {
      return a.first.seq != b.first.seq ? a.first.seq < b.first.seq : a.first.j < b.first.j;
    });
    for (size_t q = 0; q < all.size(); ++q) {
      const int k = all[q].first.seq, j = all[q].first.j;
      const int st = all[q].second;
      if (pc.rs[k][j].state == RS_DEAD) continue;
      if (st == kRegionOverflow) {
        Rare& ra = pc.rare[key(k, j)];
        if (ra.tier >= kTierGeneral) { err_ = "search capacity exceeded in the large-capacity kernel"; return VCE_ERR_CAPACITY; }
        // what outgrows a shared-memory tier goes straight to the arena tiers: every settle round is a launch plus a host
        // round trip (~1 ms), and the follow-up rounds hold a handful of regions
        const int nt = ra.tier < kTierMid ? kTierMid : ra.tier + 1;
        todo.push_back(std::make_pair(all[q].first, nt));
        ++be_->stats.escalated;
      } else if (st == kRegionWinMiss) {
        // a haplotype ran off the reference window (long repeats keep paths apart for longer than the margin): the same
        // region again with a wider window -- merging with the next region would drag that region's variants in for nothing
        Rare& ra = pc.rare[key(k, j)];
        if (ra.extraMargin < 0) { err_ = "window miss with the whole template in the window"; return VCE_ERR_CAPACITY; }
        ra.extraMargin = ra.extraMargin == 0 ? 1024 : (ra.extraMargin >= (1 << 20) ? -1 : ra.extraMargin * 16);
        todo.push_back(std::make_pair(all[q].first, std::max(ra.tier, startTier(pc, pc.rr[k][j]))));
        pc.rs[k][j].state = RS_RESIDUAL;
      } else if (st == kRegionSpill) {
        Rare& ra = pc.rare[key(k, j)];
        ++be_->stats.spilled;
        const int n = (int) pc.rs[k].size();
        int nx = nextAlive(k, j);
        RegRange& g = pc.rr[k][j];
        if (nx >= n) {
          // last region of its sequence ran off its reference window: widen it
          if (ra.extraMargin < 0) { err_ = "region spilled with the whole template in its window"; return VCE_ERR_CAPACITY; }
          ra.extraMargin = ra.extraMargin == 0 ? 256 : (ra.extraMargin >= (1 << 20) ? -1 : ra.extraMargin * 16);
          todo.push_back(std::make_pair(all[q].first, std::max(ra.tier, startTier(pc, g))));
        } else {
          // absorb the following region (and further ones while they are themselves known to spill)
          pc.undo.push_back(std::make_pair(RegKey{k, j}, g));
          bool more = true;
          while (more && nx < n) {
            RegRes& o = pc.rs[k][nx];
            const RegRange& og = pc.rr[k][nx];
            const auto oit = pc.rare.find(key(k, nx));
            more = o.state == RS_RESIDUAL && oit != pc.rare.end() && oit->second.lastStatus == kRegionSpill;
            unmark(k, nx);
            g.cCount += og.cCount;
            g.bCount += og.bCount;
            o.state = RS_DEAD;
            nx = nextAlive(k, nx);
          }
          ra.extraMargin = 0;
          if (pc.microOk && !ra.microTried && !ra.hasPrefix) {
            ra.microTried = true;
            retry.push_back(all[q].first);
          } else {
            todo.push_back(std::make_pair(all[q].first, startTier(pc, g)));
          }
        }
        pc.rs[k][j].state = RS_RESIDUAL;
      }
    }
    all.clear();
    tms.lap("  settle: statuses");
  }
  if (!settled) { err_ = "regions left unsettled after the round limit"; return VCE_ERR_CAPACITY; }
  return VCE_OK;
}

// ------------------------------------------------------------------------------------------- assembly
// PhasingEvaluator.countMisphasings, src/com/rtg/vcf/eval/PhasingEvaluator.java:55-142, on the flat pass-0 results.
namespace {
struct VSum { int start; bool phased, included, phase; };
}  // namespace

static inline void pickRow(int kind, int H, int gtPloidy, const vce_variants& in, int ii, int row, int8_t* ids, uint8_t* original);

namespace {
// PhasingEvaluator.CallIterator :160-195 streamed over the flat results of one side: merges the included and the
// excluded variants by start, included first on ties.
struct CallIter {
  const uint8_t* dec;      // decisions of the side's variants of this sequence (sorted order)
  const int32_t* start;
  const vce_variants& in;
  const int32_t* idx;
  int n, kind, H, gtPloidy;
  int i = 0, e = 0;        // next included / excluded candidate
  CallIter(const uint8_t* d, const int32_t* st, const vce_variants& v, const int32_t* ix, int n_, int kind_, int H_, int gp)
      : dec(d), start(st), in(v), idx(ix), n(n_), kind(kind_), H(H_), gtPloidy(gp) {
    while (i < n && dec[i] < 2) ++i;
    while (e < n && dec[e] != 1) ++e;
  }
  bool hasNext() const { return i < n || e < n; }
  VSum next() {
    const bool haveI = i < n, haveE = e < n;
    if (!haveI || (haveE && start[e] < start[i])) {
      const int li = e;
      do { ++e; } while (e < n && dec[e] != 1);
      const int ii = idx ? idx[li] : li;
      return VSum{start[li], in.phased && in.phased[ii], false, false};
    }
    const int li = i;
    do { ++i; } while (i < n && dec[i] < 2);
    const int ii = idx ? idx[li] : li;
    const bool ph = in.phased && in.phased[ii];
    // OrientedVariant.isOriginal() of the chosen orientation (Orientor.java): only the diploid unphased expansion
    // (first row) and the phased orientor (as-is) produce "original" orientations
    bool orig = false;
    const int row = dec[li] - 2;
    if (kind == VCE_ORIENT_PHASED && ph) orig = true;
    else if (kind == VCE_ORIENT_PHASED_INVERT && ph) orig = false;
    else if (kind == VCE_ORIENT_UNPHASED || kind == VCE_ORIENT_PHASED || kind == VCE_ORIENT_PHASED_INVERT) orig = H == 2 && row == 0;
    return VSum{start[li], ph, true, orig};
  }
};
}  // namespace

void Engine::phasing(int k, const std::vector<int32_t>& sync, int32_t out[3]) const {
  const PassCtx& pc = pass_[0];
  const SideSeq& sc = pc.ss[0][k];
  const SideSeq& sb = pc.ss[1][k];
  CallIter baseline(pc.decision[1].data() + sb.first, pc.vStart[1].data() + sb.first, *sb.in, sb.idx, sb.n, pc.kind[1], pc.H, pc.gtPloidy[1]);
  CallIter calls(pc.decision[0].data() + sc.first, pc.vStart[0].data() + sc.first, *sc.in, sc.idx, sc.n, pc.kind[0], pc.H, pc.gtPloidy[0]);
  int mis = 0, unph = 0, correct = 0;
  bool baseIsPhased = false, basePhase = false, callIsPhased = false, callPhase = false;
  bool haveCall = false, haveBase = false;
  VSum call{0, false, false, false}, base{0, false, false, false};
  // One pass per sync region, no lists: the baseline section only matters through "all phased the same way" (:143-155)
  // and its phase, and that is known before the calls of the section are looked at.
  for (int point : sync) {
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
        const VSum& c = call;
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

// Orientation row `row` of a variant (host side of K1 for the output arrays): allele ids + isOriginal.
// Fast paths for the orientors of the default / squash configurations, Orientor.orientations otherwise.
static inline void pickRow(int kind, int H, int gtPloidy, const vce_variants& in, int ii, int row, int8_t* ids, uint8_t* original) {
  const int P = in.gt_ploidy;
  const bool phased = in.phased && in.phased[ii];
  if (H == 2 && P == 2 && gtPloidy == 2 && (kind == VCE_ORIENT_UNPHASED || ((kind == VCE_ORIENT_PHASED || kind == VCE_ORIENT_PHASED_INVERT) && !phased))) {
    const int8_t g0 = in.gt[(size_t) ii * 2], g1 = in.gt[(size_t) ii * 2 + 1];  // UnphasedOrientor :77-89
    ids[0] = row == 0 ? g0 : g1;
    ids[1] = row == 0 ? g1 : g0;
    ids[2] = ids[3] = -2;
    *original = row == 0 ? 1 : 0;
    return;
  }
  VariantGt g;
  g.gtPloidy = gtPloidy;
  for (int h = 0; h < 4; ++h) g.gt[h] = h < P ? in.gt[(size_t) ii * P + h] : -2;
  g.phased = phased ? 1 : 0;
  g.slot0 = in.allele_off[ii];
  g.nSlots = in.allele_off[ii + 1] - in.allele_off[ii];
  g.aNull = in.allele_null;
  RowPick pick;
  pick.want = row;
  orientVariant(kind, H, g, pick);
  for (int h = 0; h < 4; ++h) ids[h] = pick.ids[h];
  *original = pick.original ? 1 : 0;
}

// Per sequence: sync points of every pass (Path.getSyncPoints with the Path.java:293 de-duplication), errors,
// warnings, and the pass-2 decisions by input index.  Returns false when the sequence needs no per-variant pass
// (short-circuit or error: its outputs are complete).
bool Engine::assembleHeader(int k, vce_sequence_result& res) {
  const vce_sequence& sq = seqs_[k];
  res.n_sync[0] = res.n_sync[1] = 0;
  res.phasing[0] = res.phasing[1] = res.phasing[2] = 0;
  res.n_warnings = 0;
  res.error = VCE_OK;
  res.n_regions = 0;
  res.n_iterations = 0;
  vce_side_result* sides[2] = {&res.calls, &res.baseline};
  const vce_variants* vin[2] = {&sq.calls, &sq.baseline};
  auto blank = [&](int extraStatus, int syncId) {
    for (int side = 0; side < 2; ++side) {
      vce_side_result& o = *sides[side];
      for (int i = 0; i < vin[side]->n; ++i) {
        o.status[i] = (uint8_t) ((vin[side]->status_in ? vin[side]->status_in[i] : 0) | extraStatus);
        for (int h = 0; h < VCE_MAX_PLOIDY; ++h) o.allele_ids[(size_t) i * VCE_MAX_PLOIDY + h] = -2;
        o.original[i] = 0;
        o.weight[i] = 0;
        if (o.sync_id) o.sync_id[i] = syncId;
      }
    }
  };
  if (pass_[0].shortCircuit[k]) {  // SequenceEvaluator.java:94-97
    blank(VCE_STATUS_NO_MATCH, 1);
    return false;
  }
  for (int p = 0; p < 2; ++p) seqSync_[p][k].clear();
  for (int p = 0; p < cfg_.n_passes; ++p) {
    const PassCtx& pc = pass_[p];
    std::vector<int32_t>& sync = seqSync_[p][k];
    // the sync points of the sequence's region blocks were gathered in parallel (finish()); join them here
    size_t total = 0;
    for (int t = syncPartFirst_[p][k]; t < syncPartFirst_[p][k + 1]; ++t) total += syncParts_[t].v.size();
    sync.reserve(total + 8);
    for (int t = syncPartFirst_[p][k]; t < syncPartFirst_[p][k + 1]; ++t) {
      const SyncPart& sp = syncParts_[t];
      if (sp.err) res.error = sp.err;
      size_t from = 0;
      if (!sync.empty() && !sp.v.empty() && sp.v[0] == sync.back()) from = 1;  // Path.java:293 de-duplication across the joint
      sync.insert(sync.end(), sp.v.begin() + (long) from, sp.v.end());
    }
    // re-runs append out of order: report in region order, as the sequential reference would
    std::vector<WarnEntry> ws;
    for (const WarnEntry& w : pc.warns) if (w.seq == k && pc.rs[k][w.j].state != RS_DEAD) ws.push_back(w);
    std::stable_sort(ws.begin(), ws.end(), [](const WarnEntry& a, const WarnEntry& b) { return a.j < b.j; });
    for (const WarnEntry& w : ws) {
      if (res.n_warnings < res.warn_cap) res.warnings[res.n_warnings] = vce_warning{p, w.w.paths, w.w.iterations, w.w.start1, w.w.end1};
      ++res.n_warnings;
    }
  }
  if (res.error != VCE_OK) {
    blank(0, 0);
    return false;
  }
  for (int p = 0; p < cfg_.n_passes; ++p) {
    const std::vector<int32_t>& sync = seqSync_[p][k];
    res.n_sync[p] = (int32_t) sync.size();
    if ((int) sync.size() > res.sync_cap[p]) { res.error = VCE_ERR_CAPACITY; blank(0, 0); return false; }
    std::copy(sync.begin(), sync.end(), res.sync_points[p]);
  }
  res.n_regions = (int64_t) seqSync_[0][k].size();
  if (cfg_.n_passes == 2) {  // pass-2 decision / weight of every input variant (0 when the variant did not enter pass 2)
    const PassCtx& p1 = pass_[1];
    for (int side = 0; side < 2; ++side) {
      const SideSeq& b = p1.ss[side][k];
      std::vector<uint8_t>& d1 = seqDec1_[side][k];
      d1.assign((size_t) vin[side]->n, 0);
      std::vector<double>& w1 = seqW1_[k];
      if (side == 0) w1.assign((size_t) vin[side]->n, 0.0);
      for (int li = 0; li < b.n; ++li) {
        const int ii = b.idx[li];
        d1[ii] = p1.decision[side][b.first + li];
        if (side == 0) w1[ii] = p1.weight[b.first + li];
      }
    }
  }
  return true;
}

// Statuses (SequenceEvaluator.mergeVariants :170-182, insertSkipped :190-209), orientation ids, weights and sync ids
// of the variants [lo, hi) (sorted positions of pass 1) of one side of sequence k.
void Engine::assembleBlock(int k, int side, int lo, int hi, vce_sequence_result& res) const {
  const vce_sequence& sq = seqs_[k];
  vce_side_result& o = side == 0 ? res.calls : res.baseline;
  const vce_variants& in = side == 0 ? sq.calls : sq.baseline;
  const PassCtx& p0 = pass_[0];
  const SideSeq& a = p0.ss[side][k];
  const bool two = cfg_.n_passes == 2;
  const uint8_t* dec1 = two ? seqDec1_[side][k].data() : nullptr;
  const double* w1 = two && side == 0 ? seqW1_[k].data() : nullptr;
  const std::vector<int32_t>& s0 = seqSync_[0][k];
  const std::vector<int32_t>& s1 = seqSync_[1][k];
  // InterleavingEvalSynchronizer.floorSyncPos :71-83: last sync point <= start - 1 (the first one if none); variants
  // come in ascending start order, so the floor index only moves forward
  auto floorInit = [](const std::vector<int32_t>& sp, int start) -> size_t {
    if (sp.empty()) return 0;
    const size_t u = (size_t) (std::upper_bound(sp.begin(), sp.end(), start - 1) - sp.begin());
    return u == 0 ? 0 : u - 1;
  };
  size_t f0 = 0, f1 = 0;
  if (lo < hi) {
    f0 = floorInit(s0, p0.vStart[side][a.first + lo]);
    f1 = floorInit(s1, p0.vStart[side][a.first + lo]);
  }
  auto floorVal = [](const std::vector<int32_t>& sp, size_t& cur, int start) -> int {
    if (sp.empty()) return 0;
    const int vv = start - 1;
    while (cur + 1 < sp.size() && sp[cur + 1] <= vv) ++cur;
    return sp[cur] + 1;
  };
  const uint8_t* early = earlyUsed_ ? earlyDone_[side].data() : nullptr;
  for (int li = lo; li < hi; ++li) {
    const int v = a.first + li;
    if (early != nullptr && early[v]) continue;  // written while the region's record was decoded
    const int ii = a.idx ? a.idx[li] : li;
    int st = in.status_in ? in.status_in[ii] : 0;
    const int d0 = p0.decision[side][v];
    int passSel = -1, row = 0;
    double wgt = 0.0;
    if (d0 >= 2) {
      st |= VCE_STATUS_GT_MATCH;
      passSel = 0;
      row = d0 - 2;
      wgt = side == 0 ? p0.weight[v] : 0.0;
    } else if (!two) {
      st |= d0 == 1 ? VCE_STATUS_NO_MATCH : VCE_STATUS_SKIPPED;
    } else {
      const int d1 = dec1[ii];
      if (d1 >= 2) {
        st |= VCE_STATUS_ALLELE_MATCH;
        passSel = 1;
        row = d1 - 2;
        wgt = side == 0 ? w1[ii] : 0.0;
      } else {
        st |= d1 == 1 ? VCE_STATUS_NO_MATCH : VCE_STATUS_SKIPPED;
      }
    }
    o.status[ii] = (uint8_t) st;
    int8_t* ids = o.allele_ids + (size_t) ii * VCE_MAX_PLOIDY;
    if (passSel >= 0) {
      const PassCtx& pp = pass_[passSel];
      pickRow(pp.kind[side], pp.H, pp.gtPloidy[side], in, ii, row, ids, &o.original[ii]);
    } else {
      ids[0] = ids[1] = ids[2] = ids[3] = -2;
      o.original[ii] = 0;
    }
    o.weight[ii] = wgt;
    if (o.sync_id) {
      // SYNC annotation value (InterleavingEvalSynchronizer.floorSyncPos :71-83 + WithInfoEvalSynchronizer :121-222)
      const int sv1 = floorVal(s0, f0, p0.vStart[side][v]);
      const int sv2 = floorVal(s1, f1, p0.vStart[side][v]);
      int val = 0;
      if (st & VCE_STATUS_SKIPPED) val = 0;
      else if (st & VCE_STATUS_GT_MATCH) val = sv1 + 1;
      else if (st & VCE_STATUS_ALLELE_MATCH) val = sv2 + 1;
      else if (st & VCE_STATUS_NO_MATCH) val = sv2 > 0 ? sv2 + 1 : sv1 + 1;
      o.sync_id[ii] = val;
    }
  }
}

// ------------------------------------------------------------------------------------------- device-resident results
// Single-pass vce_submit on sorted inputs with the thread-per-region kernel in play: its result records stay on the device.
bool Engine::wantDeviceResults(const PassCtx&) const {
  if (!opt_.deviceResults || !opt_.deviceBuild || opt_.disableMicro || opt_.forceGeneral || !be_->resultsSupported()) return false;
  if (cfg_.n_passes != 1) return false;
  if (!orientorPairMicro(cfg_.calls_orientor[0], cfg_.baseline_orientor[0], cfg_.ploidy[0])) return false;
  const PassCtx& pc = pass_[0];
  for (int k = 0; k < nSeq_; ++k) {
    if (pc.shortCircuit[k]) continue;
    if (pc.ss[0][k].idx != nullptr || pc.ss[1][k].idx != nullptr) return false;   // an unsorted side: host path
  }
  return true;
}

int Engine::resultsSetup(PassCtx& pc) {
  regBase_.assign((size_t) nSeq_ + 1, 0);
  for (int k = 0; k < nSeq_; ++k) regBase_[(size_t) k + 1] = regBase_[(size_t) k] + (int32_t) pc.rr[k].size();
  return VCE_OK;
}

// The region table's info words come back (4 bytes per region instead of a 16 / 32-byte record to parse): which regions the
// thread-per-region kernel settled, whether their tie-breaks looked at the prefix, their last baseline allele id.
int Engine::resultsCollect(PassCtx& pc) {
  const uint32_t* info = nullptr;
  int64_t iterations = 0;
  const int rc = be_->resultsRegInfo(&info, &iterations);
  if (rc) { err_ = be_->lastError(); return rc; }
  be_->stats.iterations += iterations;
  struct Task { int k, lo, hi; };
  std::vector<Task> tasks;
  const int block = 1 << 16;
  for (int k = 0; k < nSeq_; ++k)
    for (int lo = 0; lo < (int) pc.rr[k].size(); lo += block) tasks.push_back(Task{k, lo, std::min((int) pc.rr[k].size(), lo + block)});
  pool_->parallelFor((int) tasks.size(), [&](int t, int) {
    const Task& tk = tasks[t];
    RegRes* rs = pc.rs[tk.k].data();
    const uint8_t* cls = pc.rcls[tk.k].data();
    const uint32_t* in = info + regBase_[tk.k];
    for (int j = tk.lo; j < tk.hi; ++j) {
      if (cls[j] != 1 && cls[j] != 2) continue;           // not a packet region of this call
      if (rs[j].state != RS_PENDING) continue;
      const uint32_t v = in[j];
      if (riStatus(v) == RI_CLEAN) {
        rs[j].state = RS_MICRO_DONE;
        rs[j].flags = (uint8_t) ((riFlags(v) & 1u) | ((riFlags(v) & 2u) ? 0x20u : 0u) | 0x80u);   // bit 7: the region's results live on the device
        rs[j].lastId = (int8_t) riLastId(v);
        rs[j].nSync = 0;
      } else {
        rs[j].state = RS_RESIDUAL;
      }
    }
  });
  return VCE_OK;
}

int Engine::finishDevice(vce_sequence_result* results) {
  StageTimer tm;
  PassCtx& pc = pass_[0];
  // ---- the regions whose final state is the host's
  struct Part { std::vector<int32_t> regIdx, syncCnt, sync; std::vector<uint32_t> info; std::vector<RegRange> range; std::vector<uint8_t> dec;
                std::vector<double> w; int err = 0; };
  std::vector<Part> parts((size_t) nSeq_);
  pool_->parallelFor(nSeq_, [&](int k, int) {
    Part& pt = parts[(size_t) k];
    if (pc.shortCircuit[k]) return;
    const RegRes* rs = pc.rs[k].data();
    const int n = (int) pc.rs[k].size();
    for (int j = 0; j < n; ++j) {
      const RegRes& s = rs[j];
      if (s.state == RS_MICRO_DONE && (s.flags & 0x80u)) continue;
      const RegRange& g = pc.rr[k][j];
      const bool alive = s.state == RS_MICRO_DONE || s.state == RS_WIDE_DONE;
      if (s.state == RS_ERR_NULL) pt.err = VCE_ERR_BEST_PATH_NULL;
      else if (s.state == RS_ERR_MOVE) pt.err = VCE_ERR_MOVE_IN_VARIANT;
      else if (s.state == RS_ERR_ORDER) pt.err = VCE_ERR_UNSUPPORTED;
      else if (!alive && s.state != RS_DEAD && pt.err == 0) pt.err = VCE_ERR_CAPACITY;   // must not happen: an unsettled region
      pt.regIdx.push_back(regBase_[k] + j);
      int cnt = 0;
      if (s.state == RS_MICRO_DONE) {   // decoded on the host (a merged region the packet kernel took on the retry)
        const int base = regionStartPos(pc, k, j);
        for (int q = 0; q < s.nSync; ++q) pt.sync.push_back(base + s.rel[q]);
        cnt = s.nSync;
      } else if (s.state == RS_WIDE_DONE) {
        for (int q = 0; q < s.ext.n; ++q) pt.sync.push_back(pc.wideSync[(size_t) s.ext.idx + q]);
        cnt = s.ext.n;
      }
      pt.syncCnt.push_back(cnt);
      pt.info.push_back(riPack(alive ? RI_HOST : RI_DEAD, s.flags & 3u, alive ? (int) s.lastId : kMicroNoLastId));
      pt.range.push_back(alive ? g : RegRange{g.cFirst, 0, g.bFirst, 0});
      if (alive) {
        for (int i = 0; i < g.cCount; ++i) { pt.dec.push_back(pc.decision[0][g.cFirst + i]); pt.w.push_back(pc.weight[g.cFirst + i]); }
        for (int i = 0; i < g.bCount; ++i) pt.dec.push_back(pc.decision[1][g.bFirst + i]);
      }
    }
  });
  ResultsPatch patch;
  for (int k = 0; k < nSeq_; ++k) {
    const Part& pt = parts[(size_t) k];
    size_t so = patch.extSync.size(), po = patch.decPay.size(), wo = patch.wPay.size();
    size_t si = 0;
    for (size_t i = 0; i < pt.regIdx.size(); ++i) {
      patch.regIdx.push_back(pt.regIdx[i]);
      patch.info.push_back(pt.info[i]);
      patch.range.push_back(pt.range[i]);
      patch.syncOff.push_back((int32_t) (so + si));
      patch.syncCnt.push_back(pt.syncCnt[i]);
      si += (size_t) pt.syncCnt[i];
      patch.payOff.push_back((int32_t) po);
      patch.wOff.push_back((int32_t) wo);
      po += (size_t) pt.range[i].cCount + (size_t) pt.range[i].bCount;
      wo += (size_t) pt.range[i].cCount;
    }
    patch.extSync.insert(patch.extSync.end(), pt.sync.begin(), pt.sync.end());
    patch.decPay.insert(patch.decPay.end(), pt.dec.begin(), pt.dec.end());
    patch.wPay.insert(patch.wPay.end(), pt.w.begin(), pt.w.end());
  }
  tm.lap("device results: patch of host-settled regions");
  ResultsOut out;
  int rc = be_->resultsFinish(patch, out);
  if (rc) { err_ = be_->lastError(); return rc; }
  tm.lap("device results: final passes + copy back");
  // ---- per sequence: sync points, phasing counts, warnings, errors
  std::vector<uint8_t> live((size_t) nSeq_, 0);
  struct SyncCopy { int32_t* dst; const int32_t* src; size_t n; };
  std::vector<SyncCopy> syncCopies;
  for (int k = 0; k < nSeq_; ++k) {
    vce_sequence_result& res = results[k];
    const vce_sequence& sq = seqs_[k];
    res.n_sync[0] = res.n_sync[1] = 0;
    res.phasing[0] = res.phasing[1] = res.phasing[2] = 0;
    res.n_warnings = 0;
    res.error = VCE_OK;
    res.n_regions = 0;
    res.n_iterations = 0;
    vce_side_result* sides[2] = {&res.calls, &res.baseline};
    const vce_variants* vin[2] = {&sq.calls, &sq.baseline};
    auto blank = [&](int extraStatus, int syncId) {
      for (int side = 0; side < 2; ++side) {
        vce_side_result& o = *sides[side];
        for (int i = 0; i < vin[side]->n; ++i) {
          o.status[i] = (uint8_t) ((vin[side]->status_in ? vin[side]->status_in[i] : 0) | extraStatus);
          for (int h = 0; h < VCE_MAX_PLOIDY; ++h) o.allele_ids[(size_t) i * VCE_MAX_PLOIDY + h] = -2;
          o.original[i] = 0;
          o.weight[i] = 0;
          if (o.sync_id) o.sync_id[i] = syncId;
        }
      }
    };
    if (pc.shortCircuit[k]) { blank(VCE_STATUS_NO_MATCH, 1); continue; }   // SequenceEvaluator.java:94-97
    std::vector<WarnEntry> ws;
    for (const WarnEntry& w : pc.warns) if (w.seq == k && pc.rs[k][w.j].state != RS_DEAD) ws.push_back(w);
    std::stable_sort(ws.begin(), ws.end(), [](const WarnEntry& a, const WarnEntry& b) { return a.j < b.j; });
    for (const WarnEntry& w : ws) {
      if (res.n_warnings < res.warn_cap) res.warnings[res.n_warnings] = vce_warning{0, w.w.paths, w.w.iterations, w.w.start1, w.w.end1};
      ++res.n_warnings;
    }
    if (parts[(size_t) k].err) { res.error = parts[(size_t) k].err; blank(0, 0); continue; }
    const int s0 = out.seqSyncFirst[k], sn = out.seqSyncFirst[k + 1] - s0;
    res.n_sync[0] = sn;
    if (sn > res.sync_cap[0]) { res.error = VCE_ERR_CAPACITY; blank(0, 0); continue; }
    syncCopies.push_back(SyncCopy{res.sync_points[0], out.syncPos + s0, (size_t) sn});   // copied with the per-variant blocks below
    res.n_regions = sn;
    res.phasing[0] = (int32_t) out.counters[(size_t) k * 8 + 5];
    res.phasing[1] = (int32_t) out.counters[(size_t) k * 8 + 6];
    res.phasing[2] = (int32_t) out.counters[(size_t) k * 8 + 7];
    live[(size_t) k] = 1;
  }
  tm.lap("device results: headers");
  // ---- per-variant records -> the caller's arrays (sorted inputs: sorted position == input index)
  struct Task { int seq, side, lo, hi; };   // side 2: a block of sync points (seq = index into syncCopies)
  std::vector<Task> tasks;
  const int block = 1 << 15;
  for (int k = 0; k < nSeq_; ++k) {
    if (!live[(size_t) k]) continue;
    for (int side = 0; side < 2; ++side) {
      const int n = pc.ss[side][k].n;
      for (int lo = 0; lo < n; lo += block) tasks.push_back(Task{k, side, lo, std::min(n, lo + block)});
    }
  }
  const int syncBlock = 1 << 17;
  for (size_t c = 0; c < syncCopies.size(); ++c) {
    const int n = (int) syncCopies[c].n;
    for (int lo = 0; lo < n; lo += syncBlock) tasks.push_back(Task{(int) c, 2, lo, std::min(n, lo + syncBlock)});
  }
  pool_->parallelFor((int) tasks.size(), [&](int t, int) {
    const Task& tk = tasks[t];
    if (tk.side == 2) {
      const SyncCopy& sc = syncCopies[(size_t) tk.seq];
      std::memcpy(sc.dst + tk.lo, sc.src + tk.lo, (size_t) (tk.hi - tk.lo) * sizeof(int32_t));
      return;
    }
    const vce_sequence& sq = seqs_[tk.seq];
    vce_side_result& o = tk.side == 0 ? results[tk.seq].calls : results[tk.seq].baseline;
    const vce_variants& in = tk.side == 0 ? sq.calls : sq.baseline;
    const SideSeq& a = pc.ss[tk.side][tk.seq];
    const int stride = tk.side == 0 ? 16 : 8;
    const uint8_t* rec = (tk.side == 0 ? out.outCalls : out.outBase) + (size_t) (a.first + tk.lo) * stride;
    for (int li = tk.lo; li < tk.hi; ++li, rec += stride) {
      o.status[li] = rec[0];
      int8_t* ids = o.allele_ids + (size_t) li * VCE_MAX_PLOIDY;
      if (rec[1] != 0xff) {
        pickRow(pc.kind[tk.side], pc.H, pc.gtPloidy[tk.side], in, li, rec[1], ids, &o.original[li]);
      } else {
        ids[0] = ids[1] = ids[2] = ids[3] = -2;
        o.original[li] = 0;
      }
      int32_t sid;
      std::memcpy(&sid, rec + 4, 4);
      if (o.sync_id) o.sync_id[li] = sid;
      double w = 0.0;
      if (tk.side == 0) std::memcpy(&w, rec + 8, 8);
      o.weight[li] = w;
    }
  });
  tm.lap("device results: expansion into the caller's arrays");
  summary_.assign((size_t) nSeq_ * 8, 0);
  for (int k = 0; k < nSeq_; ++k) {
    if (live[(size_t) k]) for (int i = 0; i < 8; ++i) summary_[(size_t) k * 8 + i] = out.counters[(size_t) k * 8 + i];
  }
  {  // sequences that never reached the device (short-circuit, errors): counted from what was written for them
    std::vector<uint8_t> rest((size_t) nSeq_, 0);
    bool any = false;
    for (int k = 0; k < nSeq_; ++k) if (!live[(size_t) k]) { rest[(size_t) k] = 1; any = true; }
    if (any) {
      std::vector<int64_t> keep = summary_;
      summaryFromResults(results);
      for (int k = 0; k < nSeq_; ++k) if (!rest[(size_t) k]) for (int i = 0; i < 8; ++i) summary_[(size_t) k * 8 + i] = keep[(size_t) k * 8 + i];
    }
  }
  int rcAll = VCE_OK;
  for (int k = 0; k < nSeq_; ++k) if (results[k].error != VCE_OK && rcAll == VCE_OK) rcAll = results[k].error;
  return rcAll;
}

int Engine::finish(vce_sequence_result* results) {
  if (devRes_) return finishDevice(results);
  StageTimer tm;
  for (int p = 0; p < 2; ++p) ensureSize(seqSync_[p], (size_t) nSeq_);
  for (int side = 0; side < 2; ++side) ensureSize(seqDec1_[side], (size_t) nSeq_);
  ensureSize(seqW1_, (size_t) nSeq_);
  // largest sequences first
  std::vector<int> order(nSeq_);
  for (int k = 0; k < nSeq_; ++k) order[k] = k;
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    const int na = seqs_[a].calls.n + seqs_[a].baseline.n, nb = seqs_[b].calls.n + seqs_[b].baseline.n;
    return na != nb ? na > nb : a < b;
  });
  // sync points of every pass, gathered per block of regions (Path.getSyncPoints with the Path.java:293 de-duplication)
  // (the part objects persist across calls: their vectors keep their capacity, a steady-state call allocates nothing)
  const int rblock = 1 << 15;
  size_t nSyncParts = 0;
  for (int p = 0; p < 2; ++p) {
    syncPartFirst_[p].assign((size_t) nSeq_ + 1, (int) nSyncParts);
    if (p >= cfg_.n_passes) continue;
    for (int k = 0; k < nSeq_; ++k) {
      syncPartFirst_[p][k] = (int) nSyncParts;
      const int n = pass_[0].shortCircuit[k] ? 0 : (int) pass_[p].rs[k].size();
      for (int r0 = 0; r0 < n; r0 += rblock) {
        if (syncParts_.size() <= nSyncParts) syncParts_.emplace_back();
        SyncPart& sp = syncParts_[nSyncParts++];
        sp.pass = p; sp.seq = k; sp.r0 = r0; sp.r1 = std::min(n, r0 + rblock); sp.err = 0;
        sp.v.clear();
      }
    }
    syncPartFirst_[p][nSeq_] = (int) nSyncParts;
  }
  // One task list.  First the region blocks, largest sequence first; the worker that finishes the last block of one of
  // the few LARGEST sequences goes on with that sequence's join and phasing counts (a sequential state machine,
  // PhasingEvaluator.java:55-142: the longest chains have to start early).  The other sequences' joins follow as tasks
  // of their own at the end of the list, so that the block work is never starved of workers.
  std::vector<int> partOrder;
  partOrder.reserve(nSyncParts);
  std::vector<std::atomic<int>> remaining(nSeq_);
  std::vector<uint8_t> live(nSeq_, 0), inlineClose(nSeq_, 0);
  for (int k = 0; k < nSeq_; ++k) remaining[k].store(0);
  const int nInline = std::max(1, pool_->threads() / 4);
  std::vector<int> tailSeqs;
  for (int t = 0; t < nSeq_; ++t) {
    const int k = order[t];
    for (int p = 0; p < cfg_.n_passes; ++p) {
      for (int q = syncPartFirst_[p][k]; q < syncPartFirst_[p][k + 1]; ++q) partOrder.push_back(q);
      remaining[k].fetch_add(syncPartFirst_[p][k + 1] - syncPartFirst_[p][k]);
    }
    if (t < nInline) inlineClose[k] = 1;
    else tailSeqs.push_back(k);
  }
  auto closeSequence = [&](int k) {
    live[k] = assembleHeader(k, results[k]) ? 1 : 0;
    if (live[k]) phasing(k, seqSync_[0][k], results[k].phasing);
  };
  for (int k = 0; k < nSeq_; ++k) if (inlineClose[k] && remaining[k].load() == 0) closeSequence(k);   // nothing to gather
  const int nParts = (int) partOrder.size();
  pool_->parallelFor(nParts + (int) tailSeqs.size(), [&](int tt, int) {
    if (tt >= nParts) {
      // every block was handed out before this task: the few still running finish within a block's time
      const int k = tailSeqs[(size_t) (tt - nParts)];
      while (remaining[k].load() != 0) std::this_thread::yield();
      closeSequence(k);
      return;
    }
    const int t = partOrder[tt];
    SyncPart& sp = syncParts_[t];
    const PassCtx& pc = pass_[sp.pass];
    const int k = sp.seq;
    const RegRes* rs = pc.rs[k].data();
    std::vector<int32_t>& sync = sp.v;
    sync.reserve((size_t) (sp.r1 - sp.r0) + 8);
    for (int j = sp.r0; j < sp.r1; ++j) {
      const RegRes& s = rs[j];
      if (s.state == RS_MICRO_DONE) {
        const int base = regionStartPos(pc, k, j);
        for (int q = 0; q < s.nSync; ++q) {
          const int pos = base + s.rel[q];
          if (sync.empty() || sync.back() != pos) sync.push_back(pos);
        }
      } else if (s.state == RS_WIDE_DONE) {
        for (int q = 0; q < s.ext.n; ++q) {
          const int pos = pc.wideSync[s.ext.idx + q];
          if (sync.empty() || sync.back() != pos) sync.push_back(pos);
        }
      } else if (s.state == RS_ERR_NULL) {
        sp.err = VCE_ERR_BEST_PATH_NULL;
      } else if (s.state == RS_ERR_MOVE) {
        sp.err = VCE_ERR_MOVE_IN_VARIANT;
      } else if (s.state == RS_ERR_ORDER) {
        sp.err = VCE_ERR_UNSUPPORTED;
      }
    }
    if (remaining[k].fetch_sub(1) == 1 && inlineClose[k]) closeSequence(k);
  });
  tm.lap("assemble: sync points + phasing");
  struct Task { int seq, side, lo, hi; };
  std::vector<Task> tasks;
  const int block = 1 << 15;
  for (int t = 0; t < nSeq_; ++t) {
    const int k = order[t];
    if (!live[k]) continue;
    for (int side = 0; side < 2; ++side) {
      const int n = pass_[0].ss[side][k].n;
      for (int lo = 0; lo < n; lo += block) tasks.push_back(Task{k, side, lo, std::min(n, lo + block)});
    }
  }
  pool_->parallelFor((int) tasks.size(), [&](int t, int) {
    const Task& tk = tasks[t];
    assembleBlock(tk.seq, tk.side, tk.lo, tk.hi, results[tk.seq]);
  });
  tm.lap("assemble: variants");
  summaryFromResults(results);
  int rc = VCE_OK;
  for (int k = 0; k < nSeq_; ++k) {
    if (results[k].error != VCE_OK && rc == VCE_OK) rc = results[k].error;
  }
  return rc;
}

// The split writers' classes counted on the host (the paths that assemble on the host): same rules as the device pass of
// vce_final.h -- SplitEvalSynchronizer.handleKnownCall :112-134, handleKnownBaseline :136-150.
void Engine::summaryFromResults(const vce_sequence_result* results) {
  summary_.assign((size_t) nSeq_ * 8, 0);
  pool_->parallelFor(nSeq_, [&](int k, int) {
    int64_t* cn = summary_.data() + (size_t) k * 8;
    const vce_sequence_result& r = results[k];
    if (r.error != VCE_OK) return;
    const vce_sequence& sq = seqs_[k];
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
  });
}

}  // namespace vce
