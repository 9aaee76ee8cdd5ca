// vce_pipeline.cpp -- see vce_pipeline.h
#include "vce_pipeline.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace vce {

namespace {
const int kDummyPrefix = 0;  // "some baseline variant was included earlier, allele id unknown": the common case
}

// ------------------------------------------------------------------------------------------- packing
int Pipeline::prepare(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs) {
  cfg_ = cfg;
  nSeq_ = nSeq;
  seqs_ = seqs;
  err_.clear();
  if (cfg.n_passes < 1 || cfg.n_passes > 2) { err_ = "n_passes must be 1 or 2"; return VCE_ERR_BAD_ARGUMENT; }
  if (cfg.flag_alternates) { err_ = "flag_alternates is not supported"; return VCE_ERR_UNSUPPORTED; }
  for (int p = 0; p < cfg.n_passes; ++p) {
    if (cfg.ploidy[p] < 1 || cfg.ploidy[p] > VCE_MAX_PLOIDY) { err_ = "ploidy out of range"; return VCE_ERR_BAD_ARGUMENT; }
  }
  // allele tables: side 0 = calls, side 1 = baseline, sequences concatenated
  for (int side = 0; side < 2; ++side) {
    AlleleTable& t = alleles_[side];
    t = AlleleTable();
    slotBase_[side].assign(nSeq + 1, 0);
    size_t nSlots = 0, nNt = 0;
    for (int k = 0; k < nSeq; ++k) {
      const vce_variants& v = side == 0 ? seqs[k].calls : seqs[k].baseline;
      nSlots += v.n_slots;
      nNt += v.n_slots ? v.allele_nt_off[v.n_slots] : 0;
    }
    t.aStart.reserve(nSlots); t.aEnd.reserve(nSlots); t.aNull.reserve(nSlots); t.aNtOff.reserve(nSlots + 1);
    t.nt.reserve(nNt);
    for (int k = 0; k < nSeq; ++k) {
      const vce_variants& v = side == 0 ? seqs[k].calls : seqs[k].baseline;
      slotBase_[side][k] = (int32_t) t.aStart.size();
      const int32_t ntBase = (int32_t) t.nt.size();
      for (int s = 0; s < v.n_slots; ++s) {
        t.aStart.push_back(v.allele_start[s]);
        t.aEnd.push_back(v.allele_end[s]);
        t.aNull.push_back(v.allele_null[s]);
        t.aNtOff.push_back(ntBase + v.allele_nt_off[s]);
      }
      if (v.n_slots) t.nt.insert(t.nt.end(), v.nt, v.nt + v.allele_nt_off[v.n_slots]);
    }
    slotBase_[side][nSeq] = (int32_t) t.aStart.size();
    t.aNtOff.push_back((int32_t) t.nt.size());
    const int rc = be_->setAlleles(side, t);
    if (rc) { err_ = be_->lastError(); return rc; }
  }
  executed_ = false;
  int rc = buildPass0();
  if (rc) return rc;
  return setupPass(pass_[0], true);
}

int Pipeline::buildPass0() {
  PassState& ps = pass_[0];
  ps = PassState();
  ps.pd.pass = 0;
  ps.pd.H = cfg_.ploidy[0];
  ps.pd.kind[0] = cfg_.calls_orientor[0];
  ps.pd.kind[1] = cfg_.baseline_orientor[0];
  ps.shortCircuit.assign(nSeq_, 0);
  for (int k = 0; k < nSeq_; ++k) {
    if (seqs_[k].baseline.n == 0 || seqs_[k].calls.n == 0) ps.shortCircuit[k] = 1;  // SequenceEvaluator.java:94-97
  }
  for (int side = 0; side < 2; ++side) {
    HostSide& hs = ps.pd.side[side];
    hs.seqFirst.assign(nSeq_ + 1, 0);
    std::vector<int32_t> st, en, order;
    for (int k = 0; k < nSeq_; ++k) {
      const vce_variants& v = side == 0 ? seqs_[k].calls : seqs_[k].baseline;
      hs.seqFirst[k] = (int32_t) hs.vStart.size();
      if (v.n > 0) hs.gtPloidy = std::max(hs.gtPloidy, (int) v.gt_ploidy);
      if (ps.shortCircuit[k]) continue;
      const int n = v.n;
      st.resize(n); en.resize(n); order.resize(n);
      bool sorted = true;
      for (int i = 0; i < n; ++i) {  // Allele.getAlleleBounds, Allele.java:195-206
        int s = INT_MAX, e = INT_MIN;
        for (int a = v.allele_off[i]; a < v.allele_off[i + 1]; ++a) {
          if (!v.allele_null[a]) {
            s = std::min(s, v.allele_start[a]);
            e = std::max(e, v.allele_end[a]);
          }
        }
        st[i] = s; en[i] = e; order[i] = i;
        if (i > 0 && (st[i] < st[i - 1] || (st[i] == st[i - 1] && en[i] < en[i - 1]))) sorted = false;
      }
      if (!sorted) {  // PathFinder.java:122-126 stable sort by (start, end)
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
          if (st[a] != st[b]) return st[a] < st[b];
          return en[a] < en[b];
        });
      }
      for (int j = 0; j < n; ++j) {
        const int i = order[j];
        hs.vStart.push_back(st[i]);
        hs.vEnd.push_back(en[i]);
        hs.slot0.push_back(slotBase_[side][k] + v.allele_off[i]);
        hs.nSl.push_back(v.allele_off[i + 1] - v.allele_off[i]);
        uint8_t f = (v.phased && v.phased[i]) ? 1 : 0;
        if (v.status_in) f |= v.status_in[i] & VCE_STATUS_OUTSIDE_EVAL;
        hs.vFlags.push_back(f);
        for (int h = 0; h < 4; ++h) hs.gt.push_back(h < v.gt_ploidy ? v.gt[(size_t) i * v.gt_ploidy + h] : (int8_t) -2);
        hs.inputIdx.push_back(i);
      }
    }
    hs.seqFirst[nSeq_] = (int32_t) hs.vStart.size();
  }
  return VCE_OK;
}

int Pipeline::buildPass1() {
  const PassState& p0 = pass_[0];
  PassState& ps = pass_[1];
  ps = PassState();
  ps.pd.pass = 1;
  ps.pd.H = cfg_.ploidy[1];
  ps.pd.kind[0] = cfg_.calls_orientor[1];
  ps.pd.kind[1] = cfg_.baseline_orientor[1];
  ps.shortCircuit = p0.shortCircuit;
  for (int side = 0; side < 2; ++side) {
    const HostSide& a = p0.pd.side[side];
    HostSide& hs = ps.pd.side[side];
    hs.gtPloidy = a.gtPloidy;
    hs.seqFirst.assign(nSeq_ + 1, 0);
    for (int k = 0; k < nSeq_; ++k) {
      hs.seqFirst[k] = (int32_t) hs.vStart.size();
      for (int v = a.seqFirst[k]; v < a.seqFirst[k + 1]; ++v) {
        if (a.decision[v] != 1) continue;  // Path.getCalledExcluded / getBaselineExcluded of the first pass
        hs.vStart.push_back(a.vStart[v]);
        hs.vEnd.push_back(a.vEnd[v]);
        hs.slot0.push_back(a.slot0[v]);
        hs.nSl.push_back(a.nSl[v]);
        hs.vFlags.push_back(a.vFlags[v]);
        for (int h = 0; h < 4; ++h) hs.gt.push_back(a.gt[(size_t) v * 4 + h]);
        hs.inputIdx.push_back(a.inputIdx[v]);
      }
    }
    hs.seqFirst[nSeq_] = (int32_t) hs.vStart.size();
  }
  return VCE_OK;
}

// ------------------------------------------------------------------------------------------- regions
void Pipeline::splitRegions(PassState& ps) {
  const HostSide& C = ps.pd.side[0];
  const HostSide& B = ps.pd.side[1];
  ps.regions.clear();
  ps.seqRegionFirst.assign(nSeq_ + 1, 0);
  const int gap = std::max(1, opt_.gap);
  for (int k = 0; k < nSeq_; ++k) {
    ps.seqRegionFirst[k] = (int32_t) ps.regions.size();
    if (ps.shortCircuit[k]) continue;
    int i = C.seqFirst[k], j = B.seqFirst[k];
    const int ie = C.seqFirst[k + 1], je = B.seqFirst[k + 1];
    const int tl = seqs_[k].template_len;
    bool first = true;
    do {
      Region rg;
      RegionDesc& d = rg.d;
      std::memset(&d, 0, sizeof(d));
      d.seq = k;
      d.tmplLen = tl;
      d.cFirst = i;
      d.bFirst = j;
      long maxEnd = LONG_MIN;
      int firstStart = INT_MAX;
      while (i < ie || j < je) {
        const bool takeC = j >= je || (i < ie && C.vStart[i] <= B.vStart[j]);
        const int s = takeC ? C.vStart[i] : B.vStart[j];
        const int e = takeC ? C.vEnd[i] : B.vEnd[j];
        if (maxEnd != LONG_MIN && (long) s >= maxEnd + gap) break;
        firstStart = std::min(firstStart, s);
        maxEnd = std::max(maxEnd, (long) e);
        if (takeC) ++i; else ++j;
      }
      d.cCount = i - d.cFirst;
      d.bCount = j - d.bFirst;
      d.cNextStart = i < ie ? C.vStart[i] : kNoNext;
      d.bNextStart = j < je ? B.vStart[j] : kNoNext;
      d.startPos = first ? -1 : firstStart - 1;
      d.prefixAlleleId = first ? kPrefixEmpty : kDummyPrefix;
      rg.assumedPrefix = d.prefixAlleleId;
      first = false;
      ps.regions.push_back(rg);
    } while (i < ie || j < je);
  }
  ps.seqRegionFirst[nSeq_] = (int32_t) ps.regions.size();
}

bool Pipeline::fastEligible(const PassState& ps, const RegionDesc& d) const {
  if (ps.pd.H > 2) return false;
  if (d.cCount > FastLimits::kMaxVarPerSide || d.bCount > FastLimits::kMaxVarPerSide) return false;
  if (1 + std::max(ps.pd.maxOrient[0], ps.pd.maxOrient[1]) > FastLimits::kMaxChildren) return false;
  return true;
}

// Upper bound on the number of orientations any variant of a pass can have (Orientor.java), used to size child
// slots and decision fields before K1 has run.
int Pipeline::orientBound(int kind, int H, int gtPloidy, int maxSlots) {
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

void Pipeline::buildLaunch(PassState& ps, const std::vector<int>& ids, bool general, Launch& L) {
  L.ids = ids;
  L.descs.clear();
  L.windows.clear();
  L.syncOff.clear();
  L.descs.reserve(ids.size());
  L.syncOff.reserve(ids.size());
  const int mo = std::max(ps.pd.maxOrient[0], ps.pd.maxOrient[1]);
  const HostSide& C = ps.pd.side[0];
  const HostSide& B = ps.pd.side[1];
  size_t winBytes = 0;
  for (int id : ids) {
    Region& rg = ps.regions[id];
    RegionDesc& d = rg.d;
    if (general) {
      d.qMax = std::max(1, std::max(d.cCount, d.bCount));
      d.decBits = (mo + 2 <= 15) ? 4 : 8;
      PathLayout PL;
      PL.init(ps.pd.H, d.qMax, std::max(1, std::max(d.cCount, d.bCount)), d.cCount + d.bCount + 2, d.decBits);
      d.layoutW = PL.W;
    }
    // Reference window of the region: from just before its first variant to `margin` bases beyond the furthest
    // variant end (the whole remainder of the template when extraMargin < 0).
    int lo = INT_MAX;
    long hi = 0;
    for (int v = d.cFirst; v < d.cFirst + d.cCount; ++v) { lo = std::min(lo, C.vStart[v]); hi = std::max(hi, (long) C.vEnd[v]); }
    for (int v = d.bFirst; v < d.bFirst + d.bCount; ++v) { lo = std::min(lo, B.vStart[v]); hi = std::max(hi, (long) B.vEnd[v]); }
    if (lo == INT_MAX) lo = 0;
    const int extra = rg.extraMargin;
    const int ws = std::max(0, std::min(lo, d.startPos < 0 ? 0 : d.startPos) - 1);
    long we = extra < 0 ? (long) d.tmplLen : hi + opt_.margin + extra;
    we = std::min(we, (long) d.tmplLen);
    if (we < ws) we = ws;
    d.winStart = ws;
    d.winLen = (int32_t) (we - ws);
    d.winOff = (uint32_t) winBytes;
    winBytes += ((size_t) d.winLen + 15 + 16) & ~(size_t) 15;  // 16-byte aligned and padded: vector / bulk loads
    L.descs.push_back(d);
    L.syncOff.push_back(rg.syncOff);
  }
  L.windows.assign(winBytes + 16, 0);
  for (const RegionDesc& d : L.descs) {
    if (d.winLen > 0) std::memcpy(L.windows.data() + d.winOff, seqs_[d.seq].template_bases + d.winStart, (size_t) d.winLen);
  }
}

int Pipeline::runLaunch(PassState& ps, Launch& L, bool general, bool staged) {
  if (L.ids.empty()) return VCE_OK;
  // a re-run region reports its warnings afresh
  if (!ps.warns.empty()) {
    std::vector<uint8_t> redo(ps.regions.size(), 0);
    for (int id : L.ids) redo[id] = 1;
    std::vector<WarnRec> keep;
    for (const WarnRec& wr : ps.warns) if (!redo[wr.region]) keep.push_back(wr);
    ps.warns.swap(keep);
  }
  std::vector<RegionResult> out(L.ids.size());
  const SearchConfig sc{cfg_.max_paths, cfg_.max_iterations, cfg_.prune_no_ops, cfg_.preference};
  const size_t warn0 = ps.warns.size();
  const int rc = be_->search(ps.pd, sc, L.descs, L.windows, general, staged, out, L.syncOff, ps.syncBuf.size(), ps.warns);
  if (rc) { err_ = be_->lastError(); return rc; }
  for (size_t q = 0; q < L.ids.size(); ++q) {
    Region& rg = ps.regions[L.ids[q]];
    rg.r = out[q];
    rg.general = general;
  }
  // warnings carry the launch-local region index: translate to the pass's region id
  for (size_t wi = warn0; wi < ps.warns.size(); ++wi) ps.warns[wi].region = L.ids[ps.warns[wi].region];
  return VCE_OK;
}

int Pipeline::launch(PassState& ps, const std::vector<int>& ids, bool general) {
  if (ids.empty()) return VCE_OK;
  Launch L;
  buildLaunch(ps, ids, general, L);
  return runLaunch(ps, L, general, false);
}

// Upload the pass, split it into regions and (optionally) stage the main shared-memory launch on the device.
int Pipeline::setupPass(PassState& ps, bool stageMain) {
  for (int side = 0; side < 2; ++side) {
    const HostSide& hs = ps.pd.side[side];
    int maxSl = 2;
    for (int x : hs.nSl) maxSl = std::max(maxSl, x);
    ps.pd.maxOrient[side] = orientBound(ps.pd.kind[side], ps.pd.H, hs.gtPloidy, maxSl);
    if (ps.pd.maxOrient[side] + 2 > 255) { err_ = "too many orientations per variant for this engine"; return VCE_ERR_UNSUPPORTED; }
  }
  int rc = be_->uploadPass(ps.pd);
  if (rc) { err_ = be_->lastError(); return rc; }
  splitRegions(ps);
  size_t so = 0;  // sync-point output space: cCount + bCount + 2 per region
  for (Region& rg : ps.regions) {
    rg.syncOff = (int) so;
    so += (size_t) rg.d.cCount + rg.d.bCount + 2;
    rg.r = RegionResult();
  }
  ps.syncBuf.assign(so, 0);
  ps.initSyncSize = so;
  std::vector<int> fast;
  ps.mainGen.clear();
  for (int id = 0; id < (int) ps.regions.size(); ++id) (fastEligible(ps, ps.regions[id].d) ? fast : ps.mainGen).push_back(id);
  buildLaunch(ps, fast, false, ps.mainFast);
  ps.initRegions = ps.regions;
  if (stageMain && !ps.mainFast.ids.empty()) {
    rc = be_->stage(ps.mainFast.descs, ps.mainFast.windows, ps.mainFast.syncOff, ps.syncBuf.size());
    if (rc) { err_ = be_->lastError(); return rc; }
  }
  return VCE_OK;
}

int Pipeline::runPass(PassState& ps, bool staged) {
  int rc;
  ps.regions = ps.initRegions;
  ps.syncBuf.assign(ps.initSyncSize, 0);
  ps.warns.clear();
  ps.fetched = false;
  if ((rc = be_->orient(ps.pd))) { err_ = be_->lastError(); return rc; }
  be_->stats.regions += (int64_t) ps.regions.size();
  be_->stats.escalated += (int64_t) ps.mainGen.size();
  if ((rc = runLaunch(ps, ps.mainFast, false, staged))) return rc;
  if ((rc = launch(ps, ps.mainGen, true))) return rc;

  // ---- settle SPILL / OVERFLOW, then prefix-dependent tie-breaks
  std::vector<int> pending;  // regions whose latest run must be inspected
  pending.reserve(ps.regions.size());
  for (int id = 0; id < (int) ps.regions.size(); ++id) pending.push_back(id);
  for (int round = 0; round < 100000; ++round) {
    std::vector<int> redoFast, redoGen;
    std::sort(pending.begin(), pending.end());
    for (int id : pending) {
      Region& rg = ps.regions[id];
      if (!rg.alive) continue;
      const int st = rg.r.status;
      if (st == kRegionOverflow) {
        if (rg.general) { err_ = "search capacity exceeded in the large-capacity kernel"; return VCE_ERR_CAPACITY; }
        redoGen.push_back(id);
        ++be_->stats.escalated;
      } else if (st == kRegionSpill) {
        ++be_->stats.spilled;
        // absorb following regions of the same sequence
        const int seqEnd = ps.seqRegionFirst[rg.d.seq + 1];
        int nx = id + 1;
        while (nx < seqEnd && !ps.regions[nx].alive) ++nx;
        if (nx >= seqEnd) {
          // last region of its sequence ran off its reference window: widen it
          if (rg.extraMargin < 0) { err_ = "region spilled with the whole template in its window"; return VCE_ERR_CAPACITY; }
          rg.extraMargin = rg.extraMargin == 0 ? 256 : (rg.extraMargin >= (1 << 20) ? -1 : rg.extraMargin * 16);
        } else {
          bool more = true;
          while (more && nx < seqEnd) {
            Region& o = ps.regions[nx];
            more = o.r.status == kRegionSpill;
            rg.d.cCount += o.d.cCount;
            rg.d.bCount += o.d.bCount;
            rg.d.cNextStart = o.d.cNextStart;
            rg.d.bNextStart = o.d.bNextStart;
            o.alive = false;
            ++nx;
            while (nx < seqEnd && !ps.regions[nx].alive) ++nx;
          }
          rg.extraMargin = 0;
          rg.syncOff = (int) ps.syncBuf.size();
          ps.syncBuf.resize(ps.syncBuf.size() + (size_t) rg.d.cCount + rg.d.bCount + 2, 0);
        }
        (fastEligible(ps, rg.d) ? redoFast : redoGen).push_back(id);
      } else if (st == kRegionErrNull || st == kRegionErrMove || st == kRegionErrOrder) {
        // reported per sequence in assemble()
      }
    }
    pending.clear();
    if (redoFast.empty() && redoGen.empty()) {
      // prefix-dependent tie-breaks (MaxSumBoth.java:55,76): re-run the regions whose assumption was wrong
      std::vector<int> fixF, fixG;
      for (int k = 0; k < nSeq_; ++k) {
        int prefix = kPrefixEmpty;
        for (int id = ps.seqRegionFirst[k]; id < ps.seqRegionFirst[k + 1]; ++id) {
          Region& rg = ps.regions[id];
          if (!rg.alive) continue;
          if (rg.r.usedPrefix && rg.assumedPrefix != prefix) {
            rg.d.prefixAlleleId = prefix;
            rg.assumedPrefix = prefix;
            (rg.general ? fixG : fixF).push_back(id);
            break;  // later regions of this sequence depend on this one's outcome
          }
          if (rg.r.lastBaseAlleleId != kPrefixEmpty) prefix = rg.r.lastBaseAlleleId;
        }
      }
      if (fixF.empty() && fixG.empty()) break;
      be_->stats.prefixReruns += (int64_t) (fixF.size() + fixG.size());
      if ((rc = launch(ps, fixF, false))) return rc;
      if ((rc = launch(ps, fixG, true))) return rc;
      pending.insert(pending.end(), fixF.begin(), fixF.end());
      pending.insert(pending.end(), fixG.begin(), fixG.end());
      continue;
    }
    if ((rc = launch(ps, redoFast, false))) return rc;
    if ((rc = launch(ps, redoGen, true))) return rc;
    pending.insert(pending.end(), redoFast.begin(), redoFast.end());
    pending.insert(pending.end(), redoGen.begin(), redoGen.end());
  }
  for (const Region& rg : ps.regions) if (rg.alive) be_->stats.iterations += rg.r.iterations;
  if (const char* dump = std::getenv("VCE_DUMP_REGIONS")) {  // diagnostics: one line per settled region
    if (FILE* f = std::fopen(dump, ps.pd.pass == 0 ? "w" : "a")) {
      for (const Region& rg : ps.regions) {
        if (!rg.alive) continue;
        std::fprintf(f, "%d,%d,%d,%d,%d,%d,%d,%d\n", ps.pd.pass, rg.d.cCount, rg.d.bCount, rg.r.maxFrontier, rg.r.iterations,
                     rg.r.nSync, rg.r.status, rg.general ? 1 : 0);
      }
      std::fclose(f);
    }
  }
  return VCE_OK;
}

int Pipeline::execute() {
  int rc;
  if (executed_) {  // repeated execution of a staged job: only the device-side outputs of pass 0 are cleared
    if ((rc = be_->resetPass(pass_[0].pd))) { err_ = be_->lastError(); return rc; }
  }
  executed_ = true;
  if ((rc = runPass(pass_[0], true))) return rc;
  if (cfg_.n_passes == 2) {
    PassState& p0 = pass_[0];
    if ((rc = be_->fetchPass(p0.pd, p0.syncBuf))) { err_ = be_->lastError(); return rc; }
    p0.fetched = true;
    // the second pass needs the device for itself: its arrays replace the first pass's
    if ((rc = buildPass1())) return rc;
    if ((rc = setupPass(pass_[1], false))) return rc;
    if ((rc = runPass(pass_[1], false))) return rc;
    if ((rc = be_->fetchPass(pass_[1].pd, pass_[1].syncBuf))) { err_ = be_->lastError(); return rc; }
    pass_[1].fetched = true;
  }
  return VCE_OK;
}

// ------------------------------------------------------------------------------------------- assembly
// PhasingEvaluator.countMisphasings, src/com/rtg/vcf/eval/PhasingEvaluator.java:55-142, on the flat pass-0 results.
namespace {
struct VSum { int start; bool phased, included, phase; };
struct CallIter {  // PhasingEvaluator.CallIterator :160-195
  const std::vector<VSum>& inc;
  const std::vector<VSum>& exc;
  size_t i = 0, e = 0;
  CallIter(const std::vector<VSum>& a, const std::vector<VSum>& b) : inc(a), exc(b) {}
  bool hasNext() const { return i < inc.size() || e < exc.size(); }
  VSum next() {
    const bool haveI = i < inc.size(), haveE = e < exc.size();
    if (!haveI || (haveE && exc[e].start < inc[i].start)) return exc[e++];
    return inc[i++];
  }
};
bool groupInPhase(const std::vector<VSum>& g) {  // :143-155
  if (g.empty()) return true;
  if (!g[0].phased) return false;
  const bool ph = g[0].phase;
  for (const VSum& v : g) if (!v.phased || v.phase != ph) return false;
  return true;
}
}  // namespace

void Pipeline::phasing(int k, int32_t out[3]) const {
  const PassState& ps = pass_[0];
  std::vector<VSum> inc[2], exc[2];
  for (int side = 0; side < 2; ++side) {
    const HostSide& hs = ps.pd.side[side];
    for (int v = hs.seqFirst[k]; v < hs.seqFirst[k + 1]; ++v) {
      const int dec = hs.decision[v];
      const bool ph = hs.vFlags[v] & 1;
      if (dec == 1) exc[side].push_back(VSum{hs.vStart[v], ph, false, false});
      else if (dec >= 2) inc[side].push_back(VSum{hs.vStart[v], ph, true, hs.oOrig[hs.oOff[v] + dec - 2] != 0});
    }
  }
  // sync points of pass 0 for this sequence
  std::vector<int32_t> sync;
  for (int id = ps.seqRegionFirst[k]; id < ps.seqRegionFirst[k + 1]; ++id) {
    const Region& rg = ps.regions[id];
    if (!rg.alive) continue;
    for (int q = 0; q < rg.r.nSync; ++q) {
      const int sp = ps.syncBuf[rg.syncOff + q];
      if (sync.empty() || sync.back() != sp) sync.push_back(sp);
    }
  }
  CallIter baseline(inc[1], exc[1]);
  CallIter calls(inc[0], exc[0]);
  int mis = 0, unph = 0, correct = 0;
  bool baseIsPhased = false, basePhase = false, callIsPhased = false, callPhase = false;
  bool haveCall = false, haveBase = false;
  VSum call{0, false, false, false}, base{0, false, false, false};
  for (int point : sync) {
    std::vector<VSum> bsec, csec;
    do {
      if (haveBase && base.start < point) {
        bsec.push_back(base);
        haveBase = baseline.hasNext();
        if (haveBase) base = baseline.next();
      }
      if (!haveBase) {
        haveBase = baseline.hasNext();
        if (haveBase) base = baseline.next();
      }
    } while (haveBase && base.start < point);
    do {
      if (haveCall) csec.push_back(call);
      haveCall = calls.hasNext();
      if (haveCall) call = calls.next();
    } while (haveCall && call.start < point);
    if (!groupInPhase(bsec)) {
      for (const VSum& s : csec) if (s.phased) ++unph;
      baseIsPhased = false;
      callIsPhased = false;
      continue;
    }
    bool transition = false;
    if (!baseIsPhased) callIsPhased = false;
    for (const VSum& b : bsec) {
      if (b.phased) {
        if (basePhase != b.phase) transition = true;
        baseIsPhased = true;
      }
      basePhase = b.phase;
    }
    for (const VSum& c : csec) {
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
  out[0] = mis; out[1] = correct; out[2] = unph;
}

void Pipeline::assemble(int k, vce_sequence_result& res) {
  const vce_sequence& sq = seqs_[k];
  res.n_sync[0] = res.n_sync[1] = 0;
  res.phasing[0] = res.phasing[1] = res.phasing[2] = 0;
  res.n_warnings = 0;
  res.error = VCE_OK;
  res.n_regions = 0;
  res.n_iterations = 0;
  vce_side_result* sides[2] = {&res.calls, &res.baseline};
  const vce_variants* vin[2] = {&sq.calls, &sq.baseline};
  for (int side = 0; side < 2; ++side) {
    vce_side_result& o = *sides[side];
    const int n = vin[side]->n;
    for (int i = 0; i < n; ++i) {
      o.status[i] = vin[side]->status_in ? vin[side]->status_in[i] : 0;
      for (int h = 0; h < VCE_MAX_PLOIDY; ++h) o.allele_ids[(size_t) i * VCE_MAX_PLOIDY + h] = -2;
      o.original[i] = 0;
      o.weight[i] = 0;
      if (o.sync_id) o.sync_id[i] = 0;
    }
  }
  if (pass_[0].shortCircuit[k]) {  // SequenceEvaluator.java:94-97
    for (int side = 0; side < 2; ++side) {
      for (int i = 0; i < vin[side]->n; ++i) {
        sides[side]->status[i] |= VCE_STATUS_NO_MATCH;
        if (sides[side]->sync_id) sides[side]->sync_id[i] = 1;
      }
    }
    return;
  }
  // sync points per pass + per-sequence errors / diagnostics
  std::vector<int32_t> sync[2];
  for (int p = 0; p < cfg_.n_passes; ++p) {
    const PassState& ps = pass_[p];
    for (int id = ps.seqRegionFirst[k]; id < ps.seqRegionFirst[k + 1]; ++id) {
      const Region& rg = ps.regions[id];
      if (!rg.alive) continue;
      if (rg.r.status == kRegionErrNull) res.error = VCE_ERR_BEST_PATH_NULL;
      else if (rg.r.status == kRegionErrMove) res.error = VCE_ERR_MOVE_IN_VARIANT;
      else if (rg.r.status == kRegionErrOrder) res.error = VCE_ERR_UNSUPPORTED;
      res.n_iterations += rg.r.iterations;
      for (int q = 0; q < rg.r.nSync; ++q) {
        const int sp = ps.syncBuf[rg.syncOff + q];
        if (sync[p].empty() || sync[p].back() != sp) sync[p].push_back(sp);  // Path.java:293 de-duplication
      }
    }
    // re-runs append out of order: report in region order, as the sequential reference would
    std::vector<WarnRec> ws(ps.warns);
    std::stable_sort(ws.begin(), ws.end(), [](const WarnRec& a, const WarnRec& b) { return a.region < b.region; });
    for (const WarnRec& wr : ws) {
      if (ps.regions[wr.region].d.seq != k || !ps.regions[wr.region].alive) continue;
      if (res.n_warnings < res.warn_cap) res.warnings[res.n_warnings] = vce_warning{p, wr.paths, wr.iterations, wr.start1, wr.end1};
      ++res.n_warnings;
    }
  }
  if (res.error != VCE_OK) return;
  for (int p = 0; p < cfg_.n_passes; ++p) {
    res.n_sync[p] = (int32_t) sync[p].size();
    if ((int) sync[p].size() > res.sync_cap[p]) { res.error = VCE_ERR_CAPACITY; return; }
    std::copy(sync[p].begin(), sync[p].end(), res.sync_points[p]);
  }
  res.n_regions = (int64_t) sync[0].size();
  // statuses (SequenceEvaluator.mergeVariants :170-182, insertSkipped :190-209)
  for (int side = 0; side < 2; ++side) {
    vce_side_result& o = *sides[side];
    const HostSide& a = pass_[0].pd.side[side];
    std::vector<uint8_t> seen(vin[side]->n, 0);
    for (int v = a.seqFirst[k]; v < a.seqFirst[k + 1]; ++v) {
      const int i = a.inputIdx[v];
      const int dec = a.decision[v];
      if (dec >= 2) {
        const int row = a.oOff[v] + dec - 2;
        o.status[i] |= VCE_STATUS_GT_MATCH;
        for (int h = 0; h < 4; ++h) o.allele_ids[(size_t) i * 4 + h] = a.oIds[(size_t) row * 4 + h];
        o.original[i] = a.oOrig[row];
        o.weight[i] = side == 0 ? a.weight[v] : 0.0;
        seen[i] = 1;
      } else if (dec == 1 && cfg_.n_passes == 1) {
        o.status[i] |= VCE_STATUS_NO_MATCH;
        seen[i] = 1;
      }
    }
    if (cfg_.n_passes == 2) {
      const HostSide& b = pass_[1].pd.side[side];
      for (int v = b.seqFirst[k]; v < b.seqFirst[k + 1]; ++v) {
        const int i = b.inputIdx[v];
        const int dec = b.decision[v];
        if (dec >= 2) {
          const int row = b.oOff[v] + dec - 2;
          o.status[i] |= VCE_STATUS_ALLELE_MATCH;
          for (int h = 0; h < 4; ++h) o.allele_ids[(size_t) i * 4 + h] = b.oIds[(size_t) row * 4 + h];
          o.original[i] = b.oOrig[row];
          o.weight[i] = side == 0 ? b.weight[v] : 0.0;
          seen[i] = 1;
        } else if (dec == 1) {
          o.status[i] |= VCE_STATUS_NO_MATCH;
          seen[i] = 1;
        }
      }
    }
    for (int i = 0; i < vin[side]->n; ++i) if (!seen[i]) o.status[i] |= VCE_STATUS_SKIPPED;
    if (o.sync_id) {
      // SYNC annotation value (InterleavingEvalSynchronizer.floorSyncPos :71-83 + WithInfoEvalSynchronizer :121-222)
      auto floorVal = [](const std::vector<int32_t>& sp, int start) -> int {
        if (sp.empty()) return 0;
        const int vv = start - 1;
        auto it = std::upper_bound(sp.begin(), sp.end(), vv);
        const size_t idx = it == sp.begin() ? 0 : (size_t) (it - sp.begin()) - 1;
        return sp[idx] + 1;
      };
      for (int v = a.seqFirst[k]; v < a.seqFirst[k + 1]; ++v) {
        const int i = a.inputIdx[v];
        const int st = o.status[i];
        const int s1 = floorVal(sync[0], a.vStart[v]);
        const int s2 = floorVal(sync[1], a.vStart[v]);
        int val = 0;
        if (st & VCE_STATUS_SKIPPED) val = 0;
        else if (st & VCE_STATUS_GT_MATCH) val = s1 + 1;
        else if (st & VCE_STATUS_ALLELE_MATCH) val = s2 + 1;
        else if (st & VCE_STATUS_NO_MATCH) val = s2 > 0 ? s2 + 1 : s1 + 1;
        o.sync_id[i] = val;
      }
    }
  }
  phasing(k, res.phasing);
}

int Pipeline::finish(vce_sequence_result* results) {
  int rc = VCE_OK;
  if (!pass_[0].fetched) {
    if ((rc = be_->fetchPass(pass_[0].pd, pass_[0].syncBuf))) { err_ = be_->lastError(); return rc; }
    pass_[0].fetched = true;
  }
  for (int k = 0; k < nSeq_; ++k) {
    assemble(k, results[k]);
    if (results[k].error != VCE_OK && rc == VCE_OK) rc = results[k].error;
  }
  return rc;
}

int Pipeline::run(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs, vce_sequence_result* results) {
  int rc = prepare(cfg, nSeq, seqs);
  if (rc) return rc;
  if ((rc = execute())) return rc;
  return finish(results);
}

}  // namespace vce
