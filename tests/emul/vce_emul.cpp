// vce_emul.cpp -- TEST-ONLY host instantiation of the engine's host pipeline.
//
// There is no GPU in the build container.  To exercise the HOST logic of the product (packing, region
// splitting, SPILL merging, overflow escalation, prefix re-runs, result assembly) in `-m "not gpu"` tests,
// this file plugs a 1-lane CPU instantiation of the very same search source (vce_search.h, HostLane policy)
// behind the pipeline's Backend interface.  It is built into tests/emul/libvce_emul.so, exported under the
// `emul_` prefix, and is never linked into or loaded by the product library (rtg_tools_b200/csrc/libvce.so),
// which fails loudly without a CUDA device.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include <atomic>
#include <cstdlib>
#include <mutex>

#include "../../rtg_tools_b200/csrc/vce_engine.h"
#include "../../rtg_tools_b200/csrc/vce_orient.h"
#include "../../rtg_tools_b200/csrc/vce_pipeline.h"
#include "../../rtg_tools_b200/csrc/vce_bgzf.h"
#include "../../rtg_tools_b200/csrc/vce_split.h"
#include "../../rtg_tools_b200/csrc/vce_vcf.h"

namespace {
using namespace vce;

class EmulBackend : public Backend {
 public:
  bool forceGeneral = false;
  std::atomic<int64_t> microRegions{0}, microClean{0};
  // ---- thread-per-region path: the same MicroSearch source, one region after the other
  void* hostAlloc(size_t bytes) override { return std::malloc(bytes); }
  void hostFree(void* p) override { std::free(p); }
  int microReset() override { return 0; }
  int microUpload(MicroJob&) override { return 0; }
  template <class T>
  void microRunT(MicroJob& j, const MicroConfig& mc) {
    std::vector<uint8_t> ws(T::WS_BYTES + 16);
    for (int q = 0; q < j.n; ++q) {
      MicroSearch<T> ms;
      ms.ws = ws.data();
      ms.cfg = mc;
      const uint8_t* src = j.packets + (size_t) j.offsets[q] * 4;
      std::memcpy(ws.data() + T::OFF_PKT, src, (size_t) src[MP_NWORDS] * 4);
      ms.runRegion(j.results + (size_t) q * j.resBytes);
      if (j.results[(size_t) q * j.resBytes + MR_STATUS] == kMicroClean) ++microClean;
    }
  }
  int microRun(MicroJob& j, const MicroConfig& mc) override {
    if (j.cls == 0) microRunT<MicroA>(j, mc); else microRunT<MicroB>(j, mc);
    microRegions += j.n;
    ++stats.launches;
    return 0;
  }
  int microRunAll(std::vector<MicroJob*>& jobs, const MicroConfig& mc) override {
    for (MicroJob* j : jobs) microRun(*j, mc);
    return 0;
  }
  int microWait(MicroJob&) override { return 0; }
  // device-side packet building, emulated: the same builder source over the host slab, then the search as above
  template <class T>
  void buildRunT(const MicroBuildJob& b, const PacketView& pv, int q0, MicroJob& j, const MicroConfig& mc, std::vector<uint8_t>& pkStore,
                 std::vector<uint32_t>& offStore) {
    const RegRange* regs = reinterpret_cast<const RegRange*>(b.slab + b.regions);
    const uint32_t* winOff = reinterpret_cast<const uint32_t*>(b.slab + b.winOff);
    const size_t stride = ((size_t) T::PKT_MAX + 15) & ~(size_t) 15;
    pkStore.assign((size_t) j.n * stride + 16, 0);
    offStore.assign((size_t) j.n + 1, 0xffffffffu);
    std::vector<uint8_t> ws(T::WS_BYTES + 16);
    for (int q = 0; q < j.n; ++q) {
      uint8_t* res = j.results + (size_t) q * j.resBytes;
      uint8_t* pkq = pkStore.data() + (size_t) q * stride;
      struct { uint8_t* p; uint8_t* data() { return p; } uint8_t operator[](int i) const { return p[i]; } } pk{pkq};
      int32_t winStart = 0, winLen = 0, nextStart[2];
      int bytes = 0;
      if (microWindow(pv, regs[q0 + q], &winStart, &winLen, nextStart)) {
        if (b.resident) bytes = microBuildPacketW<T>(pv, regs[q0 + q], winStart, winLen, nextStart, WinPacked{b.t2, b.nmask, winStart}, pk.data());
        else bytes = microBuildPacket<T>(pv, regs[q0 + q], winStart, winLen, nextStart, b.slab + b.windows + winOff[q0 + q], pk.data());
      }
      if (bytes <= 0) {
        std::memset(res, 0, (size_t) j.resBytes);
        res[MR_STATUS] = kMicroFallback;
        continue;
      }
      offStore[(size_t) q] = (uint32_t) (((size_t) q * stride) >> 2);
      MicroSearch<T> ms;
      ms.ws = ws.data();
      ms.cfg = mc;
      std::memcpy(ws.data() + T::OFF_PKT, pk.data(), (size_t) pk[MP_NWORDS] * 4);
      ms.runRegion(res);
      if (res[MR_STATUS] == kMicroClean) ++microClean;
    }
  }
  // ---- device-resident results, emulated: the same scatter / patch / final-pass source (vce_final.h) over host vectors
  struct HostFinalOps {
    template <class F> void parFor(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
    void exclusiveSum(int32_t* in, int32_t* out, int n) { int32_t acc = 0; for (int i = 0; i < n; ++i) { const int32_t v = in[i]; out[i] = acc; acc += v; } }
    int unique64(int64_t* in, int64_t* out, int n) {
      int m = 0;
      for (int i = 0; i < n; ++i) if (m == 0 || out[m - 1] != in[i]) out[m++] = in[i];
      return m;
    }
    void segMinScan(const int32_t* keys, const int32_t* vals, int32_t* out, int n) {
      for (int i = 0; i < n; ++i) out[i] = (i > 0 && keys[i] == keys[i - 1]) ? std::min(out[i - 1], vals[i]) : vals[i];
    }
    int fetchInt(const int32_t* p) { return *p; }
    void zero(void* p, size_t bytes) { std::memset(p, 0, bytes); }
  };
  bool resultsSupported() const override { return resultsOn; }
  bool resultsOn = true;
  ResultsSetup rsu_;
  std::vector<FinalSeq> fseq_;
  std::vector<uint8_t> rDec_[2], rRel_, fOutC_, fOutB_, fMFlags_, fBaseSum_, fBlockEnd_, fSuperEnd_;
  std::vector<int32_t> fSuperCnt_, fSeqSuper_;
  std::vector<double> rWeight_, fRoc_;
  std::vector<uint32_t> rInfo_;
  std::vector<int32_t> rCnt_, rBase_, fRegOff_, fPos_, fSeqSync_, fKept_, fMStart_, fMSeq_, fMVal_, fMMin_, fCall0_, fBlockCnt_, fSeqBlock_;
  std::vector<int64_t> fKeys_, fUniq_, fCounters_;
  unsigned long long rIter_ = 0;
  std::atomic<int64_t> scattered{0};
  int resultsBegin(const ResultsSetup& s) override {
    rsu_ = s;
    const size_t nC = (size_t) s.nVar[0], nB = (size_t) s.nVar[1], nR = (size_t) s.nReg, nS = s.seq.size();
    rDec_[0].assign(nC + 16, 0xee); rDec_[1].assign(nB + 16, 0xee); rWeight_.assign(nC + 2, -7.0);   // poison: every entry must be written
    rInfo_.assign(nR + 2, 0); rCnt_.assign(nR + 2, 0); rBase_.assign(nR + 2, 0); rRel_.assign((nR + 2) * kFinalRel, 0);
    rIter_ = 0;
    fseq_.assign(nS, FinalSeq());
    for (size_t k = 0; k < nS; ++k) {
      for (int side = 0; side < 2; ++side) {
        fseq_[k].vStart[side] = s.seq[k].vStart[side]; fseq_[k].phased[side] = s.seq[k].phased[side]; fseq_[k].statusIn[side] = s.seq[k].statusIn[side];
        fseq_[k].first[side] = s.seq[k].first[side]; fseq_[k].n[side] = s.seq[k].n[side];
      }
      fseq_[k].regFirst = s.seq[k].regFirst; fseq_[k].nReg = s.seq[k].nReg;
    }
    return 0;
  }
  std::mutex scatterMu;
  void scatterJob(const MicroJob& j, const RegRange* regs, const int32_t* regIdx, const uint8_t* packets, const uint32_t* offsets) {
    std::lock_guard<std::mutex> lk(scatterMu);   // (host emulation: the iteration counter is a plain sum)
    ScatterArgs a;
    a.results = j.results; a.regs = regs; a.regIdx = regIdx; a.packets = packets; a.offsets = offsets;
    a.n = j.n; a.resBytes = j.resBytes;
    a.kv = j.cls == 0 ? MicroA::KV : MicroB::KV; a.smax = j.cls == 0 ? MicroA::SMAX : MicroB::SMAX;
    a.dec[0] = rDec_[0].data(); a.dec[1] = rDec_[1].data(); a.weight = rWeight_.data();
    a.regInfo = rInfo_.data(); a.regCnt = rCnt_.data(); a.regBase = rBase_.data(); a.regRel = rRel_.data();
    a.iterations = &rIter_;
    for (int q = 0; q < j.n; ++q) scatterOne(a, q);
    scattered += j.n;
  }
  int resultsRegInfo(const uint32_t** info, int64_t* iterations) override {
    *info = rInfo_.data();
    *iterations = (int64_t) rIter_;
    return 0;
  }
  int resultsFinish(const ResultsPatch& p, ResultsOut& out) override {
    const size_t nC = (size_t) rsu_.nVar[0], nB = (size_t) rsu_.nVar[1], nR = (size_t) rsu_.nReg, nS = rsu_.seq.size();
    PatchArgs pa;
    pa.n = (int) p.regIdx.size();
    pa.regIdx = p.regIdx.data(); pa.info = p.info.data(); pa.syncOff = p.syncOff.data(); pa.syncCnt = p.syncCnt.data(); pa.range = p.range.data();
    pa.payOff = p.payOff.data(); pa.decPay = p.decPay.data(); pa.wOff = p.wOff.data(); pa.wPay = p.wPay.data();
    pa.dec[0] = rDec_[0].data(); pa.dec[1] = rDec_[1].data(); pa.weight = rWeight_.data();
    pa.regInfo = rInfo_.data(); pa.regCnt = rCnt_.data(); pa.regBase = rBase_.data();
    for (int q = 0; q < pa.n; ++q) patchOne(pa, q);
    const size_t nFlatMax = nC + nB + 2 * nR + 4;
    fRegOff_.assign(nR + 2, 0); fKeys_.assign(nFlatMax, 0); fUniq_.assign(nFlatMax, 0); fPos_.assign(nFlatMax, 0); fSeqSync_.assign(nS + 2, 0);
    fOutC_.assign(nC * 16 + 16, 0); fOutB_.assign(nB * 8 + 16, 0); fCounters_.assign(nS * 8 + 8, 0); fKept_.assign(nC + 2, 0);
    fMStart_.assign(nC + 2, 0); fMFlags_.assign(nC + 2, 0); fMSeq_.assign(nC + 2, 0); fMVal_.assign(nC + 2, 0); fMMin_.assign(nC + 2, 0);
    fBaseSum_.assign(nFlatMax, 0); fCall0_.assign(nFlatMax + nS + 2, 0);
    const size_t nBlocksMax = nFlatMax / kPhaseBlock + nS + 2;
    fBlockEnd_.assign(nBlocksMax * 16, 0); fBlockCnt_.assign(nBlocksMax * 16 * 3, 0); fSeqBlock_.assign(nS + 2, 0);
    const size_t nSuperMax = nBlocksMax / kPhaseGroup + nS + 2;
    fSuperEnd_.assign(nSuperMax * 16, 0); fSuperCnt_.assign(nSuperMax * 16 * 3, 0); fSeqSuper_.assign(nS + 2, 0);
    fRoc_.assign(nC * 3 + 3, 0.0);
    FinalArrays A;
    A.nSeq = (int) nS; A.H = rsu_.H; A.kind[0] = rsu_.kind[0]; A.kind[1] = rsu_.kind[1];
    A.seq = fseq_.data();
    A.nVar[0] = rsu_.nVar[0]; A.nVar[1] = rsu_.nVar[1]; A.nReg = rsu_.nReg;
    A.dec[0] = rDec_[0].data(); A.dec[1] = rDec_[1].data(); A.weight = rWeight_.data();
    A.regInfo = rInfo_.data(); A.regCnt = rCnt_.data(); A.regBase = rBase_.data(); A.regRel = rRel_.data(); A.extSync = p.extSync.data();
    A.regOff = fRegOff_.data(); A.syncKeys = fKeys_.data(); A.syncUniq = fUniq_.data(); A.syncPos = fPos_.data(); A.seqSyncFirst = fSeqSync_.data();
    A.outCalls = fOutC_.data(); A.outBase = fOutB_.data(); A.counters = fCounters_.data(); A.rocTuples = fRoc_.data();
    A.keptScan = fKept_.data(); A.mStart = fMStart_.data(); A.mFlags = fMFlags_.data(); A.mSeq = fMSeq_.data(); A.mVal = fMVal_.data();
    A.mMin = fMMin_.data(); A.regBaseSum = fBaseSum_.data(); A.regCall0 = fCall0_.data(); A.blockEnd = fBlockEnd_.data();
    A.blockCnt = fBlockCnt_.data(); A.seqBlockFirst = fSeqBlock_.data();
    A.superEnd = fSuperEnd_.data(); A.superCnt = fSuperCnt_.data(); A.seqSuperFirst = fSeqSuper_.data();
    HostFinalOps ops;
    const int nSync = finalPasses(ops, A);
    out.outCalls = fOutC_.data(); out.outBase = fOutB_.data(); out.syncPos = fPos_.data(); out.seqSyncFirst = fSeqSync_.data();
    out.counters = fCounters_.data(); out.rocTuples = fRoc_.data(); out.nSync = nSync;
    ++finalRuns;
    return 0;
  }
  std::atomic<int64_t> finalRuns{0};

  // resident inputs, emulated: "device" pointers are the page-locked host blocks themselves
  struct Tmpl { const uint8_t* bases; int32_t len; std::vector<uint32_t> t2, nmask; };
  static std::vector<Tmpl>& templates() { static std::vector<Tmpl> t; return t; }
  bool templateLookup(const uint8_t* bases, int32_t len, const uint32_t** t2, const uint32_t** nmask) override {
    for (const Tmpl& t : templates()) if (t.bases == bases && len <= t.len) { *t2 = t.t2.data(); *nmask = t.nmask.data(); return true; }
    return false;
  }
  int residentUpload(const void* pinned, size_t, void** dev, int* eventId) override {
    *dev = const_cast<void*>(pinned);
    *eventId = 0;
    ++residentUploads;
    return 0;
  }
  std::atomic<int64_t> residentUploads{0};
  int microBuildRun(MicroBuildJob& b, MicroJob& a, MicroJob& bj, const MicroConfig& mc) override {
    const PacketView pv = b.resident ? b.rv : microBuildView(b, b.slab);
    std::vector<uint8_t> pkStore[2];
    std::vector<uint32_t> offStore[2];
    buildRunT<MicroA>(b, pv, 0, a, mc, pkStore[0], offStore[0]);
    buildRunT<MicroB>(b, pv, b.nA, bj, mc, pkStore[1], offStore[1]);
    if (b.scatter) {
      const RegRange* regs = reinterpret_cast<const RegRange*>(b.slab + b.regions);
      const int32_t* regIdx = reinterpret_cast<const int32_t*>(b.slab + b.regIdx);
      if (a.n > 0) scatterJob(a, regs, regIdx, pkStore[0].data(), offStore[0].data());
      if (bj.n > 0) scatterJob(bj, regs + b.nA, regIdx + b.nA, pkStore[1].data(), offStore[1].data());
    }
    microRegions += a.n + bj.n;
    stats.launches += 2;
    return 0;
  }

  int setAlleles(int side, const AlleleTable& t) override { al_copy_[side] = t; al_[side] = &al_copy_[side]; return 0; }

  int uploadPass(PassData& pd) override { return resetPass(pd); }
  int resetPass(PassData& pd) override {
    for (int side = 0; side < 2; ++side) {
      pd.side[side].decision.assign(pd.side[side].n(), 0);
      pd.side[side].weight.assign(pd.side[side].n(), 0.0);
    }
    return 0;
  }
  int stage(const std::vector<RegionDesc>&, const std::vector<uint8_t>&, const std::vector<int32_t>&, size_t) override { return 0; }
  int orient(PassData& pd) override {
    for (int side = 0; side < 2; ++side) {
      HostSide& hs = pd.side[side];
      const size_t n = hs.n();
      hs.oOff.assign(n + 1, 0);
      int mo = 0;
      for (size_t v = 0; v < n; ++v) {
        VariantGt g = gtOf(hs, side, v);
        OrientCount cnt;
        orientVariant(pd.kind[side], pd.H, g, cnt);
        hs.oOff[v + 1] = hs.oOff[v] + cnt.n;
        mo = std::max(mo, cnt.n);
      }
      if (mo > pd.maxOrient[side]) { err_ = "orientation bound violated"; return VCE_ERR_UNSUPPORTED; }
      const size_t rows = hs.oOff[n];
      oSlot_[side].assign(rows * pd.H + 1, -1);
      hs.oIds.assign(rows * 4 + 4, -2);
      hs.oOrig.assign(rows + 1, 0);
      for (size_t v = 0; v < n; ++v) {
        VariantGt g = gtOf(hs, side, v);
        OrientFill f{hs.oOff[v], pd.H, g.slot0, g.nSlots, g.aNull, oSlot_[side].data(), hs.oIds.data(), hs.oOrig.data()};
        orientVariant(pd.kind[side], pd.H, g, f);
      }
    }
    return 0;
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
    pendOut_.assign(regs.size(), RegionResult());
    pendWarns_.clear();
    for (int tier = 0; tier <= kTierGeneral; ++tier) {
      if (tierFirst[tier + 1] > tierFirst[tier]) {
        runTier(pd, sc, regs, (size_t) tierFirst[tier], (size_t) tierFirst[tier + 1], windows, tier, pendOut_, syncOff, syncTotal, pendWarns_);
      }
    }
    return 0;
  }
  int searchFinish(std::vector<RegionResult>& out, std::vector<WarnRec>& warns) override {
    out = pendOut_;
    warns.insert(warns.end(), pendWarns_.begin(), pendWarns_.end());
    return 0;
  }
  std::vector<RegionResult> pendOut_;
  std::vector<WarnRec> pendWarns_;

  int runTier(PassData& pd, const SearchConfig& sc, const std::vector<RegionDesc>& regs, size_t q0, size_t q1,
              const std::vector<uint8_t>& windows, int tier, std::vector<RegionResult>& out, const std::vector<int32_t>& syncOff,
              size_t syncTotal, std::vector<WarnRec>& warns) {
    const bool general = tier >= kTierMid;
    const FastShape fs = fastShape(general ? 0 : tier);
    ++stats.launches;
    std::vector<int32_t>& syncOut = sync_[pd.pass];
    if (syncOut.size() < syncTotal) syncOut.resize(syncTotal, 0);
    for (size_t q = q0; q < q1; ++q) {
      if ((tier == kTierMid || tier == kTierCta) && !forceGeneral) runRegionCta(pd, sc, regs, q, windows, tier, out, syncOff, warns, syncOut);
      else if (general || forceGeneral) runRegion<true>(pd, sc, regs, q, windows, tier, general, fs, out, syncOff, warns, syncOut);
      else runRegion<false>(pd, sc, regs, q, windows, tier, general, fs, out, syncOff, warns, syncOut);
    }
    return 0;
  }
  // The CTA kernel's source (vce_cta.h) with one lane: "shared memory" and the HBM arena are plain host buffers.
  void runRegionCta(PassData& pd, const SearchConfig& sc, const std::vector<RegionDesc>& regs, size_t q,
                    const std::vector<uint8_t>& windows, int tier, std::vector<RegionResult>& out,
                    const std::vector<int32_t>& syncOff, std::vector<WarnRec>& warns, std::vector<int32_t>& syncOut) {
    RegionConst rk;
    CtaSearch<HostLaneMove> cs(&rk);
    cs.rs.R = regs[q];
    cs.rs.cfg = sc;
    for (int side = 0; side < 2; ++side) {
      HostSide& hs = pd.side[side];
      SideArrays& S = cs.rs.S[side];
      S.vStart = hs.vStart.data(); S.vEnd = hs.vEnd.data(); S.oOff = hs.oOff.data(); S.oSlot = oSlot_[side].data();
      S.oIds = hs.oIds.data(); S.oOrig = hs.oOrig.data();
      S.aStart = al_[side]->aStart.data(); S.aEnd = al_[side]->aEnd.data(); S.aNtOff = al_[side]->aNtOff.data();
      S.nt = al_[side]->nt.data(); S.vFlags = hs.vFlags.data();
      S.outDecision = hs.decision.data(); S.outWeight = hs.weight.data();
    }
    cs.rs.win = windows.data() + regs[q].winOff;
    const int mo = std::max(pd.maxOrient[0], pd.maxOrient[1]);
    cs.rs.maxChildren = 1 + mo;
    cs.shape = ctaShape(tier);
    if (ctaNbOverride > 0) cs.shape.nb = ctaNbOverride;
    const bool layoutOk = ctaLayout(cs.rs.L, pd.H, regs[q].cCount, regs[q].bCount, regs[q].decBits ? regs[q].decBits : 4,
                                     ctaNarrowRanks(regs[q].cSlotN, regs[q].bSlotN));
    std::vector<uint32_t> sm((size_t) ctaSmemWords(cs.shape) + 64);
    cs.carve(sm.data());
    std::vector<uint16_t> ranks((size_t) regs[q].cSlotN + regs[q].bSlotN + 2);
    cs.rank[0] = ranks.data();
    cs.rank[1] = ranks.data() + regs[q].cSlotN;
    long long pool = std::min<long long>(cs.shape.capacity(), (long long) sc.maxPaths + 1 + cs.rs.maxChildren) + 64;
    if (ctaPoolOverride > 0) pool = ctaPoolOverride;
    const long long words = ctaArenaWords(cs.shape, cs.rs.L.W, pool);
    std::vector<int32_t> arena((size_t) words + 16);
    cs.bindArena(arena.data(), words);
    std::vector<WarnRec> lw(64);
    cs.rs.warnOut = lw.data(); cs.rs.warnCap = 64; cs.rs.regionId = (int) q;
    cs.rs.usedPrefix = cs.rs.nWarn = cs.rs.maxFrontier = cs.rs.totalIters = 0;
    cs.computeRanks();
    int resultSlot = -1;
    const int st = layoutOk ? cs.run(resultSlot) : kRegionOverflow;
    RegionResult rr;
    std::memset(&rr, 0, sizeof(rr));
    rr.status = st;
    rr.usedPrefix = cs.rs.usedPrefix;
    rr.nWarn = cs.rs.nWarn;
    rr.maxFrontier = cs.rs.maxFrontier;
    rr.iterations = cs.rs.totalIters;
    rr.lastBaseAlleleId = kPrefixEmpty;
    if (st == kRegionClean || st == kRegionFinished) {
      cs.rs.writeResults(resultSlot, syncOut.data() + syncOff[q], &rr);
      for (int i = 0; i < cs.rs.nWarn && i < 64; ++i) warns.push_back(lw[i]);
    }
    ++ctaRegions;
    if (st == kRegionOverflow) ++ctaOverflows;
    if (std::getenv("VCE_EMUL_DEBUG")) {
      fprintf(stderr, "[emul] cta tier %d region %zu: %dx%d status %d iterations %d maxFrontier %d flags %d layoutOk %d W %d SMAX %d\n", tier, q,
              regs[q].cCount, regs[q].bCount, st, cs.rs.totalIters, cs.rs.maxFrontier, cs.rs.ctl[RegionSearch<HostLaneMove, false>::CTL_FLAGS],
              (int) layoutOk, cs.rs.L.W, cs.rs.L.SMAX);
      fprintf(stderr, "[emul]    block cache: %lld hits %lld misses; inserts into the front block %lld, elsewhere %lld\n", cs.statHit, cs.statMiss,
              cs.statInsFront, cs.statInsFar);
    }
    out[q] = rr;
  }
  int ctaNbOverride = 0;
  long long ctaPoolOverride = 0;
  std::atomic<int64_t> ctaRegions{0}, ctaOverflows{0};
  template <bool CH>
  void runRegion(PassData& pd, const SearchConfig& sc, const std::vector<RegionDesc>& regs, size_t q,
                 const std::vector<uint8_t>& windows, int tier, bool general, const FastShape& fs, std::vector<RegionResult>& out,
                 const std::vector<int32_t>& syncOff, std::vector<WarnRec>& warns, std::vector<int32_t>& syncOut) {
      RegionConst rk;
      RegionSearch<HostLane, CH> rs(&rk);
      rs.R = regs[q];
      rs.cfg = sc;
      for (int side = 0; side < 2; ++side) {
        HostSide& hs = pd.side[side];
        SideArrays& S = rs.S[side];
        S.vStart = hs.vStart.data(); S.vEnd = hs.vEnd.data(); S.oOff = hs.oOff.data(); S.oSlot = oSlot_[side].data();
        S.oIds = hs.oIds.data(); S.oOrig = hs.oOrig.data();
        S.aStart = al_[side]->aStart.data(); S.aEnd = al_[side]->aEnd.data(); S.aNtOff = al_[side]->aNtOff.data();
        S.nt = al_[side]->nt.data(); S.vFlags = hs.vFlags.data();
        S.outDecision = hs.decision.data(); S.outWeight = hs.weight.data();
      }
      rs.win = windows.data() + regs[q].winOff;
      const int mo = std::max(pd.maxOrient[0], pd.maxOrient[1]);
      if (general || forceGeneral) {
        rs.L.init(pd.H, regs[q].qMax > 0 ? regs[q].qMax : 1, std::max(1, std::max(regs[q].cCount, regs[q].bCount)),
                  regs[q].cCount + regs[q].bCount + 2, regs[q].decBits ? regs[q].decBits : 4);
        rs.maxChildren = 1 + mo;
        rs.cap = sc.maxPaths + 1 + rs.maxChildren;
        if (tier == kTierMid && !forceGeneral) rs.cap = std::min(rs.cap, kMidCap);
        // grow the arena lazily in the emulation: start small, retry bigger on overflow
      } else {
        rs.L.init(pd.H, fs.q, fs.kv, fs.smax, 4);
        rs.maxChildren = fs.maxChildren;
        rs.cap = fs.cap;
      }
      int capTry = general ? std::min(rs.cap, 256) : rs.cap;
      const int capFull = rs.cap;
      RegionResult rr;
      while (true) {
        rs.cap = capTry;
        rs.nSlots = rs.cap + 1 + rs.maxChildren;
        rs.ordCap = rs.chunked ? RegionSearch<HostLane, CH>::chunkWords(rs.cap) : 2 * rs.cap + 2;
        std::vector<int32_t> slots((size_t) (rs.nSlots + 2) * rs.L.W), order(rs.ordCap), freeList(rs.nSlots), ctl(16 + 2 * rs.maxChildren);
        rs.slots = slots.data(); rs.order = order.data(); rs.freeList = freeList.data(); rs.ctl = ctl.data();
        std::vector<WarnRec> lw(64);
        rs.warnOut = lw.data(); rs.warnCap = 64; rs.regionId = (int) q;
        int resultSlot = -1;
        const int st = rs.run(resultSlot);
        if (st == kRegionOverflow && general && capTry < capFull) { capTry = std::min(capFull, capTry * 8); continue; }
        std::memset(&rr, 0, sizeof(rr));
        rr.status = st;
        rr.usedPrefix = rs.usedPrefix;
        rr.nWarn = rs.nWarn;
        rr.maxFrontier = rs.maxFrontier;
        rr.iterations = rs.totalIters;
        rr.lastBaseAlleleId = kPrefixEmpty;
        if (st == kRegionClean || st == kRegionFinished) {
          rs.writeResults(resultSlot, syncOut.data() + syncOff[q], &rr);
          for (int i = 0; i < rs.nWarn && i < 64; ++i) warns.push_back(lw[i]);
        }
        break;
      }
      out[q] = rr;
  }
  int fetchPass(PassData& pd, std::vector<int32_t>& syncBuf) override {
    std::vector<int32_t>& so = sync_[pd.pass];
    for (size_t i = 0; i < syncBuf.size() && i < so.size(); ++i) syncBuf[i] = so[i];
    return 0;
  }
  const char* lastError() const override { return err_.c_str(); }

 private:
  VariantGt gtOf(const HostSide& hs, int side, size_t v) const {
    VariantGt g;
    g.gtPloidy = hs.gtPloidy;
    for (int h = 0; h < 4; ++h) g.gt[h] = hs.gt[v * 4 + h];
    g.phased = hs.vFlags[v] & 1;
    g.slot0 = hs.slot0[v];
    g.nSlots = hs.nSl[v];
    g.aNull = al_[side]->aNull.data();
    return g;
  }
  AlleleTable al_copy_[2];
  const AlleleTable* al_[2] = {nullptr, nullptr};
  std::vector<int32_t> oSlot_[2];
  std::vector<int32_t> sync_[2];
  std::string err_;
};

int g_gap = 0 /* 0 = the product default */, g_margin = 24, g_forceGeneral = 0, g_disableMicro = 0, g_threads = 3, g_chunk = 0, g_split = 1 << 16;
int g_forceCta = 0, g_ctaNb = 0;
long long g_ctaPool = 0;
int64_t g_stats[8];
}  // namespace

extern "C" {
void emul_set_options(int gap, int margin, int forceGeneral) { g_gap = gap; g_margin = margin; g_forceGeneral = forceGeneral; }
// regions, escalated, spilled, prefixReruns, launches of the last emul_vce_submit
void emul_get_stats(int64_t* out) { for (int i = 0; i < 8; ++i) out[i] = g_stats[i]; }

// disableMicro: no thread-per-region kernel; threads: host worker threads; chunk: regions per chunk (0 = automatic)
void emul_set_engine(int disableMicro, int threads, int chunk) { g_disableMicro = disableMicro; g_threads = threads; g_chunk = chunk; }
void emul_set_split(int variantsPerSegment) { g_split = variantsPerSegment; }
int g_deviceBuild = 1;   // the product default
void emul_set_device_build(int on) { g_deviceBuild = on; }

std::vector<int64_t> g_summary;
int emul_vce_submit_summary(int64_t* out, int32_t n_seq) {
  if ((size_t) n_seq * 8 > g_summary.size()) return VCE_ERR_BAD_ARGUMENT;
  std::memcpy(out, g_summary.data(), (size_t) n_seq * 8 * sizeof(int64_t));
  return 0;
}
int g_deviceResults = 1;
int64_t g_finalRuns = 0;
void emul_set_device_results(int on) { g_deviceResults = on; }
int64_t emul_final_runs() { return g_finalRuns; }
// test hook of vce_template_register for the emulation (host copies); len <= 0 forgets every template
void emul_template_register(const uint8_t* bases, int32_t len) {
  std::vector<EmulBackend::Tmpl>& ts = EmulBackend::templates();
  if (len <= 0) { ts.clear(); return; }
  ts.emplace_back();
  ts.back().bases = bases;
  ts.back().len = len;
  packTemplate(bases, len, nullptr, ts.back().t2, ts.back().nmask);
  // slack, as on the device: the builder may look one granule beyond the last base
  ts.back().t2.resize(ts.back().t2.size() + 8, 0);
  ts.back().nmask.resize(ts.back().nmask.size() + 8, 0);
}
// test hook: what WinPacked delivers for one template position (0 = the position's granule holds a non-ACGT base)
uint32_t emul_template_word(const uint8_t* bases, int32_t len, int32_t pos, int32_t) {
  std::vector<uint32_t> t2, nm;
  packTemplate(bases, len, nullptr, t2, nm);
  t2.resize(t2.size() + 8, 0);
  nm.resize(nm.size() + 8, 0);
  const WinPacked w{t2.data(), nm.data(), pos};
  return (uint32_t) w.at(0);
}
int64_t g_residentUploads = 0;
int64_t emul_resident_uploads() { return g_residentUploads; }

// forceCta: every warp-path region starts in the CTA kernel's tiers; nb / pool: shrink its frontier directory / record pool
// (exercises the overflow hand-over to the large-capacity kernel)
void emul_set_cta(int forceCta, int nb, long long pool) { g_forceCta = forceCta; g_ctaNb = nb; g_ctaPool = pool; }

int emul_vce_submit(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_sequence_result* results) {
  EmulBackend be;
  be.forceGeneral = g_forceGeneral != 0;
  be.ctaNbOverride = g_ctaNb;
  be.ctaPoolOverride = g_ctaPool;
  be.resultsOn = g_deviceResults != 0;
  WorkerPool pool(g_threads);
  EngineOptions opt;
  if (g_gap > 0) opt.gap = g_gap;
  opt.margin = g_margin;
  opt.forceGeneral = g_forceGeneral != 0;
  opt.forceCta = g_forceCta != 0;
  opt.disableMicro = g_disableMicro != 0;
  opt.chunkRegions = g_chunk;
  opt.splitSegment = g_split;
  opt.deviceBuild = g_deviceBuild != 0;
  int rc;
  {
    Engine en(&be, &pool, opt);
    rc = en.run(*cfg, n_seq, seqs, results);
    if (rc && !en.error().empty()) fprintf(stderr, "emul: %s\n", en.error().c_str());
    g_summary = en.summary();
  }
  g_stats[0] = be.stats.regions; g_stats[1] = be.stats.escalated; g_stats[2] = be.stats.spilled;
  g_stats[3] = be.stats.prefixReruns; g_stats[4] = be.stats.launches; g_stats[5] = be.microRegions.load(); g_stats[6] = be.microClean.load();
  g_stats[7] = be.ctaRegions.load() | (be.ctaOverflows.load() << 32);
  g_residentUploads = be.residentUploads.load();
  g_finalRuns = be.finalRuns.load();
  return rc;
}

// vce_submit over `n_dev` emulated devices with the sequence splitting of vce_split.h (what vce_capi.cpp does with real devices):
// the items of every "device" go through an engine of their own, one after the other.  opts = {min variants, overlap, gap,
// imbalance * 100, pieces per device, force (no cost estimate)}; out_stats = {sequences cut, pieces, sequences evaluated again in one piece}.
int emul_vce_submit_split(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_sequence_result* results, int32_t n_dev,
                          const int64_t* opts, int64_t* out_stats) {
  vce::SplitOptions so;
  so.minVariants = opts[0]; so.overlap = (int) opts[1]; so.gap = (int) opts[2]; so.imbalance = (double) opts[3] / 100.0;
  so.piecesPerDevice = (int) opts[4];
  so.force = opts[5] != 0;
  vce::SplitPlan plan;
  const auto t0 = std::chrono::steady_clock::now();
  vce::splitPlan(*cfg, n_seq, seqs, results, n_dev, so, plan);
  if (std::getenv("VCE_TIMING")) fprintf(stderr, "[emul] split planning (serial) %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  std::vector<int64_t> summary((size_t) n_seq * 8, 0);
  int hard = 0;
  auto round = [&](std::vector<vce_sequence>& xs, std::vector<vce_sequence_result>& xr, std::vector<int64_t>& xsum) {
    std::vector<int64_t> w(xs.size());
    for (size_t i = 0; i < xs.size(); ++i) w[i] = (int64_t) xs[i].baseline.n + xs[i].calls.n;
    std::vector<int> dev;
    vce::splitAssign(w, n_dev, dev);
    xsum.assign(xs.size() * 8, 0);
    for (int d = 0; d < n_dev; ++d) {
      std::vector<int32_t> idx;
      for (size_t i = 0; i < xs.size(); ++i) if (dev[i] == d) idx.push_back((int32_t) i);
      if (idx.empty()) continue;
      std::vector<vce_sequence> sub(idx.size());
      std::vector<vce_sequence_result> res(idx.size());
      for (size_t i = 0; i < idx.size(); ++i) { sub[i] = xs[(size_t) idx[i]]; res[i] = xr[(size_t) idx[i]]; }
      const int rc = emul_vce_submit(cfg, (int32_t) idx.size(), sub.data(), res.data());
      bool perSequence = false;
      for (size_t i = 0; i < idx.size(); ++i) {
        xr[(size_t) idx[i]] = res[i];
        if (rc && res[i].error == rc) perSequence = true;
        if (g_summary.size() >= (i + 1) * 8) for (int q = 0; q < 8; ++q) xsum[(size_t) idx[i] * 8 + (size_t) q] = g_summary[i * 8 + (size_t) q];
      }
      if (rc && !perSequence && !hard) hard = rc;
    }
  };
  std::vector<vce_sequence> xs;
  std::vector<vce_sequence_result> xr;
  for (int32_t k : plan.whole) { xs.push_back(seqs[k]); xr.push_back(results[k]); }
  for (auto& p : plan.pieces) { xs.push_back(p->view); xr.push_back(p->res); }
  std::vector<int64_t> xsum;
  round(xs, xr, xsum);
  if (hard) return hard;
  for (size_t i = 0; i < plan.whole.size(); ++i) {
    results[plan.whole[i]] = xr[i];
    for (int q = 0; q < 8; ++q) summary[(size_t) plan.whole[i] * 8 + (size_t) q] = xsum[i * 8 + (size_t) q];
  }
  for (size_t i = 0; i < plan.pieces.size(); ++i) plan.pieces[i]->res = xr[plan.whole.size() + i];
  std::vector<int32_t> redo;
  const auto t1 = std::chrono::steady_clock::now();
  vce::splitMergeAll(plan, seqs, results, summary.data(), redo, vce::SplitSerial());
  if (std::getenv("VCE_TIMING")) fprintf(stderr, "[emul] split stitching (serial) %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
  if (out_stats) { out_stats[0] = (int64_t) plan.split.size(); out_stats[1] = (int64_t) plan.pieces.size(); out_stats[2] = (int64_t) redo.size(); }
  if (!redo.empty()) {
    std::vector<vce_sequence> ys;
    std::vector<vce_sequence_result> yr;
    for (int32_t k : redo) { ys.push_back(seqs[k]); yr.push_back(results[k]); }
    std::vector<int64_t> ysum;
    round(ys, yr, ysum);
    if (hard) return hard;
    for (size_t i = 0; i < redo.size(); ++i) {
      results[redo[i]] = yr[i];
      for (int q = 0; q < 8; ++q) summary[(size_t) redo[i] * 8 + (size_t) q] = ysum[i * 8 + (size_t) q];
    }
  }
  g_summary = summary;
  for (int32_t k = 0; k < n_seq; ++k) if (results[k].error) return results[k].error;
  return VCE_OK;
}

// staged job on the emulation: prepare once, execute `runs` times, finish (exercises the re-runnable path)
int emul_vce_job(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, vce_sequence_result* results, int32_t runs) {
  EmulBackend be;
  WorkerPool pool(g_threads);
  EngineOptions opt;
  if (g_gap > 0) opt.gap = g_gap;
  opt.margin = g_margin;
  opt.chunkRegions = g_chunk;
  opt.splitSegment = g_split;
  Engine en(&be, &pool, opt);
  int rc = en.prepare(*cfg, n_seq, seqs, true);
  for (int i = 0; i < runs && !rc; ++i) rc = en.execute();
  if (!rc) rc = en.finish(results);
  if (rc && !en.error().empty()) fprintf(stderr, "emul: %s\n", en.error().c_str());
  return rc;
}
// host-side micro-benchmark: `reps` staged prepare() calls on one engine (warm arenas); stage times with VCE_TIMING=1
int emul_bench_prepare(const vce_config* cfg, int32_t n_seq, const vce_sequence* seqs, int32_t reps) {
  EmulBackend be;
  WorkerPool pool(g_threads);
  EngineOptions opt;
  opt.chunkRegions = g_chunk;
  opt.splitSegment = g_split;
  Engine en(&be, &pool, opt);
  int rc = 0;
  for (int i = 0; i < reps && !rc; ++i) rc = en.prepare(*cfg, n_seq, seqs, true);
  return rc;
}
// Unit-test hook: the product's K1 source (vce_orient.h) for one variant, same signature as oracle_orientations.
int emul_orientations(int32_t kind, int32_t H, int32_t gt_ploidy, const int8_t* gt, int32_t phased, int32_t n_slots,
                      const uint8_t* slot_null, int8_t* out_ids, uint8_t* out_original, int32_t cap) {
  VariantGt g;
  g.gtPloidy = gt_ploidy;
  for (int h = 0; h < 4; ++h) g.gt[h] = h < gt_ploidy ? gt[h] : -2;
  g.phased = phased;
  g.slot0 = 0;
  g.nSlots = n_slots;
  g.aNull = slot_null;
  OrientCount cnt;
  orientVariant(kind, H, g, cnt);
  if (cnt.n > cap) return cnt.n;
  std::vector<int32_t> oSlot((size_t) cnt.n * H + 4);
  OrientFill f{0, H, 0, n_slots, slot_null, oSlot.data(), out_ids, out_original};
  orientVariant(kind, H, g, f);
  return cnt.n;
}
int emul_vce_roc(const vce_roc_in*, vce_roc_out*) { return VCE_ERR_UNSUPPORTED; }

// ---- VCF loader (vce_vcf.h) with the passes run as loops
}  // extern "C"
namespace {
struct HostVcfOps {
  std::vector<std::unique_ptr<uint8_t[]>> blocks;
  template <class T> T* alloc(size_t n) {
    blocks.emplace_back(new uint8_t[n * sizeof(T) + 64]);
    std::memset(blocks.back().get(), 0xcd, n * sizeof(T) + 64);   // poison: every entry read must have been written
    return reinterpret_cast<T*>(blocks.back().get());
  }
  void upload(void* d, const void* h, size_t bytes) { if (bytes) std::memcpy(d, h, bytes); }
  void download(void* h, const void* d, size_t bytes) { if (bytes) std::memcpy(h, d, bytes); }
  template <class F> void parFor(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
  void exclusiveSum(int32_t* in, int32_t* out, int n) { int32_t acc = 0; for (int i = 0; i < n; ++i) { const int32_t v = in[i]; out[i] = acc; acc += v; } }
  int fetchInt(const int32_t* p) { return *p; }
  void mark(int) {}
  // one lane per "warp": the very same inflateBlock / CRC source as k_bgzf_inflate / k_bgzf_crc
  void inflate(int nb, const vce::BgzfBlock* blocks, const uint8_t* data, uint8_t* out, int32_t* status, bool checkCrc) {
    std::unique_ptr<vce::InflateWs> ws(new vce::InflateWs());
    uint32_t table[256];
    for (uint32_t i = 0; i < 256; ++i) table[i] = vce::crc32TableEntry(i);
    for (int b = 0; b < nb; ++b) {
      std::memset(ws.get(), 0xcd, sizeof(vce::InflateWs));
      status[b] = vce::inflateBlock<vce::HostWarp>(data + blocks[b].src, blocks[b].srcLen, out + blocks[b].dst, blocks[b].dstLen, ws.get());
      if (checkCrc && status[b] == vce::kInfOk && vce::crc32Bytes(out + blocks[b].dst, blocks[b].dstLen, table) != blocks[b].crc) status[b] = vce::kInfCrc;
    }
  }
  static void hostFree(uint8_t* p, size_t) { delete[] p; }
  uint8_t* hostAlloc(size_t bytes, void (**release)(uint8_t*, size_t)) {
    *release = &hostFree;
    uint8_t* p = new uint8_t[bytes];
    std::memset(p, 0xcd, bytes);
    return p;
  }
};
std::string g_vcfErr;
}  // namespace
extern "C" {

int emul_vce_vcf_parse(const vce_vcf_in* in, void** result) {
  *result = nullptr;
  vce::VcfConfig cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.sample = in->sample; cfg.ploidy = in->ploidy; cfg.passOnly = in->pass_only; cfg.maxLength = in->max_length;
  cfg.templateLen = in->template_len;
  cfg.refOverlap = in->ref_overlap;
  cfg.evalRegions = in->eval_regions;
  cfg.nEvalRegions = in->eval_regions ? in->n_eval_regions : -1;
  cfg.regionBeg = in->region_begin;
  cfg.regionEnd = in->region_end;
  if (in->score_field && std::strcmp(in->score_field, "QUAL") != 0) {
    const size_t sl = std::strlen(in->score_field);
    if (sl == 0 || sl >= sizeof(cfg.scoreField)) return VCE_ERR_BAD_ARGUMENT;
    std::memcpy(cfg.scoreField, in->score_field, sl);
    cfg.scoreLen = (int32_t) sl;
  }
  if (in->name) {
    const size_t nl = std::strlen(in->name);
    if (nl == 0 || nl >= sizeof(cfg.name)) return VCE_ERR_BAD_ARGUMENT;
    std::memcpy(cfg.name, in->name, nl);
    cfg.nameLen = (int32_t) nl;
  }
  std::unique_ptr<vce::VcfResult> R(new vce::VcfResult());
  HostVcfOps ops;
  const int rc = vce::vcfDrive(ops, reinterpret_cast<const uint8_t*>(in->text), in->text_len, cfg, *R, g_vcfErr);
  if (rc) return rc == 2 ? VCE_ERR_CUDA : VCE_ERR_BAD_ARGUMENT;
  *result = R.release();
  return VCE_OK;
}
const char* emul_vce_vcf_error() { return g_vcfErr.c_str(); }
int emul_vce_vcf_variants(const void* result, vce_variants* out, vce_vcf_stats* stats) {
  const vce::VcfResult& R = *static_cast<const vce::VcfResult*>(result);
  out->n = R.n;
  out->gt_ploidy = R.ploidy;
  out->id = R.id; out->phased = R.phased; out->status_in = R.statusIn; out->gt = R.gt;
  out->allele_off = R.alleleOff;
  out->n_slots = R.nSlots;
  out->allele_start = R.alleleStart; out->allele_end = R.alleleEnd; out->allele_null = R.alleleNull;
  out->allele_nt_off = R.alleleNtOff; out->nt = R.nt;
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->lines = R.stats.lines; stats->records = R.stats.records; stats->kept = R.stats.kept; stats->skipped = R.stats.skipped;
  }
  return VCE_OK;
}
void emul_vce_vcf_free(void* result) { delete static_cast<vce::VcfResult*>(result); }
int emul_vce_vcf_roc_fields(const void* result, const uint32_t** filter_mask, const double** qual) {
  const vce::VcfResult& R = *static_cast<const vce::VcfResult*>(result);
  *filter_mask = R.rocMask;
  *qual = R.qual;
  return VCE_OK;
}

// ---- on-disk formats (vce_bgzf.h) through the host ops
static int emulVcfConfig(const vce_vcf_in* in, vce::VcfConfig& cfg) {
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.sample = in->sample; cfg.ploidy = in->ploidy; cfg.passOnly = in->pass_only; cfg.maxLength = in->max_length;
  cfg.templateLen = in->template_len;
  cfg.refOverlap = in->ref_overlap;
  cfg.evalRegions = in->eval_regions;
  cfg.nEvalRegions = in->eval_regions ? in->n_eval_regions : -1;
  cfg.regionBeg = in->region_begin;
  cfg.regionEnd = in->region_end;
  if (in->score_field && std::strcmp(in->score_field, "QUAL") != 0) {
    const size_t sl = std::strlen(in->score_field);
    if (sl == 0 || sl >= sizeof(cfg.scoreField)) return VCE_ERR_BAD_ARGUMENT;
    std::memcpy(cfg.scoreField, in->score_field, sl);
    cfg.scoreLen = (int32_t) sl;
  }
  if (in->name) {
    const size_t nl = std::strlen(in->name);
    if (nl == 0 || nl >= sizeof(cfg.name)) return VCE_ERR_BAD_ARGUMENT;
    std::memcpy(cfg.name, in->name, nl);
    cfg.nameLen = (int32_t) nl;
  }
  return VCE_OK;
}
int emul_vce_bgzf_text_size(const uint8_t* data, int64_t len, int64_t vb, int64_t ve, int64_t* text_len) {
  g_vcfErr = vce::bgzfTextSize(data, len, vb, ve, text_len);
  return g_vcfErr.empty() ? VCE_OK : VCE_ERR_BAD_ARGUMENT;
}
int emul_vce_bgzf_inflate(const uint8_t* data, int64_t len, int64_t vb, int64_t ve, int32_t check_crc, uint8_t* text, int64_t text_cap,
                          int64_t* text_len, vce_bgzf_stats* stats) {
  HostVcfOps ops;
  const uint8_t* dText = nullptr;
  std::vector<uint8_t> out;
  vce::BgzfStats st;
  const int rc = vce::bgzfDrive(ops, data, len, vb, ve, check_crc != 0, &dText, text_len, &out, st, g_vcfErr);
  if (rc) return rc == 2 ? VCE_ERR_CUDA : VCE_ERR_BAD_ARGUMENT;
  if (stats) { std::memset(stats, 0, sizeof(*stats)); stats->blocks = st.blocks; stats->compressed_bytes = st.compressed; stats->text_bytes = st.text; }
  if ((int64_t) out.size() > text_cap) return VCE_ERR_CAPACITY;
  if (!out.empty()) std::memcpy(text, out.data(), out.size());
  return VCE_OK;
}
int emul_vce_vcf_parse_bgzf(const vce_vcf_in* in, int64_t vb, int64_t ve, int32_t check_crc, void** result, vce_bgzf_stats* stats) {
  *result = nullptr;
  vce::VcfConfig cfg;
  if (emulVcfConfig(in, cfg)) return VCE_ERR_BAD_ARGUMENT;
  std::unique_ptr<vce::VcfResult> R(new vce::VcfResult());
  HostVcfOps ops;
  const uint8_t* dText = nullptr;
  int64_t textLen = 0;
  vce::BgzfStats st;
  int rc = vce::bgzfDrive(ops, reinterpret_cast<const uint8_t*>(in->text), in->text_len, vb, ve, check_crc != 0, &dText, &textLen, nullptr, st, g_vcfErr);
  if (rc == 0) rc = vce::vcfDrive(ops, nullptr, textLen, cfg, *R, g_vcfErr, dText);
  if (rc) return rc == 2 ? VCE_ERR_CUDA : VCE_ERR_BAD_ARGUMENT;
  if (stats) { std::memset(stats, 0, sizeof(*stats)); stats->blocks = st.blocks; stats->compressed_bytes = st.compressed; stats->text_bytes = st.text; }
  *result = R.release();
  return VCE_OK;
}
int emul_vce_tabix_open(const uint8_t* tbi, int64_t len, void** index) {
  *index = nullptr;
  if (len < 18 || tbi[0] != 0x1f || tbi[1] != 0x8b) { g_vcfErr = "not a valid TABIX index. (index file is not block compressed)"; return VCE_ERR_BAD_ARGUMENT; }
  std::unique_ptr<vce::TabixIndex> t(new vce::TabixIndex());
  HostVcfOps ops;
  const uint8_t* dText = nullptr;
  int64_t textLen = 0;
  vce::BgzfStats st;
  if (vce::bgzfDrive(ops, tbi, len, 0, -1, true, &dText, &textLen, &t->raw, st, g_vcfErr)) return VCE_ERR_BAD_ARGUMENT;
  g_vcfErr = vce::tabixParse(*t);
  if (!g_vcfErr.empty()) return VCE_ERR_BAD_ARGUMENT;
  *index = t.release();
  return VCE_OK;
}
int emul_vce_tabix_info(const void* index, int32_t out[7]) {
  const vce::TabixIndex& t = *static_cast<const vce::TabixIndex*>(index);
  out[0] = (int32_t) t.names.size(); out[1] = t.format; out[2] = t.seqCol; out[3] = t.begCol; out[4] = t.endCol; out[5] = t.meta; out[6] = t.skip;
  return VCE_OK;
}
const char* emul_vce_tabix_name(const void* index, int32_t i) {
  const vce::TabixIndex& t = *static_cast<const vce::TabixIndex*>(index);
  return i >= 0 && i < (int32_t) t.names.size() ? t.names[(size_t) i].c_str() : nullptr;
}
int emul_vce_tabix_query(const void* index, const char* name, int32_t b0, int32_t e0, int64_t* vb, int64_t* ve) {
  *vb = *ve = -1;
  const int r = vce::tabixQuery(*static_cast<const vce::TabixIndex*>(index), name, b0, e0, vb, ve, g_vcfErr);
  if (r < 0) return VCE_ERR_BAD_ARGUMENT;
  if (r == 0) *vb = *ve = -1;
  return VCE_OK;
}
void emul_vce_tabix_close(void* index) { delete static_cast<vce::TabixIndex*>(index); }
// SDF bit planes: the granule / mask bodies as loops; the packed words and the mask come back for comparison with packTemplate
int emul_vce_sdf_decode(const uint8_t* seqdata, int64_t seqdata_len, int64_t first, int32_t len, uint8_t* bases_out, uint32_t* packed_out,
                        uint32_t* nmask_out) {
  const int64_t g0 = first >> 6, rel = first - (g0 << 6), need = ((rel + len + 63) >> 6) * 24;
  if (len <= 0 || first < 0 || g0 * 24 + need > seqdata_len) return VCE_ERR_BAD_ARGUMENT;
  const int32_t granules = (len + 15) / 16, mwords = (granules + 31) / 32;
  std::vector<uint8_t> flag((size_t) granules);
  for (int32_t g = 0; g < granules; ++g) vce::sdfGranule(seqdata + g0 * 24, rel, len, g, packed_out, flag.data(), bases_out);
  for (int32_t m = 0; m < mwords; ++m) vce::sdfMaskWord(flag.data(), granules, m, nmask_out);
  return VCE_OK;
}
// packTemplate (vce_engine.cpp) for the comparison above
int emul_vce_pack_template(const uint8_t* bases, int32_t len, uint32_t* packed_out, uint32_t* nmask_out) {
  std::vector<uint32_t> p, m;
  vce::packTemplate(bases, len, nullptr, p, m);
  std::memcpy(packed_out, p.data(), p.size() * 4);
  std::memcpy(nmask_out, m.data(), m.size() * 4);
  return VCE_OK;
}
}
