// vce_pipeline.h -- the interface between the host engine (vce_engine.h) and the device work: flat host-side arrays
// of a batch of regions, launch descriptors, and `Backend`.
//
// The product library links exactly one backend, the CUDA one (vce_cuda.cu).  tests/emul builds a host
// instantiation of the same search sources to exercise the host logic where no GPU exists; it is never linked
// into the product.
#ifndef VCE_PIPELINE_H_
#define VCE_PIPELINE_H_

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/vce.h"
#include "vce_device.h"
#include "vce_micro.h"
#include "vce_packet.h"
#include "vce_search.h"
#include "vce_cta.h"
#include "vce_final.h"

namespace vce {

struct AlleleTable {  // one side, whole batch
  std::vector<int32_t> aStart, aEnd, aNtOff;
  std::vector<uint8_t> aNull, nt;
};

struct HostSide {  // one side of one pass, whole batch, sorted by (sequence, start, end) -- stable
  int gtPloidy = 0;
  std::vector<int32_t> vStart, vEnd, slot0, nSl;
  std::vector<uint8_t> vFlags;      // bit0 phased, bit4 outside-eval (== VCE_STATUS_OUTSIDE_EVAL)
  std::vector<int8_t> gt;           // n * 4
  std::vector<int32_t> inputIdx;    // index in the caller's arrays of that sequence
  std::vector<int32_t> seqFirst;    // [nSeq + 1]
  // K1 output (host copy)
  std::vector<int32_t> oOff;
  std::vector<int8_t> oIds;
  std::vector<uint8_t> oOrig;
  // K2/K3 output
  std::vector<uint8_t> decision;
  std::vector<double> weight;
  size_t n() const { return vStart.size(); }
};

struct PassData {
  bool wantOrientTables = true;   // fetchPass also copies the orientation tables (oOff / oIds / oOrig) back
  int pass = 0;
  int H = 2;
  int kind[2] = {0, 0};  // [0] = calls orientor, [1] = baseline orientor
  HostSide side[2];      // [0] = calls, [1] = baseline
  int maxOrient[2] = {0, 0};
};

struct SearchStats {
  int64_t launches = 0;
  double searchMs = 0;   // device time of the search kernels
  double deviceMs = 0;   // device time of everything launched
  int64_t regions = 0, escalated = 0, spilled = 0, prefixReruns = 0;
  int64_t h2dBytes = 0, d2hBytes = 0;
  int64_t iterations = 0;
};

// One chunk of thread-per-region work (vce_micro.h): packets in, fixed-size result records out.
struct MicroJob {
  int cls = 0;                        // 0 = MicroA, 1 = MicroB
  int n = 0;                          // regions in the chunk
  const uint8_t* packets = nullptr;   // host buffer (from Backend::hostAlloc), packetBytes long
  size_t packetBytes = 0;
  const uint32_t* offsets = nullptr;  // [n] offset of each packet in 4-byte words (host, from hostAlloc)
  uint8_t* results = nullptr;         // [n * resBytes] host (from hostAlloc)
  int resBytes = 0;
  // backend-private
  void* dPackets = nullptr;
  void* dOffsets = nullptr;
  void* dResults = nullptr;
  int lane = 0;
  int event = -1;
  void* evHandle = nullptr;
};

// ---- device-side packet building (vce_packet.h).  One slab per chunk: wholesale copies of the slices of the caller's
// arrays that the chunk's regions touch, the regions themselves and their reference windows.  Offsets are in bytes from
// the start of the slab, every array starts on a 16-byte boundary.
struct MicroBuildSide {
  int64_t vStart, vEnd;          // int32[nV + 1]   element 0 = pass-wide variant index v0 (the extra one: the next variant)
  int64_t gt;                    // int8[nV * gtPloidy]   element 0 = input index i0
  int64_t phased, statusIn;      // uint8[nV] or -1 when the caller passed no such array
  int64_t alleleOff;             // int32[nV + 1]
  int64_t aStart, aEnd;          // int32[nSlots]   element 0 = slot slot0
  int64_t aNull;                 // uint8[nSlots]
  int64_t aNtOff;                // int32[nSlots + 1]
  int64_t nt;                    // uint8[nNt]      element 0 = nt offset nt0
  int32_t v0, i0, slot0, nt0;
  int32_t sideEnd, first, gtPloidy;
};
struct MicroBuildJob {
  const uint8_t* slab = nullptr;   // host buffer (Backend::hostAlloc)
  size_t slabBytes = 0;
  MicroBuildSide side[2];
  int32_t tl = 0, H = 0, expectPloidy[2] = {0, 0}, margin = 0;
  // resident mode (registered template + whole-sequence upload, Backend::residentUpload): the slab holds the regions only,
  // `rv` carries device pointers to the sequence's arrays, the windows are cut out of the device copy of the template
  bool resident = false;
  PacketView rv;
  const uint32_t* t2 = nullptr;
  const uint32_t* nmask = nullptr;
  int waitEvents[2] = {-1, -1};    // uploads (Backend::residentUpload) the builder has to wait for
  // device-resident results (Backend::resultsBegin): the chunk's result records are scattered into the pass-wide arrays on
  // the device instead of being copied back; regIdx = slab offset of int32[nA + nB], the global region index of each region
  bool scatter = false;
  int64_t regIdx = 0;
  int64_t regions = 0;             // RegRange[nA + nB]: the MicroA-shaped regions, then the MicroB-shaped ones
  int64_t winOff = 0;              // uint32[nA + nB]: offset of each region's window bases within `windows`
  int64_t windows = 0;             // uint8[]: reference bases from each region's window start
  int32_t nA = 0, nB = 0;
};
// The packet builder's view of a slab that lives at `base` (host or device address).
inline PacketView microBuildView(const MicroBuildJob& b, const uint8_t* base) {
  PacketView pv;
  for (int side = 0; side < 2; ++side) {
    const MicroBuildSide& m = b.side[side];
    PacketSideView& s = pv.s[side];
    s.vStart = reinterpret_cast<const int32_t*>(base + m.vStart) - m.v0;
    s.vEnd = reinterpret_cast<const int32_t*>(base + m.vEnd) - m.v0;
    s.sideEnd = m.sideEnd;
    s.first = m.first;
    s.idx = nullptr;
    s.gtPloidy = m.gtPloidy;
    s.gt = reinterpret_cast<const int8_t*>(base + m.gt) - (ptrdiff_t) m.i0 * m.gtPloidy;
    s.phased = m.phased >= 0 ? base + m.phased - m.i0 : nullptr;
    s.statusIn = m.statusIn >= 0 ? base + m.statusIn - m.i0 : nullptr;
    s.alleleOff = reinterpret_cast<const int32_t*>(base + m.alleleOff) - m.i0;
    s.aStart = reinterpret_cast<const int32_t*>(base + m.aStart) - m.slot0;
    s.aEnd = reinterpret_cast<const int32_t*>(base + m.aEnd) - m.slot0;
    s.aNull = base + m.aNull - m.slot0;
    s.aNtOff = reinterpret_cast<const int32_t*>(base + m.aNtOff) - m.slot0;
    s.nt = base + m.nt - m.nt0;
    pv.expectPloidy[side] = b.expectPloidy[side];
  }
  pv.tl = b.tl;
  pv.H = b.H;
  pv.margin = b.margin;
  return pv;
}

// ---- device-resident results (vce_final.h): what the engine tells the backend about the pass, the regions the host
// settled, and what comes back
struct ResultsSeq {
  int32_t first[2] = {0, 0}, n[2] = {0, 0};     // slices of the pass-wide per-variant arrays; side 0 = calls, 1 = baseline
  int32_t regFirst = 0, nReg = 0;               // slice of the region table
  const int32_t* vStart[2] = {nullptr, nullptr};     // device pointers (Backend::residentUpload), sequence-local index
  const uint8_t* phased[2] = {nullptr, nullptr};
  const uint8_t* statusIn[2] = {nullptr, nullptr};
};
struct ResultsSetup {
  int H = 2, kind[2] = {0, 0};
  std::vector<ResultsSeq> seq;
  int32_t nVar[2] = {0, 0};
  int32_t nReg = 0;
  bool rocTuples = false;
};
struct ResultsPatch {     // one entry per region whose final state is the host's: settled by the warp / CTA kernels, merged, dead
  std::vector<int32_t> regIdx, syncOff, syncCnt, payOff, wOff;
  std::vector<uint32_t> info;       // riPack(RI_HOST | RI_DEAD, flags, lastId)
  std::vector<RegRange> range;      // the variants the region owns (empty for dead regions)
  std::vector<uint8_t> decPay;      // decisions, calls then baseline, region by region
  std::vector<double> wPay;         // call weights, region by region
  std::vector<int32_t> extSync;     // sync points, region by region
};
struct ResultsOut {       // host copies, valid until the next resultsBegin
  const uint8_t* outCalls = nullptr;      // [nVar[0] * 16] status, row (0xff none), original, pad, syncId i32, weight f64
  const uint8_t* outBase = nullptr;       // [nVar[1] * 8]
  const int32_t* syncPos = nullptr;       // Path.getSyncPoints() of every sequence, back to back
  const int32_t* seqSyncFirst = nullptr;  // [nSeq + 1]
  const int64_t* counters = nullptr;      // [nSeq * 8] baseline TP, FN, call TP, FP, skipped, mis-, correct, unphaseable phasings
  const double* rocTuples = nullptr;      // [nVar[0] * 3] when requested
  int nSync = 0;
};

class Backend {
 public:
  virtual ~Backend() {}
  // ---- device-resident results
  virtual bool resultsSupported() const { return false; }
  virtual int resultsBegin(const ResultsSetup&) { return VCE_ERR_UNSUPPORTED; }
  // Blocks until every chunk handed over so far has been searched and scattered; info[r] = riPack(...) of every region.
  virtual int resultsRegInfo(const uint32_t** /*info*/, int64_t* /*iterations*/) { return VCE_ERR_UNSUPPORTED; }
  // Applies the patch, runs the final passes and copies the results to the host.
  virtual int resultsFinish(const ResultsPatch&, ResultsOut&) { return VCE_ERR_UNSUPPORTED; }
  // ---- thread-per-region path.  Calls for different jobs may come from different host threads.
  virtual void* hostAlloc(size_t bytes) = 0;   // page-locked on the CUDA backend
  virtual void hostFree(void* p) = 0;
  virtual int microReset() = 0;                               // forgets the device buffers of earlier jobs
  virtual void microMark() {}                                 // remembers the device buffers allocated so far ...
  virtual void microRewind() {}                               // ... and forgets everything allocated after the mark
  virtual int microUpload(MicroJob& j) = 0;                   // async host -> device copy of packets and offsets
  virtual int microRun(MicroJob& j, const MicroConfig& mc) = 0;  // async kernel + device -> host copy of the results
  // one kernel launch over all (uploaded) jobs + one async result copy per job: the staged / HBM-resident path
  virtual int microRunAll(std::vector<MicroJob*>& jobs, const MicroConfig& mc) = 0;
  virtual int microWait(MicroJob& j) = 0;                     // blocks until the results of j are on the host
  virtual void microCollect() {}                              // reads the device times of finished microRunAll launches
  // Uploads the slab, builds the packets of its regions on the device (k_build), runs the search kernels and starts the
  // device -> host copies of the results: a.n == b.nA records for the MicroA-shaped regions, bj.n == b.nB for the others.
  // A region that does not fit a packet comes back with status kMicroFallback.  Asynchronous; microWait() collects.
  virtual int microBuildRun(MicroBuildJob& b, MicroJob& a, MicroJob& bj, const MicroConfig& mc) = 0;
  // ---- resident inputs.  A registered template (vce_template_register) lives on the device as 2 bits per base + a map of
  // the 16-base granules that hold anything else; false when `bases` is not registered on this backend's device.
  virtual bool templateLookup(const uint8_t* /*bases*/, int32_t /*len*/, const uint32_t** /*t2*/, const uint32_t** /*nmask*/) { return false; }
  // Asynchronous host -> device copy of one block of page-locked memory (from hostAlloc); *dev is valid right away for
  // pointer arithmetic, the data once event *eventId has passed (MicroBuildJob::waitEvents).
  virtual int residentUpload(const void* /*pinned*/, size_t /*bytes*/, void** /*dev*/, int* /*eventId*/) { return VCE_ERR_UNSUPPORTED; }
  // Caller memory that was page-locked with vce_host_register goes to the device without a staging copy:
  // hostPinned = the whole range lies inside a registered block; residentAlloc + residentCopy = one device block, filled by
  // asynchronous copies straight from the caller's arrays (event as for residentUpload, recorded after the last copy of a lane).
  virtual bool hostPinned(const void* /*p*/, size_t /*bytes*/) const { return false; }
  virtual int residentAlloc(size_t /*bytes*/, void** /*dev*/, int* /*lane*/) { return VCE_ERR_UNSUPPORTED; }
  virtual int residentCopy(void* /*dst*/, const void* /*src*/, size_t /*bytes*/, int /*lane*/) { return VCE_ERR_UNSUPPORTED; }
  virtual int residentMark(int /*lane*/, int* /*eventId*/) { return VCE_ERR_UNSUPPORTED; }
  // The backend may hold the MicroB-shaped regions of the chunks back and search them in one launch: call this once every
  // chunk has been handed over (the MicroJob objects have to stay where they are until then).
  virtual int microBuildFlush(const MicroConfig&) { return 0; }
  virtual int setAlleles(int side, const AlleleTable& t) = 0;
  // Host -> device copy of the pass's variant arrays; clears the per-variant outputs.
  virtual int uploadPass(PassData& pd) = 0;
  // Clears the per-variant outputs again (a staged job may be executed repeatedly).
  virtual int resetPass(PassData& pd) = 0;
  // K1: orientation expansion on the device (fills the device-side orientation table).
  virtual int orient(PassData& pd) = 0;
  // Host -> device copy of one search launch's inputs (region descriptors, reference windows, sync offsets) ahead of
  // time, so that a following search(..., staged = true) starts with everything resident in HBM.
  virtual int stage(const std::vector<RegionDesc>& regs, const std::vector<uint8_t>& windows,
                    const std::vector<int32_t>& syncOff, size_t syncTotal) = 0;
  // Runs the warp-per-region search (K2 + K3 epilogue) over `regs`.  tier 0 / 1: shared-memory kernel with the fixed
  // capacities of fastShape(tier) (reports kRegionOverflow beyond them); tier kTierGeneral: large-capacity kernel with
  // per-region arenas in HBM.
  // staged=true: the launch inputs were already copied by stage(); regs/windows/syncOff are then only read for sizes.
  virtual int search(PassData& pd, const SearchConfig& sc, const std::vector<RegionDesc>& regs,
                     const std::vector<uint8_t>& windows, int tier, bool staged, std::vector<RegionResult>& out,
                     const std::vector<int32_t>& syncOff, size_t syncTotal, std::vector<WarnRec>& warns) = 0;
  // Asynchronous form for a batch whose regions are grouped by tier: regs [tierFirst[t], tierFirst[t+1]) run in tier t
  // (one launch per non-empty tier, all on the backend's stream).  searchStart returns once everything is enqueued;
  // searchFinish waits and delivers the per-region results.  One batch can be in flight at a time.
  virtual int searchStart(PassData& pd, const SearchConfig& sc, const std::vector<RegionDesc>& regs,
                          const std::vector<uint8_t>& windows, const int* tierFirst, const std::vector<int32_t>& syncOff,
                          size_t syncTotal) = 0;
  virtual int searchFinish(std::vector<RegionResult>& out, std::vector<WarnRec>& warns) = 0;
  // Copies decision / weight / orientation-id arrays of the pass and the sync-point buffer back to the host.
  virtual int fetchPass(PassData& pd, std::vector<int32_t>& syncBuf) = 0;
  virtual const char* lastError() const = 0;
  // Device-side stopwatch on the backend's stream (CUDA events): brackets a whole execute().
  virtual void timerStart() {}
  virtual double timerStopMs() { return 0.0; }
  SearchStats stats;
};

// Warp-per-region tiers.  Tiers 0 / 1: shared-memory kernel with the fixed capacities of fastShape(tier) -- the whole
// search state of a region sits in shared memory, which bounds the resident warps per SM.  Tier kTierMid: the
// large-capacity kernel (path records in per-warp HBM arenas, tables in shared memory) with a frontier capacity of
// kMidCap and several dozen resident warps per SM: past ~50 paths the search is latency bound and resident warps are
// worth more than shared-memory records.  Tier kTierGeneral: the same kernel at the full PathFinder budget
// (--max-paths), fewer warps with large arenas.
struct FastShape {
  int kv;            // variants per side (4-bit decisions in one word per side)
  int q;             // pending alleles per haplotype
  int smax;          // sync points per region
  int cap;           // frontier capacity
  int maxChildren;   // exclude + up to 6 orientations (ALLELE_GT)
  int tabBytes;      // shared memory per warp for the region's variant / orientation / allele tables
  VCE_HD int slots() const { return cap + 1 + maxChildren; }
  VCE_HD int ordCap() const { return 2 * cap; }
};
constexpr int kTierMid = 2;       // CTA kernel (vce_cta.h), small shape: frontiers up to kMidCap paths
constexpr int kTierCta = 3;       // CTA kernel, full shape: the reference's own budget (--max-paths 50 000)
constexpr int kTierGeneral = 4;   // large-capacity warp kernel: whatever the CTA kernel's fixed shapes cannot hold
constexpr int kMidCap = 2048;
inline CtaShape ctaShape(int tier) {
  // nb blocks of up to 128 entries; split blocks stay at least half full: capacity = 64 * (nb - 1) paths
  return tier == kTierMid ? CtaShape{40, 2, 6144, 512} : CtaShape{832, 3, 10240, 1024};
}
inline FastShape fastShape(int tier) {
  switch (tier) {
    case 0: return FastShape{8, 2, 8, 24, 7, 1536};      // ~9 KB per warp
    default: return FastShape{8, 8, 12, 48, 7, 2048};    // ~22 KB per warp
  }
}
struct FastLimits {
  static constexpr int kMaxVarPerSide = 8;
  static constexpr int kMaxChildren = 7;
};

}  // namespace vce
#endif
