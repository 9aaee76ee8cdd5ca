// vce_pipeline.h -- host side of the engine: packs a batch of sequences into flat arrays, splits every
// sequence into independent search regions, drives the search kernels (re-running the few regions that
// report SPILL / OVERFLOW / prefix-dependent tie-breaks) and assembles the per-variant results the
// reference's SequenceEvaluator.run() would hand to EvalSynchronizer.write
// (src/com/rtg/vcf/eval/SequenceEvaluator.java:87-155).
//
// The device work is behind `Backend`.  The product library links exactly one backend, the CUDA one
// (vce_cuda.cu).  tests/emul builds a 1-lane host instantiation of the same search source to exercise this
// host logic where no GPU exists; it is never linked into the product.
#ifndef VCE_PIPELINE_H_
#define VCE_PIPELINE_H_

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/vce.h"
#include "vce_device.h"
#include "vce_search.h"

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

class Backend {
 public:
  virtual ~Backend() {}
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
  // Runs the region search (K2 + K3 epilogue) over `regs`.  general=false: shared-memory kernel with fixed small
  // capacities (reports kRegionOverflow beyond them); general=true: large-capacity kernel with per-region arenas.
  // staged=true: the launch inputs were already copied by stage(); regs/windows/syncOff are then only read for sizes.
  virtual int search(PassData& pd, const SearchConfig& sc, const std::vector<RegionDesc>& regs,
                     const std::vector<uint8_t>& windows, bool general, bool staged, std::vector<RegionResult>& out,
                     const std::vector<int32_t>& syncOff, size_t syncTotal, std::vector<WarnRec>& warns) = 0;
  // Copies decision / weight / orientation-id arrays of the pass and the sync-point buffer back to the host.
  virtual int fetchPass(PassData& pd, std::vector<int32_t>& syncBuf) = 0;
  virtual const char* lastError() const = 0;
  // Device-side stopwatch on the backend's stream (CUDA events): brackets a whole execute().
  virtual void timerStart() {}
  virtual double timerStopMs() { return 0.0; }
  SearchStats stats;
};

// Fixed capacities of the shared-memory ("fast") search kernel.
struct FastLimits {
  static constexpr int kMaxVarPerSide = 8;   // 4-bit decisions in one word per side
  static constexpr int kQ = 2;               // pending alleles per haplotype
  static constexpr int kSMax = 8;            // sync points per region
  static constexpr int kCap = 24;            // frontier capacity
  static constexpr int kMaxChildren = 7;     // exclude + up to 6 orientations (ALLELE_GT)
  static constexpr int kSlots = kCap + 1 + kMaxChildren;
  static constexpr int kOrdCap = 2 * kCap;
};

struct PipelineOptions {
  int gap = 1;        // split when next.start >= maxEnd + gap  (gap >= 1)
  int margin = 24;    // reference bases kept beyond the last variant end of a region
  bool wantSyncId = true;
};

class Pipeline {
 public:
  Pipeline(Backend* b, const PipelineOptions& o) : be_(b), opt_(o) {}
  int run(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs, vce_sequence_result* results);
  // staged variant (bench): prepare = pack + split + upload, execute = kernels + re-run loops, finish = assembly
  int prepare(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs);
  int execute();
  int finish(vce_sequence_result* results);
  const std::string& error() const { return err_; }
  Backend* backend() { return be_; }

 private:
  struct Region {
    RegionDesc d;
    RegionResult r;
    int syncOff = 0;
    bool alive = true;
    bool general = false;
    int assumedPrefix = 0;
    int extraMargin = 0;   // reference bases added to the window after a window miss (< 0: rest of the template)
  };
  struct Launch {  // one search launch, ready to go
    std::vector<int> ids;
    std::vector<RegionDesc> descs;
    std::vector<uint8_t> windows;
    std::vector<int32_t> syncOff;
  };
  struct PassState {
    PassData pd;
    bool fetched = false;
    std::vector<Region> regions;          // sequence order
    std::vector<Region> initRegions;      // as split, before any merge (a staged job can be executed repeatedly)
    size_t initSyncSize = 0;
    Launch mainFast;                      // staged in prepare() for pass 0
    std::vector<int> mainGen;
    std::vector<int32_t> seqRegionFirst;  // [nSeq+1]
    std::vector<int32_t> syncBuf;
    std::vector<WarnRec> warns;
    std::vector<uint8_t> shortCircuit;    // per sequence
  };
  int buildPass0();
  int buildPass1();
  void splitRegions(PassState& ps);
  bool fastEligible(const PassState& ps, const RegionDesc& d) const;
  int setupPass(PassState& ps, bool stageMain);
  int runPass(PassState& ps, bool staged);
  void buildLaunch(PassState& ps, const std::vector<int>& ids, bool general, Launch& L);
  int runLaunch(PassState& ps, Launch& L, bool general, bool staged);
  int launch(PassState& ps, const std::vector<int>& ids, bool general);
  static int orientBound(int kind, int H, int gtPloidy, int maxSlots);
  void assemble(int k, vce_sequence_result& res);
  void phasing(int k, int32_t out[3]) const;

  Backend* be_;
  PipelineOptions opt_;
  vce_config cfg_{};
  int32_t nSeq_ = 0;
  const vce_sequence* seqs_ = nullptr;
  AlleleTable alleles_[2];
  std::vector<int32_t> slotBase_[2];  // per sequence
  PassState pass_[2];
  bool executed_ = false;
  std::string err_;
};

}  // namespace vce
#endif
