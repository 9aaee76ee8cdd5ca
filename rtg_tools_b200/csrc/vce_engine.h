// vce_engine.h -- host side of the engine.  Replaces SequenceEvaluator.run() for a batch of sequences
// (src/com/rtg/vcf/eval/SequenceEvaluator.java:87-155):
//
//   digest   (worker pool)  variant bounds, (start,end) order check, split of every sequence into independent search
//                           regions, one self-contained PACKET per small region written into page-locked memory
//   device   (CUDA streams) per chunk of regions: H2D of the packets, thread-per-region search kernel (vce_micro.h),
//                           D2H of the fixed-size result records -- overlapped with the digest of later chunks
//   decode   (worker pool)  result records -> per-variant decisions / weights, per-region sync points
//   residual                regions the small kernel cannot take or did not settle (first / last region of a
//                           sequence, large or N-containing regions, SPILL, capacity, prefix-dependent tie-breaks) are
//                           merged / re-run with the warp-per-region kernels (vce_search.h) until every region is exact
//   assemble (worker pool)  status merge, orientation ids, sync ids, phasing counts into the caller's arrays
//
// The device work is behind `Backend` (vce_pipeline.h).  The product library links the CUDA backend only.
#ifndef VCE_ENGINE_H_
#define VCE_ENGINE_H_

#include <cstdint>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "vce_packet.h"
#include "vce_pipeline.h"
#include "vce_pool.h"

namespace vce {

struct EngineOptions {
  int gap = 4;            // split when next.start >= maxEnd + gap (any gap >= 1 is exact); a few bases
                          // keep an indel and its shifted re-representation in one region instead of a SPILL + merge
  int margin = 24;        // reference bases beyond the last variant end of a region (warp-per-region kernels)
  int microMargin = 12;   // same, thread-per-region kernel
  int wideMargin = 192;   // same, regions that start in the CTA tiers
  int denseGap = 256;     // neighbouring regions closer than this are joined up front when one of them already holds ...
  int denseCount = 8;     // ... this many variants on a side (dense clusters in repeats spill into each other otherwise)
  bool forceGeneral = false;   // tests: every region through the large-capacity kernel
  bool forceCta = false;       // tests: every warp-path region through the CTA kernel
  bool disableMicro = false;   // tests: no thread-per-region kernel
  int chunkRegions = 0;        // regions per chunk, 0 = automatic
  int splitSegment = 1 << 16;  // variants per segment of the parallel region split
  bool sortPackets = true;     // hand the packets of a chunk to the kernel grouped by shape class (lock-step lanes)
  bool deviceBuild = true;     // vce_submit: pack the packets on the device from wholesale copies of the caller's arrays
  bool residentInputs = true;  // sequences with a registered template: whole-sequence upload, windows cut out on the device
  bool deviceResults = true;   // single-pass vce_submit: result records scattered, sync ids / classes / phasing computed on the device
};

// Host side of vce_template_register: 2 bits per base (base - 1; 16 bases per word) and one bit per 16-base granule that
// holds a base other than A/C/G/T (bit g & 31 of word g >> 5, g = position >> 4).
void packTemplate(const uint8_t* bases, int32_t len, WorkerPool* pool, std::vector<uint32_t>& packed, std::vector<uint32_t>& nmask);

class Engine {
 public:
  Engine(Backend* b, WorkerPool* pool, const EngineOptions& o);
  ~Engine();
  int run(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs, vce_sequence_result* results);
  // staged variant: prepare = digest + upload of pass 0, execute = device work and settling (repeatable),
  // finish = assembly into the caller's arrays
  int prepare(const vce_config& cfg, int32_t nSeq, const vce_sequence* seqs, bool staged);
  int execute();
  int finish(vce_sequence_result* results);
  const std::string& error() const { return err_; }
  // Per-sequence summary of the last finish(): [k * 8 + i] = baseline TP, baseline FN, call TP, call FP, skipped variants
  // (both sides), mis-phasings, correct phasings, unphaseable -- the split writers' classes (SplitEvalSynchronizer.java:112-156)
  const std::vector<int64_t>& summary() const { return summary_; }
  EngineOptions& options() { return opt_; }
  Backend* backend() { return be_; }
  WorkerPool* pool() { return pool_; }
  // diagnostics of the last execute(): regions of the thread-per-region launches, variants of the MicroA class
  double microMs = 0;
  int64_t microRegions = 0, microVariants = 0;

 private:
  struct SideSeq {            // one side of one sequence in one pass
    const vce_variants* in = nullptr;
    int32_t first = 0, n = 0; // slice of the pass-wide arrays
    const int32_t* idx = nullptr;  // (start,end)-sorted position -> input index, nullptr = identity
  };
  struct RegRes {
    uint8_t state, flags;     // RS_*, bit0 usedPrefix (emptiness), bit5 usedPrefix (allele id), bit1 finished (last region of the sequence), bits 2-4 tier, bit7 results on the device
    int8_t lastId;            // alleleId() of the last baseline variant included in the region, -128 = none
    uint8_t nSync;            // thread-per-region results: sync points relative to the region start position
    union {
      uint8_t rel[8];
      struct { int32_t idx, n; } ext;   // warp-per-region results: slice of PassCtx::wideSync
    };
  };
  enum : uint8_t { RS_PENDING = 0, RS_MICRO_DONE = 1, RS_WIDE_DONE = 2, RS_DEAD = 3, RS_RESIDUAL = 4, RS_ERR_NULL = 5,
                   RS_ERR_MOVE = 6, RS_ERR_ORDER = 7, RS_EARLY = 8 };
  static constexpr uint8_t kClsEarly = 255;   // region started early in the warp-per-region kernels
  struct RegKey { int32_t seq, j; };
  struct Rare {               // state only re-run regions carry
    int extraMargin = 0;      // reference bases added after a window miss (< 0: rest of the template)
    int assumedPrefix = 0;
    bool hasPrefix = false;   // an explicit prefix was set (otherwise the default assumption)
    bool microTried = false;  // the (merged) region already went through the thread-per-region kernel
    int tier = 0;             // warp-per-region tier of the last run (kTierGeneral = large-capacity kernel)
    int lastStatus = 0;       // kRegion* of the last wide run
  };
  struct Chunk {
    int seq = 0, r0 = 0, r1 = 0;
    std::vector<int32_t> list;           // explicit region indices (settling), empty = the range [r0, r1)
    bool transient = false;              // settling: leaves the per-region packet class of the main chunks alone
    MicroJob job[2];                     // one job per packet class (MicroA, MicroB)
    std::vector<int32_t> pktRegion[2];   // packet -> region index within the sequence
    int64_t variants = 0, variantsA = 0;
  };
  struct WarnEntry { int32_t seq, j; WarnRec w; };
  struct WideBatch {           // one batch of regions in the warp-per-region kernels
    std::vector<RegKey> ids;   // grouped by tier
    std::vector<int> tier;
    int tierFirst[kTierGeneral + 2] = {0};
    PassData pd;
    AlleleTable at[2];
    std::vector<RegionDesc> descs;
    std::vector<int32_t> syncOff;
    std::vector<uint8_t> windows;
    size_t so = 0;
    double searchBefore = 0;
    bool started = false;
    bool packed = false;      // host-side image is complete (a staged job packs its early batch once)
  };

  struct PassCtx {
    int pass = 0, H = 2, kind[2] = {0, 0};
    int gtPloidy[2] = {0, 0};
    bool microOk = false;
    MicroConfig mc{};
    std::vector<SideSeq> ss[2];                  // [side][seq]; side 0 = calls, 1 = baseline
    std::vector<int32_t> vStart[2], vEnd[2];
    std::vector<uint8_t> decision[2];
    std::vector<double> weight;                  // calls
    std::vector<std::vector<int32_t>> idxStore[2];
    std::vector<std::vector<RegRange>> rr;       // per sequence
    std::vector<std::vector<RegRes>> rs;
    std::vector<std::vector<uint8_t>> rcls;      // 0 = warp-per-region path, 1 = MicroA, 2 = MicroB
    std::vector<Chunk> chunks;
    std::vector<std::pair<RegKey, int>> earlyItems;   // (region, tier) started early in the warp-per-region kernels
    WideBatch earlyBatch;                             // staged jobs: packed once, launched by every execute()
    std::vector<int32_t> wideSync;
    std::vector<WarnEntry> warns;
    std::unordered_map<uint64_t, Rare> rare;
    std::vector<std::pair<RegKey, RegRange>> undo;   // ranges changed by merges (a staged job can be executed repeatedly)
    std::vector<uint8_t> shortCircuit;
    bool launched = false;                       // chunks were started during prepare()
  };
  struct HostBlock { uint8_t* p; size_t cap, used; };

  static uint64_t key(int seq, int j) { return ((uint64_t) (uint32_t) seq << 32) | (uint32_t) j; }
  uint8_t* arenaAlloc(size_t bytes);
  void arenaReset();
  void arenaMark();
  void arenaRewind();

  int setupPass0();
  int setupPass1();
  void computeBounds(PassCtx& pc);
  void resetPass(PassCtx& pc);
  void splitSlice(const PassCtx& pc, int i, int ie, int j, int je, long maxEnd, std::vector<RegRange>& out, bool* continues,
                  std::vector<uint8_t>* near = nullptr) const;
  void splitAll(PassCtx& pc);
  void planChunks(PassCtx& pc);
  struct PacketSrc {          // what buildPacket reads, resolved once per chunk
    const RegRange* rr;
    PacketView pv;            // the caller's arrays of the chunk's sequence (vce_packet.h)
    const uint8_t* tmpl;
  };
  PacketView packetView(const PassCtx& pc, int k) const;
  template <class T> int buildPacket(const PassCtx& pc, const PacketSrc& ps, int j, uint8_t* out) const;
  int buildChunk(PassCtx& pc, Chunk& ch, int worker, bool launch);
  int buildChunkDevice(PassCtx& pc, Chunk& ch, int worker);
  // resident inputs (registered template + whole-sequence upload): see residentSetup
  struct ResidentSeq {
    bool on = false;               // whole-sequence upload + registered template: packets are built from `dv`
    bool have = false;             // at least the arrays of the final passes are on the device (fVStart / fPhased / fStatus)
    const int32_t* fVStart[2] = {nullptr, nullptr};   // device pointers, sequence-local index
    const uint8_t* fPhased[2] = {nullptr, nullptr};
    const uint8_t* fStatus[2] = {nullptr, nullptr};
    PacketView dv;                 // device pointers to the sequence's arrays
    const uint32_t* t2 = nullptr;  // device copy of the template
    const uint32_t* nmask = nullptr;
    int events[2] = {-1, -1};
  };
  std::vector<ResidentSeq> resident_;
  int residentSetup(PassCtx& pc);
  int buildChunkResident(PassCtx& pc, Chunk& ch, int worker);
  // device-resident results (vce_final.h)
  bool devRes_ = false;
  std::vector<int64_t> summary_;
  void summaryFromResults(const vce_sequence_result* results);
  std::vector<int32_t> regBase_;   // [nSeq + 1] global region index of every sequence's first region
  bool wantDeviceResults(const PassCtx& pc) const;
  int resultsSetup(PassCtx& pc);
  int resultsCollect(PassCtx& pc);
  int finishDevice(vce_sequence_result* results);
  void configurePass(PassCtx& pc);
  int runChunks(PassCtx& pc);
  void decodeChunk(PassCtx& pc, Chunk& ch, int64_t& iterations);
  void assembleRegionEarly(PassCtx& pc, int k, int j);
  template <class T> void decodeJob(PassCtx& pc, int k, const MicroJob& job, const std::vector<int32_t>& pktRegion, int64_t& iterations);
  int settle(PassCtx& pc, WideBatch& early);
  int wideStart(PassCtx& pc, const std::vector<std::pair<RegKey, int>>& items, WideBatch& wb);
  int widePack(PassCtx& pc, const std::vector<std::pair<RegKey, int>>& items, WideBatch& wb);
  int wideLaunch(WideBatch& wb);
  int wideFinish(PassCtx& pc, WideBatch& wb, std::vector<int>& statusOut);
  int startEarly(PassCtx& pc, WideBatch& early);
  int startTier(const PassCtx& pc, const RegRange& g) const;
  int regionStartPos(const PassCtx& pc, int k, int j) const;
  bool assembleHeader(int k, vce_sequence_result& res);
  void assembleBlock(int k, int side, int lo, int hi, vce_sequence_result& res) const;
  void phasing(int k, const std::vector<int32_t>& sync, int32_t out[3]) const;
  void resetPassRun(PassCtx& pc);

  Backend* be_;
  WorkerPool* pool_;
  EngineOptions opt_;
  vce_config cfg_{};
  int32_t nSeq_ = 0;
  const vce_sequence* seqs_ = nullptr;
  PassCtx pass_[2];
  vce_sequence_result* liveResults_ = nullptr;   // vce_submit, single pass: outputs are written as records are decoded
  bool earlyUsed_ = false;
  std::vector<uint8_t> earlyDone_[2];            // [side][variant of pass 0] outputs already written
  bool executed_ = false;
  bool staged_ = false;
  std::string err_;
  std::mutex errMu_;
  std::mutex arenaMu_;
  std::vector<HostBlock> arena_;
  std::vector<size_t> arenaMarks_;
  struct alignas(64) SyncPart { int pass, seq, r0, r1; int err; std::vector<int32_t> v; };
  std::vector<SyncPart> syncParts_;               // sync points per block of regions (assembly)
  struct alignas(64) SplitOut { std::vector<RegRange> v; std::vector<uint8_t> near; };   // one cache line each: neighbours are filled concurrently
  std::vector<SplitOut> splitOut_;                // regions per segment of the parallel split
  std::vector<int> syncPartFirst_[2];             // [pass][seq] first part of the sequence
  std::vector<std::vector<int32_t>> seqSync_[2];   // [pass][seq] sync points (assembly)
  std::vector<std::vector<uint8_t>> seqDec1_[2];   // [side][seq] pass-2 decision by input index
  std::vector<std::vector<double>> seqW1_;         // [seq] pass-2 call weight by input index
  std::vector<std::vector<uint32_t>> classCount_, classKeys_;   // per worker: counting sort of packets by shape class
  std::vector<std::vector<int32_t>> classOrder_;
  std::vector<uint8_t> (*scratch_)[2] = nullptr;     // [worker][class] packet staging
  std::vector<uint32_t> (*scratchOff_)[2] = nullptr;
};

}  // namespace vce
#endif
