// vce_device.h -- POD layouts shared by the host pipeline and the CUDA kernels.
//
// Everything the search kernels touch lives in flat int32/uint8 arrays ("structure of arrays")
// so that a warp reads neighbouring variants / alleles with coalesced loads.
#ifndef VCE_DEVICE_H_
#define VCE_DEVICE_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define VCE_HD __host__ __device__ __forceinline__
// out-of-line helpers: ONE copy in the instruction stream (the search kernels are bound by instruction fetch); they are
// static and take what they need as arguments so that the caller's state can stay in registers
#define VCE_HDN __host__ __device__ __noinline__
#else
#define VCE_HD inline
#define VCE_HDN inline
#endif

namespace vce {

constexpr int32_t kNoNext = 0x7fffffff;      // "no further variant on this side"
constexpr int32_t kPrefixEmpty = -0x7fffffff; // baseline included list is empty before this region

// Region outcome codes written by the search kernel.
enum : int32_t {
  kRegionPending = 0,
  kRegionClean = 1,      // ended with exactly one in-sync path at the boundary: results valid
  kRegionFinished = 2,   // last region of a sequence, best path selected: results valid
  kRegionSpill = 3,      // not independent of the next region: merge and re-run
  kRegionOverflow = 4,   // exceeded this kernel's capacity: re-run in the large-capacity kernel
  kRegionErrNull = 5,    // best path null (SequenceEvaluator.java:105)
  kRegionErrMove = 6,    // moveForward while in a variant (HaplotypePlayback.java:260)
  kRegionErrOrder = 7,   // out of order alleles (HaplotypePlayback.java:208)
  kRegionWinMiss = 8     // a haplotype ran off the staged reference window: re-run with a wider window
};

// One independent search region (a run of variants on both sides separated from its neighbours by a gap).
struct RegionDesc {
  int32_t seq;            // sequence index within the batch
  int32_t tmplLen;        // template length of that sequence
  int32_t cFirst, cCount; // call variants [cFirst, cFirst+cCount) in the pass's sorted call arrays
  int32_t bFirst, bCount; // same for baseline
  int32_t cNextStart;     // start of the first call variant after the region, kNoNext if none
  int32_t bNextStart;
  int32_t startPos;       // position of the fresh in-sync path; -1 = true start of the sequence
  int32_t winStart;       // reference window covers template positions [winStart, winStart+winLen)
  int32_t winLen;
  uint32_t winOff;        // byte offset of the window in the packed window buffer
  int32_t prefixAlleleId; // OrientedVariant.alleleId() of the last baseline variant included before the
                          // region, kPrefixEmpty if none (MaxSumBoth.java:55,76)
  int32_t layoutW;        // words per path state (general kernel) -- see PathLayout
  int32_t qMax;           // pending-allele queue capacity per haplotype
  int32_t decBits;        // bits per include/exclude decision (4 or 8)
  int32_t cSlot0, cSlotN; // allele slots [cSlot0, cSlot0 + cSlotN) of the call side hold every allele of the region's calls
  int32_t bSlot0, bSlotN; // same for baseline (the kernels stage these ranges into shared memory)
};

// Per-region result header.
struct RegionResult {
  int32_t status;         // kRegion*
  int32_t usedPrefix;     // bit0: a tie-break depended on whether a baseline variant was included before the region, bit1: on its allele id
  int32_t nSync;          // number of sync points written for this region
  int32_t nWarn;          // too-complex events in this region
  int32_t lastBaseAlleleId; // alleleId() of the last baseline variant included in this region, kPrefixEmpty if none
  int32_t maxFrontier;    // diagnostics
  int32_t iterations;     // total best-first iterations in this region
  int32_t pad;
};

struct WarnRec { int32_t region, paths, iterations, start1, end1; };

// One side (calls or baseline) of one pass, all sequences of the batch concatenated, sorted by (start,end).
struct SideArrays {
  const int32_t* vStart;  // [nVar]
  const int32_t* vEnd;    // [nVar]
  const int32_t* oOff;    // [nVar+1] orientation rows of variant v are [oOff[v], oOff[v+1])
  const int32_t* oSlot;   // [nRows*H] allele slot per haplotype (index into a*), -1 = null allele
  const int8_t*  oIds;    // [nRows*4] allele ids (for output)
  const uint8_t* oOrig;   // [nRows]   isOriginal
  const int32_t* aStart;  // [nSlots]
  const int32_t* aEnd;    // [nSlots]
  const int32_t* aNtOff;  // [nSlots+1]
  const uint8_t* nt;      // packed allele bases
  const uint8_t* vFlags;  // [nVar] bit0 phased, bit4 outside-eval
  // outputs (indexed by sorted variant index)
  uint8_t* outDecision;   // [nVar] 0 = untouched (skipped), 1 = excluded, 2+k = included with orientation k
  double*  outWeight;     // [nVar]
};

}  // namespace vce
#endif
