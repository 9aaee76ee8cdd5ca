// vce_roc.cu -- K4: RocContainer accumulate + sort + cumulative sums on the device.
//
// Reference behaviour (src/com/rtg/vcf/eval/RocContainer.java):
//   addRocLine :225-258  per filter, a TreeMap<score, RocPoint>; equal scores are summed with sequential
//                        double += in ARRIVAL order (RocPoint.add, RocPoint.java:70-74); NaN / Inf scores go
//                        to one NaN bucket that sorts last (:232-237, :447-457)
//   writeRocs  :296-309  buckets visited in comparator order, cumulative = sequential double += per bucket
// Floating-point addition is not associative, so both levels are reproduced as the same sequential chains:
//   1. scores -> order-preserving 64-bit keys, one stable radix sort of (key, arrival index) shared by all filters
//   2. bucket sums: one thread per (bucket, filter) adds its members in arrival order
//   3. cumulative sums: one warp per filter walks the buckets in key order; lanes fetch 32 bucket sums with one
//      coalesced load, the 3 dependent add chains (tp, fp, tpRaw) run through warp shuffles, lane j keeps the
//      value after step j, rows of non-empty buckets are written compacted.
// The radix sort and the prefix sums are CUB (plumbing); the ROC-specific kernels are below.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vce.h"
#include "vce_cuda.h"
#include "vce_pool.h"

namespace vce {

__device__ __forceinline__ unsigned long long rocKey(double s, int ascending) {
  if (isnan(s) || isinf(s)) return ~0ull;  // the "None" bucket is last in both orders
  unsigned long long b = (unsigned long long) __double_as_longlong(s);
  b = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);  // ascending total order, -0.0 < +0.0 (Double.compareTo)
  if (!ascending) b = ~b;
  if (b == ~0ull) b = ~0ull - 1;  // keep the NaN key unique (only reachable by -NaN patterns, already excluded)
  return b;
}

__global__ void k_roc_keys(long long n, const double* __restrict__ score, int ascending, unsigned long long* __restrict__ key,
                           unsigned int* __restrict__ idx, unsigned long long* __restrict__ noScore) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  if (i < n) {
    const double s = score[i];
    bad = (isnan(s) || isinf(s)) ? 1 : 0;
    key[i] = rocKey(s, ascending);
    idx[i] = (unsigned int) i;
  }
  const unsigned m = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(noScore, (unsigned long long) __popc(m));
}

__global__ void k_roc_heads(long long n, const unsigned long long* __restrict__ key, int* __restrict__ head) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// bucketStart[b] = first sorted position of bucket b; bucketOf = inclusive scan of head - 1
__global__ void k_roc_starts(long long n, const int* __restrict__ head, const int* __restrict__ scan, int* __restrict__ bucketStart,
                             int nBuckets) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && head[i]) bucketStart[scan[i] - 1] = (int) i;
  if (i == 0) bucketStart[nBuckets] = (int) n;
}

// One thread per (bucket, filter): sequential sums in arrival order.
__global__ void k_roc_bucket_sums(int nBuckets, int nFilters, const int* __restrict__ bucketStart, const unsigned int* __restrict__ idx,
                                  const double* __restrict__ tp, const double* __restrict__ fp, const double* __restrict__ raw,
                                  const unsigned int* __restrict__ mask, double* __restrict__ sTp, double* __restrict__ sFp,
                                  double* __restrict__ sRaw, int* __restrict__ nonEmpty) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const bool have = b < nBuckets;
  const int s = have ? bucketStart[b] : 0, e = have ? bucketStart[b + 1] : 0;
  double a = 0, c = 0, d = 0;
  int cnt = 0;
  // small buckets: the owning lane walks its members; LARGE buckets (few distinct scores, e.g. QUAL at one decimal) are
  // walked by the whole warp: 32 members are fetched with one coalesced index load + one gather, the additions still
  // happen in arrival order (the dependent chain runs through shuffles), so the FP64 sums keep the reference's bits
  const bool big = e - s > 64;
  if (!big) {
    for (int j = s; j < e; ++j) {
      const unsigned int i = idx[j];
      const unsigned int m = mask ? mask[i] : 0xffffffffu;
      if ((m >> f) & 1u) {
        if (cnt == 0) { a = tp[i]; c = fp[i]; d = raw[i]; }  // new RocPoint<>(point), RocContainer.java:256
        else { a += tp[i]; c += fp[i]; d += raw[i]; }          // RocPoint.add
        ++cnt;
      }
    }
  }
  unsigned todo = __ballot_sync(0xffffffffu, big);
  while (todo) {
    const int owner = __ffs((int) todo) - 1;
    todo &= todo - 1;
    const int os = __shfl_sync(0xffffffffu, s, owner), oe = __shfl_sync(0xffffffffu, e, owner);
    double wa = 0, wc = 0, wd = 0;
    int wcnt = 0;
    for (int base = os; base < oe; base += 32) {
      const int j = base + lane;
      double vt = 0, vf = 0, vr = 0;
      int in = 0;
      if (j < oe) {
        const unsigned int i = idx[j];
        const unsigned int m = mask ? mask[i] : 0xffffffffu;
        if ((m >> f) & 1u) { in = 1; vt = tp[i]; vf = fp[i]; vr = raw[i]; }
      }
      const unsigned mm = __ballot_sync(0xffffffffu, in);
#pragma unroll 4
      for (int q = 0; q < 32; ++q) {
        const double xt = __shfl_sync(0xffffffffu, vt, q);
        const double xf = __shfl_sync(0xffffffffu, vf, q);
        const double xr = __shfl_sync(0xffffffffu, vr, q);
        if ((mm >> q) & 1u) {
          if (wcnt == 0) { wa = xt; wc = xf; wd = xr; }
          else { wa += xt; wc += xf; wd += xr; }
          ++wcnt;
        }
      }
    }
    if (lane == owner) { a = wa; c = wc; d = wd; cnt = wcnt; }
  }
  if (!have) return;
  const size_t o = (size_t) f * nBuckets + b;
  sTp[o] = a; sFp[o] = c; sRaw[o] = d;
  nonEmpty[o] = cnt > 0 ? 1 : 0;
}

// One CTA per filter: cumulative sums over the buckets in key order, rows compacted to non-empty buckets.  The sums are a
// left fold in bucket order (RocContainer.java:297-311 adds the points of one threshold after the other, FP64), so the
// additions stay a dependent chain -- but nothing else does: a tile of bucket sums is staged into shared memory by the
// whole block, three threads (one per quantity, in three different warps) fold it in place, and the row numbers
// (prefix count of non-empty buckets), the threshold gather and the stores are done by all threads again.
constexpr int kChainTile = 1024;
constexpr int kChainThreads = 256;
__global__ void __launch_bounds__(kChainThreads) k_roc_chain(int nBuckets, const int* __restrict__ bucketStart, const unsigned int* __restrict__ idx,
                            const double* __restrict__ score, const double* __restrict__ sTp, const double* __restrict__ sFp,
                            const double* __restrict__ sRaw, const int* __restrict__ nonEmpty, long long cap,
                            double* __restrict__ oThr, double* __restrict__ oTp, double* __restrict__ oFp, double* __restrict__ oRaw,
                            long long* __restrict__ nRows) {
  __shared__ double sv[3][kChainTile];
  __shared__ int sne[kChainTile];
  __shared__ int warpCnt[kChainThreads / 32];
  __shared__ double carry[3];
  const int f = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t fo = (size_t) f * nBuckets;
  if (tid < 3) carry[tid] = 0.0;   // new RocPoint<>() : zeros, RocContainer.java:297
  long long row = 0;               // rows written so far (the same value in every thread)
  __syncthreads();
  for (int base = 0; base < nBuckets; base += kChainTile) {
    const int cnt = min(kChainTile, nBuckets - base);
    for (int j = tid; j < cnt; j += kChainThreads) {
      const size_t o = fo + base + j;
      sne[j] = nonEmpty[o];
      sv[0][j] = sTp[o]; sv[1][j] = sFp[o]; sv[2][j] = sRaw[o];
    }
    __syncthreads();
    if (lane == 0 && warp < 3) {
      double* v = sv[warp];
      double c = carry[warp];
      int j = 0;
      // an empty bucket holds +0.0 (k_roc_bucket_sums) and the running sum starts at +0.0, so it is never -0.0: adding
      // the zero leaves every bit of it alone, and the chain is one DADD per bucket with nothing else on it
      for (; j + 8 <= cnt; j += 8) {
        double x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = v[j + k];
#pragma unroll
        for (int k = 0; k < 8; ++k) { c += x[k]; x[k] = c; }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[j + k] = x[k];
      }
      for (; j < cnt; ++j) { c += v[j]; v[j] = c; }
      carry[warp] = c;
    }
    __syncthreads();
    for (int j0 = 0; j0 < cnt; j0 += kChainThreads) {
      const int j = j0 + tid;
      const int ne = j < cnt ? sne[j] : 0;
      const unsigned m = __ballot_sync(0xffffffffu, ne);
      if (lane == 0) warpCnt[warp] = __popc(m);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kChainThreads / 32; ++w) {
        const int c = warpCnt[w];
        if (w < warp) before += c;
        total += c;
      }
      if (ne) {
        const long long r = row + before + __popc(m & ((1u << lane) - 1u));
        if (r < cap) {
          const size_t o = (size_t) f * cap + r;
          double s = score[idx[bucketStart[base + j]]];
          if (isnan(s) || isinf(s)) s = nan("");
          oThr[o] = s; oTp[o] = sv[0][j]; oFp[o] = sv[1][j]; oRaw[o] = sv[2][j];
        }
      }
      row += total;
      __syncthreads();   // warpCnt is rewritten by the next round, the tile by the next pass
    }
  }
  if (tid == 0) nRows[f] = row;
}

#define ROC_TRY(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char b_[512];                                                                           \
      snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      err = b_;                                                                               \
      cleanup();                                                                              \
      return VCE_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

// Per-device K4 context, kept across calls: stream, events, ONE device slab and ONE page-locked staging slab (both grow
// only), so that a steady-state vce_roc allocates nothing.  One call at a time per device (callers queue on the mutex).
namespace {
struct RocCtx {
  int device = -1;
  std::mutex mu;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  uint8_t* dSlab = nullptr; size_t dCap = 0, dUsed = 0;
  uint8_t* hSlab = nullptr; size_t hCap = 0;
  cudaError_t init(int dev) {
    device = dev;
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaEventCreate(&e2);
    return e;
  }
  cudaError_t reserveDevice(size_t bytes) {
    dUsed = 0;
    if (bytes <= dCap) return cudaSuccess;
    if (dSlab) cudaFree(dSlab);
    dSlab = nullptr; dCap = 0;
    const size_t want = bytes + bytes / 8 + (1 << 20);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dSlab), want);
    if (e == cudaSuccess) dCap = want;
    return e;
  }
  cudaError_t reserveHost(size_t bytes) {
    if (bytes <= hCap) return cudaSuccess;
    if (hSlab) cudaFreeHost(hSlab);
    hSlab = nullptr; hCap = 0;
    const size_t want = bytes + bytes / 8 + (1 << 20);
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&hSlab), want, cudaHostAllocDefault);
    if (e == cudaSuccess) hCap = want;
    return e;
  }
  void* take(size_t bytes) {   // bump allocation from the device slab (256-byte aligned)
    bytes = (bytes + 255) & ~(size_t) 255;
    if (dUsed + bytes > dCap) return nullptr;
    void* p = dSlab + dUsed;
    dUsed += bytes;
    return p;
  }
};
std::mutex g_rocMu;
std::vector<RocCtx*> g_rocCtx;
RocCtx* rocCtxFor(int device, cudaError_t* e) {
  std::lock_guard<std::mutex> lk(g_rocMu);
  for (RocCtx* c : g_rocCtx) if (c->device == device) return c;
  RocCtx* c = new RocCtx();
  *e = c->init(device);
  if (*e != cudaSuccess) { delete c; return nullptr; }
  g_rocCtx.push_back(c);
  return c;
}
}  // namespace

void rocRelease() {
  std::lock_guard<std::mutex> lk(g_rocMu);
  for (RocCtx* c : g_rocCtx) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    if (c->dSlab) cudaFree(c->dSlab);
    if (c->hSlab) cudaFreeHost(c->hSlab);
    cudaEventDestroy(c->e0); cudaEventDestroy(c->e1); cudaEventDestroy(c->e2);
    cudaStreamDestroy(c->st);
    delete c;
  }
  g_rocCtx.clear();
}

int rocRun(int device, const vce_roc_in* in, vce_roc_out* out, std::string& err, RocTiming* timing, WorkerPool* pool) {
  const long long n = in->n;
  const int nF = in->n_filters;
  if (n < 0 || nF < 1 || nF > 32 || n > 0x7fffffffLL - 64) { err = "bad ROC arguments"; return VCE_ERR_BAD_ARGUMENT; }
  for (int f = 0; f < nF; ++f) out->n_rows[f] = 0;
  *out->no_score = 0;
  if (n == 0) return VCE_OK;
  auto cleanup = [&]() {};
  if (cudaSetDevice(device) != cudaSuccess) { err = "cudaSetDevice failed"; return VCE_ERR_CUDA; }
  cudaError_t ce = cudaSuccess;
  RocCtx* ctx = rocCtxFor(device, &ce);
  if (!ctx) { err = std::string("ROC context: ") + cudaGetErrorString(ce); return VCE_ERR_CUDA; }
  std::lock_guard<std::mutex> ctxLock(ctx->mu);
  cudaStream_t st = ctx->st;
  cudaEvent_t e0 = ctx->e0, e1 = ctx->e1, e2 = ctx->e2;
  // everything the call needs on the device, up front (the bucket-level arrays are sized by n: no allocation after the sort)
  {
    size_t tmpA = 0, tmpB = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmpA, (unsigned long long*) nullptr, (unsigned long long*) nullptr, (unsigned int*) nullptr,
                                    (unsigned int*) nullptr, (int) n, 0, 64, st);
    cub::DeviceScan::InclusiveSum(nullptr, tmpB, (int*) nullptr, (int*) nullptr, (int) n, st);
    const size_t un = (size_t) n;
    size_t need = 4 * (un * 8 + 256) + (un * 4 + 256) + 2 * (un * 8 + 256) + 512 + 2 * (un * 4 + 256) + 3 * (un * 4 + 1024) + std::max(tmpA, tmpB) + 1024;
    // bucket-level arrays (sums, cumulative rows per filter): at most one bucket per call record
    const size_t fbMax = (size_t) nF * un;
    need += 7 * (fbMax * 8 + 256) + (fbMax * 4 + 256) + ((size_t) nF * 8 + 256) + 1024;
    ROC_TRY(ctx->reserveDevice(need));
    ROC_TRY(ctx->reserveHost(un * 36 + 4096));
  }
  auto alloc = [&](size_t bytes, void** p) -> cudaError_t {
    *p = ctx->take(bytes ? bytes : 16);
    if (*p) return cudaSuccess;
    // the bucket-level arrays come after the sort, when their size is known: a separate, also persistent, allocation would
    // do; growing the slab is simpler and happens once
    return cudaErrorMemoryAllocation;
  };
  double *dScore, *dTp, *dFp, *dRaw;
  unsigned int* dMask = nullptr;
  unsigned long long *dKey, *dKey2, *dNoScore;
  unsigned int *dIdx, *dIdx2;
  int *dHead, *dScan, *dStart;
  ROC_TRY(cudaEventRecord(e0, st));
  ROC_TRY(alloc(n * 8, (void**) &dScore)); ROC_TRY(alloc(n * 8, (void**) &dTp)); ROC_TRY(alloc(n * 8, (void**) &dFp));
  ROC_TRY(alloc(n * 8, (void**) &dRaw));
  if (in->filter_mask) ROC_TRY(alloc(n * 4, (void**) &dMask));
  ROC_TRY(alloc(n * 8, (void**) &dKey)); ROC_TRY(alloc(n * 8, (void**) &dKey2)); ROC_TRY(alloc(8, (void**) &dNoScore));
  ROC_TRY(alloc(n * 4, (void**) &dIdx)); ROC_TRY(alloc(n * 4, (void**) &dIdx2));
  ROC_TRY(alloc(n * 4, (void**) &dHead)); ROC_TRY(alloc(n * 4, (void**) &dScan)); ROC_TRY(alloc((n + 1) * 4, (void**) &dStart));
  {
    // the caller's arrays are pageable: they go through the page-locked slab in parallel pieces (a direct copy from pageable
    // memory runs at a fraction of the PCIe rate), every piece is uploaded as soon as it is staged
    struct Piece { const uint8_t* src; uint8_t* stage; uint8_t* dst; size_t bytes; };
    std::vector<Piece> pieces;
    const size_t chunk = (size_t) 4 << 20;
    size_t hoff = 0;
    auto add = [&](const void* src, void* dst, size_t bytes) {
      for (size_t o = 0; o < bytes; o += chunk) {
        const size_t b = std::min(chunk, bytes - o);
        pieces.push_back(Piece{static_cast<const uint8_t*>(src) + o, ctx->hSlab + hoff, static_cast<uint8_t*>(dst) + o, b});
        hoff += b;
      }
    };
    add(in->score, dScore, (size_t) n * 8);
    add(in->tp, dTp, (size_t) n * 8);
    add(in->fp, dFp, (size_t) n * 8);
    add(in->tp_raw, dRaw, (size_t) n * 8);
    if (dMask) add(in->filter_mask, dMask, (size_t) n * 4);
    std::atomic<int> bad(0);
    std::mutex issue;
    auto work = [&](int t, int) {
      const Piece& pc = pieces[(size_t) t];
      std::memcpy(pc.stage, pc.src, pc.bytes);
      std::lock_guard<std::mutex> lk(issue);   // (one stream: the copies are issued from whichever thread staged the piece)
      cudaSetDevice(device);
      if (cudaMemcpyAsync(pc.dst, pc.stage, pc.bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) bad.store(1);
    };
    if (pool) pool->parallelFor((int) pieces.size(), work);
    else for (int t = 0; t < (int) pieces.size(); ++t) work(t, 0);
    if (bad.load()) { err = "ROC upload failed"; return VCE_ERR_CUDA; }
  }
  ROC_TRY(cudaMemsetAsync(dNoScore, 0, 8, st));
  ROC_TRY(cudaEventRecord(e1, st));
  const int tb = 256;
  const int nb = (int) ((n + tb - 1) / tb);
  k_roc_keys<<<nb, tb, 0, st>>>(n, dScore, in->ascending, dKey, dIdx, dNoScore);
  size_t tmpBytes = 0, tmp2 = 0;
  ROC_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, dKey, dKey2, dIdx, dIdx2, (int) n, 0, 64, st));
  ROC_TRY(cub::DeviceScan::InclusiveSum(nullptr, tmp2, dHead, dScan, (int) n, st));
  void* dTmp;
  ROC_TRY(alloc(std::max(tmpBytes, tmp2) + 64, &dTmp));
  ROC_TRY(cub::DeviceRadixSort::SortPairs(dTmp, tmpBytes, dKey, dKey2, dIdx, dIdx2, (int) n, 0, 64, st));
  k_roc_heads<<<nb, tb, 0, st>>>(n, dKey2, dHead);
  ROC_TRY(cub::DeviceScan::InclusiveSum(dTmp, tmp2, dHead, dScan, (int) n, st));
  int nBuckets = 0;
  ROC_TRY(cudaMemcpyAsync(&nBuckets, dScan + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  ROC_TRY(cudaStreamSynchronize(st));
  k_roc_starts<<<nb, tb, 0, st>>>(n, dHead, dScan, dStart, nBuckets);
  double *sTp, *sFp, *sRaw, *oThr, *oTp, *oFp, *oRaw;
  int* dNe;
  long long* dRows;
  const size_t fb = (size_t) nF * nBuckets;
  const long long cap = nBuckets;
  ROC_TRY(alloc(fb * 8, (void**) &sTp)); ROC_TRY(alloc(fb * 8, (void**) &sFp)); ROC_TRY(alloc(fb * 8, (void**) &sRaw));
  ROC_TRY(alloc(fb * 4, (void**) &dNe)); ROC_TRY(alloc(nF * 8, (void**) &dRows));
  ROC_TRY(alloc(fb * 8, (void**) &oThr)); ROC_TRY(alloc(fb * 8, (void**) &oTp)); ROC_TRY(alloc(fb * 8, (void**) &oFp));
  ROC_TRY(alloc(fb * 8, (void**) &oRaw));
  dim3 g((nBuckets + tb - 1) / tb, nF);
  k_roc_bucket_sums<<<g, tb, 0, st>>>(nBuckets, nF, dStart, dIdx2, dTp, dFp, dRaw, dMask, sTp, sFp, sRaw, dNe);
  k_roc_chain<<<nF, kChainThreads, 0, st>>>(nBuckets, dStart, dIdx2, dScore, sTp, sFp, sRaw, dNe, cap, oThr, oTp, oFp, oRaw, dRows);
  ROC_TRY(cudaGetLastError());
  ROC_TRY(cudaEventRecord(e2, st));
  std::vector<long long> rows(nF);
  unsigned long long noScore = 0;
  ROC_TRY(cudaMemcpyAsync(rows.data(), dRows, nF * 8, cudaMemcpyDeviceToHost, st));
  ROC_TRY(cudaMemcpyAsync(&noScore, dNoScore, 8, cudaMemcpyDeviceToHost, st));
  ROC_TRY(cudaStreamSynchronize(st));
  int rc = VCE_OK;
  for (int f = 0; f < nF; ++f) {
    out->n_rows[f] = rows[f];
    long long k = rows[f];
    if (k > out->cap) { k = out->cap; rc = VCE_ERR_CAPACITY; err = "ROC output capacity too small"; }
    if (k > 0) {
      ROC_TRY(cudaMemcpyAsync(out->threshold + (size_t) f * out->cap, oThr + (size_t) f * cap, k * 8, cudaMemcpyDeviceToHost, st));
      ROC_TRY(cudaMemcpyAsync(out->cum_tp + (size_t) f * out->cap, oTp + (size_t) f * cap, k * 8, cudaMemcpyDeviceToHost, st));
      ROC_TRY(cudaMemcpyAsync(out->cum_fp + (size_t) f * out->cap, oFp + (size_t) f * cap, k * 8, cudaMemcpyDeviceToHost, st));
      ROC_TRY(cudaMemcpyAsync(out->cum_tp_raw + (size_t) f * out->cap, oRaw + (size_t) f * cap, k * 8, cudaMemcpyDeviceToHost, st));
    }
  }
  ROC_TRY(cudaStreamSynchronize(st));
  *out->no_score = (int64_t) noScore;
  if (timing) {
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, e0, e1);
    cudaEventElapsedTime(&b, e1, e2);
    timing->h2dMs = a;
    timing->kernelMs = b;
    timing->launches = 9;
  }
  return rc;
}

}  // namespace vce
