// vce_vcf.cu -- CUDA instantiation of the loader passes (vce_vcf.h, vce_bgzf.h) and their entry points.
//   k_vcf_for<F>   : one thread per byte / line / run of equal starts / kept variant (the pass bodies of vce_vcf.h)
//   k_bgzf_inflate : one BGZF block per warp (DEFLATE, vce_bgzf.h), k_bgzf_crc: CRC32 of a block's text per thread
//   k_sdf_granules / k_sdf_mask : SDF bit planes -> resident 2-bit template + N mask + byte-per-base array
//   CUB            : exclusive scans (line numbers, stream positions, kept positions, slot and base offsets)
// One loader context per process (stream, events, grow-only device slab), calls queue on its mutex.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/vce.h"
#include "vce_bgzf.h"
#include "vce_cuda.h"
#include "vce_vcf.h"

namespace vce {

template <class F>
__global__ void __launch_bounds__(256) k_vcf_for(int n, F f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f(i);
}

constexpr int kInflateWarps = 4;   // BGZF blocks per CTA: 4 x 10.4 KB of tables + text ring (static shared memory)

__global__ void __launch_bounds__(kInflateWarps * 32) k_bgzf_inflate(int nBlocks, const BgzfBlock* blocks, const uint8_t* data,
                                                                   uint8_t* out, int32_t* status) {
  __shared__ InflateWs ws[kInflateWarps];
  const int w = (int) (threadIdx.x >> 5);
  const int b = (int) blockIdx.x * kInflateWarps + w;
  if (b >= nBlocks) return;
  const BgzfBlock blk = blocks[b];
  const int st = inflateBlock<CudaWarp>(data + blk.src, blk.srcLen, out + blk.dst, blk.dstLen, &ws[w]);
  if ((threadIdx.x & 31u) == 0) status[b] = st;
}

__global__ void __launch_bounds__(64) k_bgzf_crc(int nBlocks, const BgzfBlock* blocks, const uint8_t* out, int32_t* status) {
  __shared__ uint32_t table[256];
  for (int i = (int) threadIdx.x; i < 256; i += 64) table[i] = crc32TableEntry((uint32_t) i);
  __syncthreads();
  const int b = (int) (blockIdx.x * 64 + threadIdx.x);
  if (b >= nBlocks || status[b] != kInfOk) return;
  const BgzfBlock blk = blocks[b];
  if (crc32Bytes(out + blk.dst, blk.dstLen, table) != blk.crc) status[b] = kInfCrc;
}

__global__ void __launch_bounds__(256) k_sdf_granules(const uint8_t* file, int64_t first, int32_t len, int32_t granules, uint32_t* packed,
                                                     uint8_t* otherFlag, uint8_t* bases) {
  const int g = (int) (blockIdx.x * 256 + threadIdx.x);
  if (g < granules) sdfGranule(file, first, len, g, packed, otherFlag, bases);
}
__global__ void __launch_bounds__(256) k_sdf_mask(const uint8_t* otherFlag, int32_t granules, int32_t mwords, uint32_t* nmask) {
  const int m = (int) (blockIdx.x * 256 + threadIdx.x);
  if (m < mwords) sdfMaskWord(otherFlag, granules, m, nmask);
}

namespace {
// page-locked result blocks are recycled: pinning memory costs more than the copy it speeds up
std::mutex g_pinMu;
std::vector<std::pair<uint8_t*, size_t>> g_pinFree;
void pinRelease(uint8_t* p, size_t) {
  std::lock_guard<std::mutex> lk(g_pinMu);
  // the capacity travels in the 16 bytes before the block
  const size_t cap = *reinterpret_cast<size_t*>(p - 16);
  if (g_pinFree.size() < 6) g_pinFree.emplace_back(p, cap);
  else cudaFreeHost(p - 16);
}
uint8_t* pinAcquire(size_t bytes) {
  {
    std::lock_guard<std::mutex> lk(g_pinMu);
    size_t best = g_pinFree.size();
    for (size_t i = 0; i < g_pinFree.size(); ++i) {
      if (g_pinFree[i].second >= bytes && (best == g_pinFree.size() || g_pinFree[i].second < g_pinFree[best].second)) best = i;
    }
    if (best != g_pinFree.size()) {
      uint8_t* p = g_pinFree[best].first;
      g_pinFree.erase(g_pinFree.begin() + (long) best);
      return p;
    }
  }
  const size_t cap = bytes + bytes / 8 + 4096;
  uint8_t* raw = nullptr;
  if (cudaHostAlloc(reinterpret_cast<void**>(&raw), cap + 16, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  *reinterpret_cast<size_t*>(raw) = cap;
  return raw + 16;
}
void pinDrain() {
  std::lock_guard<std::mutex> lk(g_pinMu);
  for (auto& e : g_pinFree) cudaFreeHost(e.first - 16);
  g_pinFree.clear();
}

struct CudaVcfOps {
  cudaStream_t st = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  uint8_t* slab = nullptr;
  size_t cap = 0, used = 0;
  void* tmp = nullptr;
  size_t tmpBytes = 0;
  int launches = 0;
  cudaError_t err = cudaSuccess;
  std::vector<void*> overflow;   // allocations that did not fit the slab (freed after the call; the slab then grows)
  size_t wanted = 0;

  void note(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }
  template <class T>
  T* alloc(size_t n) {
    const size_t bytes = (n * sizeof(T) + 255) & ~(size_t) 255;
    wanted += bytes;
    if (used + bytes <= cap) {
      T* p = reinterpret_cast<T*>(slab + used);
      used += bytes;
      return p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { note(cudaErrorMemoryAllocation); return nullptr; }
    overflow.push_back(p);
    return static_cast<T*>(p);
  }
  void upload(void* d, const void* h, size_t bytes) { if (bytes) note(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st)); }
  void download(void* h, const void* d, size_t bytes) {
    if (!bytes) return;
    note(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
    note(cudaStreamSynchronize(st));
  }
  template <class F>
  void parFor(int n, F f) {
    if (n <= 0 || err != cudaSuccess) return;
    k_vcf_for<<<(n + 255) / 256, 256, 0, st>>>(n, f);
    note(cudaGetLastError());
    ++launches;
  }
  void exclusiveSum(int32_t* in, int32_t* out, int n) {
    if (n <= 0 || err != cudaSuccess) return;
    size_t bytes = 0;
    note(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st));
    if (bytes > tmpBytes) {
      if (tmp) cudaFree(tmp);
      tmp = nullptr;
      tmpBytes = 0;
      if (cudaMalloc(&tmp, bytes + 4096) == cudaSuccess) tmpBytes = bytes + 4096; else { note(cudaErrorMemoryAllocation); return; }
    }
    note(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, st));
    launches += 2;
  }
  int fetchInt(const int32_t* p) {
    if (err != cudaSuccess) return 0;
    int32_t h = 0;
    note(cudaMemcpyAsync(&h, p, sizeof(h), cudaMemcpyDeviceToHost, st));
    note(cudaStreamSynchronize(st));
    return h;
  }
  bool chained = false;   // bgzfDrive recorded the first two marks; vcfDrive's own are passed over
  void mark(int k) { if (ev[k] && !(chained && k < 2)) note(cudaEventRecord(ev[k], st)); }
  void inflate(int nb, const BgzfBlock* blocks, const uint8_t* data, uint8_t* out, int32_t* status, bool checkCrc) {
    if (nb <= 0 || err != cudaSuccess) return;
    k_bgzf_inflate<<<(nb + kInflateWarps - 1) / kInflateWarps, kInflateWarps * 32, 0, st>>>(nb, blocks, data, out, status);
    note(cudaGetLastError());
    ++launches;
    if (checkCrc) {
      k_bgzf_crc<<<(nb + 63) / 64, 64, 0, st>>>(nb, blocks, out, status);
      note(cudaGetLastError());
      ++launches;
    }
  }
  uint8_t* hostAlloc(size_t bytes, void (**release)(uint8_t*, size_t)) {
    *release = &pinRelease;
    return pinAcquire(bytes);
  }
};

struct VcfCtx {
  std::mutex mu;
  int device = -1;
  CudaVcfOps ops;
};
VcfCtx g_vcf;

void vcfRelease() {
  CudaVcfOps& o = g_vcf.ops;
  if (g_vcf.device < 0) return;
  cudaSetDevice(g_vcf.device);
  if (o.st) cudaStreamSynchronize(o.st);
  for (cudaEvent_t& e : o.ev) { if (e) cudaEventDestroy(e); e = nullptr; }
  if (o.slab) cudaFree(o.slab);
  if (o.tmp) cudaFree(o.tmp);
  if (o.st) cudaStreamDestroy(o.st);
  o = CudaVcfOps();
  g_vcf.device = -1;
  pinDrain();
}
}  // namespace

// Makes the loader context current on `device` (one context per process, calls queue on g_vcf.mu, held by the caller).
static int vcfAcquire(int device, std::string& err) {
  if (g_vcf.device != device) {
    vcfRelease();
    if (cudaSetDevice(device) != cudaSuccess) { err = "cudaSetDevice failed"; return VCE_ERR_CUDA; }
    CudaVcfOps& o = g_vcf.ops;
    if (cudaStreamCreateWithFlags(&o.st, cudaStreamNonBlocking) != cudaSuccess) { err = "cudaStreamCreate failed"; return VCE_ERR_CUDA; }
    for (cudaEvent_t& e : o.ev) if (cudaEventCreate(&e) != cudaSuccess) { err = "cudaEventCreate failed"; return VCE_ERR_CUDA; }
    g_vcf.device = device;
  }
  if (cudaSetDevice(device) != cudaSuccess) { err = "cudaSetDevice failed"; return VCE_ERR_CUDA; }
  CudaVcfOps& o = g_vcf.ops;
  o.used = 0; o.wanted = 0; o.launches = 0; o.err = cudaSuccess; o.chained = false;
  return VCE_OK;
}

// After a call: event times, slab growth, error mapping.  `marks`: how many of the four events were recorded.
static int vcfSettle(int rc, int marks, float ms[3], std::string& err, const char* what) {
  CudaVcfOps& o = g_vcf.ops;
  cudaStreamSynchronize(o.st);
  ms[0] = ms[1] = ms[2] = 0;
  if (rc == 0 && o.err == cudaSuccess) {
    for (int k = 0; k + 1 < marks; ++k) cudaEventElapsedTime(&ms[k], o.ev[k], o.ev[k + 1]);
  }
  // what did not fit the slab this time fits next time
  for (void* p : o.overflow) cudaFree(p);
  if (!o.overflow.empty() || o.wanted > o.cap) {
    o.overflow.clear();
    if (o.slab) cudaFree(o.slab);
    o.slab = nullptr;
    o.cap = 0;
    const size_t want = o.wanted + o.wanted / 8 + (1 << 20);
    if (cudaMalloc(reinterpret_cast<void**>(&o.slab), want) == cudaSuccess) o.cap = want;
  }
  if (o.err != cudaSuccess) {
    err = std::string(what) + ": " + cudaGetErrorString(o.err);
    cudaGetLastError();
    return VCE_ERR_CUDA;
  }
  if (rc == 3) return VCE_ERR_BAD_ARGUMENT;
  if (rc == 2) return VCE_ERR_CUDA;
  if (rc == 1) return VCE_ERR_BAD_ARGUMENT;
  return VCE_OK;
}

// Called by vce_capi.cpp (which owns the device choice and the error string).
int vcfParseOnDevice(int device, const uint8_t* text, int64_t len, const VcfConfig& cfg, VcfResult& R, std::string& err) {
  std::lock_guard<std::mutex> lk(g_vcf.mu);
  if (int rc = vcfAcquire(device, err)) return rc;
  CudaVcfOps& o = g_vcf.ops;
  const int rc = vcfDrive(o, text, len, cfg, R, err);
  float ms[3];
  const int out = vcfSettle(rc, 4, ms, err, "VCF loader");
  R.stats.h2dMs = ms[0]; R.stats.kernelMs = ms[1]; R.stats.d2hMs = ms[2];
  R.stats.launches = o.launches;
  return out;
}

// BGZF text between two virtual offsets, inflated on the device and copied back (vce_bgzf_inflate).
int bgzfInflateOnDevice(int device, const uint8_t* data, int64_t len, int64_t vBeg, int64_t vEnd, bool checkCrc, std::vector<uint8_t>& text,
                        BgzfStats& st, std::string& err) {
  std::lock_guard<std::mutex> lk(g_vcf.mu);
  if (int rc = vcfAcquire(device, err)) return rc;
  CudaVcfOps& o = g_vcf.ops;
  const uint8_t* dText = nullptr;
  int64_t textLen = 0;
  const int rc = bgzfDrive(o, data, len, vBeg, vEnd, checkCrc, &dText, &textLen, nullptr, st, err);
  if (rc == 0) {
    o.mark(2);
    text.resize((size_t) textLen);
    o.download(text.data(), dText, (size_t) textLen);
    o.mark(3);
  }
  float ms[3];
  const int out = vcfSettle(rc, 4, ms, err, "BGZF decoder");
  st.h2dMs = ms[0]; st.kernelMs = ms[1]; st.d2hMs = ms[2];
  st.launches = o.launches;
  return out;
}

// The chained form: compressed blocks in, variants out; the text never leaves the device.
int vcfParseBgzfOnDevice(int device, const uint8_t* data, int64_t len, int64_t vBeg, int64_t vEnd, bool checkCrc, const VcfConfig& cfg,
                         VcfResult& R, BgzfStats& st, std::string& err) {
  std::lock_guard<std::mutex> lk(g_vcf.mu);
  if (int rc = vcfAcquire(device, err)) return rc;
  CudaVcfOps& o = g_vcf.ops;
  const uint8_t* dText = nullptr;
  int64_t textLen = 0;
  int rc = bgzfDrive(o, data, len, vBeg, vEnd, checkCrc, &dText, &textLen, nullptr, st, err);
  if (rc == 0) {
    o.chained = true;
    rc = vcfDrive(o, nullptr, textLen, cfg, R, err, dText);
  }
  float ms[3];
  const int out = vcfSettle(rc, 4, ms, err, "BGZF + VCF loader");
  R.stats.h2dMs = ms[0]; R.stats.kernelMs = ms[1]; R.stats.d2hMs = ms[2];
  R.stats.launches = o.launches;
  st.h2dMs = ms[0]; st.kernelMs = ms[1]; st.d2hMs = ms[2];
  st.launches = o.launches;
  return out;
}

// SDF sequence data -> resident template.  `file` holds the bit-plane longs from group `firstGroup` on (64 residues per
// group of three longs); `first` is the sequence's first residue relative to that group's first residue.
int sdfTemplateOnDevice(int device, const uint8_t* file, int64_t fileLen, int64_t first, int32_t len, uint8_t* basesOut, double ms[3],
                        std::string& err) {
  if (first < 0 || len <= 0 || ((first + len + 63) >> 6) * 24 > fileLen) { err = "SDF: the data does not cover the sequence"; return VCE_ERR_BAD_ARGUMENT; }
  std::lock_guard<std::mutex> lk(g_vcf.mu);
  if (int rc = vcfAcquire(device, err)) return rc;
  CudaVcfOps& o = g_vcf.ops;
  const int32_t granules = (len + 15) / 16, mwords = (granules + 31) / 32;
  const size_t need = (size_t) (((first + len + 63) >> 6) * 24);
  uint32_t* d2 = nullptr;
  uint32_t* dn = nullptr;
  // the template arrays outlive the call: they belong to the registry (same slack as templateRegistryAdd)
  if (cudaMalloc(&d2, ((size_t) granules + 8) * 4) != cudaSuccess || cudaMalloc(&dn, ((size_t) mwords + 8) * 4) != cudaSuccess) {
    if (d2) cudaFree(d2);
    cudaGetLastError();
    err = "out of device memory for the resident template";
    return VCE_ERR_CUDA;
  }
  o.mark(0);
  uint8_t* dFile = o.alloc<uint8_t>(need + 16);
  uint8_t* dFlag = o.alloc<uint8_t>((size_t) granules + 32);
  uint8_t* dBases = o.alloc<uint8_t>((size_t) len + 16);
  int rc = 0;
  if (!dFile || !dFlag || !dBases) { err = "out of device memory for the SDF data"; rc = 2; }
  if (rc == 0) {
    o.note(cudaMemsetAsync(d2 + granules, 0, 32, o.st));
    o.note(cudaMemsetAsync(dn + mwords, 0, 32, o.st));
    o.upload(dFile, file, need);
    o.mark(1);
    k_sdf_granules<<<(granules + 255) / 256, 256, 0, o.st>>>(dFile, first, len, granules, d2, dFlag, dBases);
    o.note(cudaGetLastError());
    k_sdf_mask<<<(mwords + 255) / 256, 256, 0, o.st>>>(dFlag, granules, mwords, dn);
    o.note(cudaGetLastError());
    o.launches += 2;
    o.mark(2);
    o.download(basesOut, dBases, (size_t) len);
    o.mark(3);
  }
  float f[3];
  int out = vcfSettle(rc, 4, f, err, "SDF decoder");
  if (ms) { ms[0] = f[0]; ms[1] = f[1]; ms[2] = f[2]; }
  if (out == VCE_OK) out = templateRegistryAdopt(device, basesOut, len, d2, dn, err);
  else { cudaFree(d2); cudaFree(dn); }
  return out;
}

void vcfShutdown() {
  std::lock_guard<std::mutex> lk(g_vcf.mu);
  vcfRelease();
}

}  // namespace vce
