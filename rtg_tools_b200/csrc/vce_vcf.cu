// vce_vcf.cu -- CUDA instantiation of the VCF loader passes (vce_vcf.h) and its C-ABI entry points.
//   k_vcf_for<F>   : one thread per byte / line / run of equal starts / kept variant (the pass bodies of vce_vcf.h)
//   CUB            : exclusive scans (line numbers, stream positions, kept positions, slot and base offsets)
// One loader context per process (stream, events, grow-only device slab), calls queue on its mutex.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/vce.h"
#include "vce_cuda.h"
#include "vce_vcf.h"

namespace vce {

template <class F>
__global__ void __launch_bounds__(256) k_vcf_for(int n, F f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f(i);
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
  void mark(int k) { if (ev[k]) note(cudaEventRecord(ev[k], st)); }
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

// Called by vce_capi.cpp (which owns the device choice and the error string).
int vcfParseOnDevice(int device, const uint8_t* text, int64_t len, const VcfConfig& cfg, VcfResult& R, std::string& err) {
  std::lock_guard<std::mutex> lk(g_vcf.mu);
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
  o.used = 0; o.wanted = 0; o.launches = 0; o.err = cudaSuccess;
  const int rc = vcfDrive(o, text, len, cfg, R, err);
  cudaStreamSynchronize(o.st);
  float a = 0, b = 0, c = 0;
  if (rc == 0 && o.err == cudaSuccess) {
    cudaEventElapsedTime(&a, o.ev[0], o.ev[1]);
    cudaEventElapsedTime(&b, o.ev[1], o.ev[2]);
    cudaEventElapsedTime(&c, o.ev[2], o.ev[3]);
  }
  R.stats.h2dMs = a; R.stats.kernelMs = b; R.stats.d2hMs = c;
  R.stats.launches = o.launches;
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
    err = std::string("VCF loader: ") + cudaGetErrorString(o.err);
    cudaGetLastError();
    return VCE_ERR_CUDA;
  }
  if (rc == 3) return VCE_ERR_BAD_ARGUMENT;
  if (rc == 2) return VCE_ERR_CUDA;
  if (rc == 1) return VCE_ERR_BAD_ARGUMENT;
  return VCE_OK;
}

void vcfShutdown() {
  std::lock_guard<std::mutex> lk(g_vcf.mu);
  vcfRelease();
}

}  // namespace vce
