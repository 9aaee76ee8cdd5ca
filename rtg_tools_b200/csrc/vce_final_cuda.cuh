// vce_final_cuda.cuh -- CUDA instantiation of the final passes (vce_final.h): element-wise bodies as kernels, scans and
// the de-duplication through CUB, device buffers that persist across calls.  Included by vce_cuda.cu only.
#ifndef VCE_FINAL_CUDA_CUH_
#define VCE_FINAL_CUDA_CUH_

#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "vce_final.h"

namespace vce {

template <class F>
__global__ void __launch_bounds__(256) k_final_for(int n, F f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f(i);
}
__global__ void __launch_bounds__(256) k_scatter(const ScatterArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < a.n) scatterOne(a, q);
}
__global__ void __launch_bounds__(128) k_patch(const PatchArgs a) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < a.n) patchOne(a, q);
}

struct MinOp {
  __host__ __device__ __forceinline__ int32_t operator()(int32_t x, int32_t y) const { return x < y ? x : y; }
};

struct CudaFinalOps {
  cudaStream_t st = nullptr;
  void* tmp = nullptr;
  size_t tmpBytes = 0;
  int32_t* dCount = nullptr;   // one device int for the selection count
  int64_t launches = 0;
  cudaError_t err = cudaSuccess;
  void note(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }
  void needTmp(size_t bytes) {
    if (bytes <= tmpBytes) return;
    if (tmp) cudaFree(tmp);
    tmp = nullptr;
    tmpBytes = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaMalloc(&tmp, want) == cudaSuccess) tmpBytes = want; else note(cudaErrorMemoryAllocation);
  }
  template <class F>
  void parFor(int n, F f) {
    if (n <= 0 || err != cudaSuccess) return;
    k_final_for<<<(n + 255) / 256, 256, 0, st>>>(n, f);
    note(cudaGetLastError());
    ++launches;
  }
  void exclusiveSum(int32_t* in, int32_t* out, int n) {
    if (n <= 0 || err != cudaSuccess) return;
    size_t bytes = 0;
    note(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st));
    needTmp(bytes);
    if (err != cudaSuccess) return;
    note(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, st));
    launches += 2;
  }
  int unique64(int64_t* in, int64_t* out, int n) {
    if (n <= 0 || err != cudaSuccess) return 0;
    size_t bytes = 0;
    note(cub::DeviceSelect::Unique(nullptr, bytes, in, out, dCount, n, st));
    needTmp(bytes);
    if (err != cudaSuccess) return 0;
    note(cub::DeviceSelect::Unique(tmp, bytes, in, out, dCount, n, st));
    launches += 2;
    return fetchInt(dCount);
  }
  void segMinScan(const int32_t* keys, const int32_t* vals, int32_t* out, int n) {
    if (n <= 0 || err != cudaSuccess) return;
    size_t bytes = 0;
    note(cub::DeviceScan::InclusiveScanByKey(nullptr, bytes, keys, vals, out, MinOp(), n, cub::Equality(), st));
    needTmp(bytes);
    if (err != cudaSuccess) return;
    note(cub::DeviceScan::InclusiveScanByKey(tmp, bytes, keys, vals, out, MinOp(), n, cub::Equality(), st));
    launches += 2;
  }
  int fetchInt(const int32_t* p) {
    if (err != cudaSuccess) return 0;
    int32_t h = 0;
    note(cudaMemcpyAsync(&h, p, sizeof(h), cudaMemcpyDeviceToHost, st));
    note(cudaStreamSynchronize(st));
    return h;
  }
  void zero(void* p, size_t bytes) {
    if (bytes == 0 || err != cudaSuccess) return;
    note(cudaMemsetAsync(p, 0, bytes, st));
  }
};

}  // namespace vce
#endif
