// vce_cuda.h -- entry points of the CUDA translation units used by the C-ABI layer.
#ifndef VCE_CUDA_H_
#define VCE_CUDA_H_

#include <string>

#include "../../include/vce.h"
#include "vce_pipeline.h"

namespace vce {

int cudaDeviceCount();
// Creates the CUDA backend on `device`; returns nullptr and fills `err` on failure.
Backend* createCudaBackend(int device, std::string& err);
// Device milliseconds and region count of the last shared-memory search launch (roofline reporting).
void cudaBackendFastStats(Backend* b, double* ms, long long* regions, long long* variants);

// Device milliseconds and region count of the last single-launch thread-per-region search (staged jobs).
void cudaBackendMicroStats(Backend* b, double* ms, long long* regions);

struct RocTiming { double h2dMs = 0, kernelMs = 0; int launches = 0; };
int rocRun(int device, const vce_roc_in* in, vce_roc_out* out, std::string& err, RocTiming* timing);

}  // namespace vce
#endif
