// vce_cuda.h -- entry points of the CUDA translation units used by the C-ABI layer.
#ifndef VCE_CUDA_H_
#define VCE_CUDA_H_

#include <string>
#include <vector>

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

// Registered templates (vce_template_register): device-resident copies, 2 bits per base (base - 1, 16 bases per word) and
// one bit per 16-base granule that holds a base other than A/C/G/T.
bool templateRegistryLookup(int device, const uint8_t* bases, int32_t len, const uint32_t** t2, const uint32_t** nmask);
int templateRegistryAdd(int device, const uint8_t* bases, int32_t len, const uint32_t* packed, size_t words, const uint32_t* nmask,
                        size_t mwords, std::string& err);
int templateRegistryAdopt(int device, const uint8_t* bases, int32_t len, uint32_t* d2, uint32_t* dn, std::string& err);
void templateRegistryRemove(int device, const uint8_t* bases);   // device < 0: every device; bases == nullptr: every template

// Caller memory page-locked through vce_host_register (one registry for the process; the ranges are portable across devices).
bool pinnedRegistryCovers(const void* p, size_t bytes);
int pinnedRegistryAdd(const void* p, size_t bytes, std::string& err);
void pinnedRegistryRemove(const void* p);   // nullptr: everything

struct RocTiming { double h2dMs = 0, kernelMs = 0; int launches = 0; };
class WorkerPool;
// `pool`: host threads that stage the caller's (pageable) arrays into page-locked memory in parallel; nullptr = the caller's thread
int rocRun(int device, const vce_roc_in* in, vce_roc_out* out, std::string& err, RocTiming* timing, WorkerPool* pool);
void rocRelease();   // frees the per-device K4 contexts (vce_shutdown)

// VCF loader (vce_vcf.cu / vce_vcf.h)
struct VcfConfig;
struct VcfResult;
int vcfParseOnDevice(int device, const uint8_t* text, int64_t len, const VcfConfig& cfg, VcfResult& R, std::string& err);
void vcfShutdown();

// On-disk formats either side of the loader (vce_bgzf.h): BGZF blocks inflated on the device, alone or chained into the VCF
// parse without the text leaving the device; SDF bit planes decoded into a resident template.
struct BgzfStats;
int bgzfInflateOnDevice(int device, const uint8_t* data, int64_t len, int64_t vBeg, int64_t vEnd, bool checkCrc, std::vector<uint8_t>& text,
                        BgzfStats& st, std::string& err);
int vcfParseBgzfOnDevice(int device, const uint8_t* data, int64_t len, int64_t vBeg, int64_t vEnd, bool checkCrc, const VcfConfig& cfg,
                         VcfResult& R, BgzfStats& st, std::string& err);
int sdfTemplateOnDevice(int device, const uint8_t* file, int64_t fileLen, int64_t first, int32_t len, uint8_t* basesOut, double ms[3],
                        std::string& err);

}  // namespace vce
#endif
