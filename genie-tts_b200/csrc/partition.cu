// Spatial partitioning of one B200 between the two kinds of work of the synthesis path (CUDA green contexts,
// driver API 12.4+, resolved at run time through cudaGetDriverEntryPoint: no link-time dependency on libcuda).
//
// The T2S decode step is a chain of ~170 latency-bound kernels per token; the SoVITS decoder and the T2S prefill
// are throughput-bound kernels whose CTAs fill every SM.  Time-sharing them does not work: a decode kernel that
// becomes ready while vocoder CTAs own the shared memory / registers of all 148 SMs waits for CTAs to retire
// before each of its 170 dependent launches (measured: decode of one batch + vocoder of another at the same time
// take the SUM of their solo times).  With the SMs split into a decode partition and a bulk partition, a batch's
// decode runs next to another batch's prefill / vocoder without either waiting for the other's CTAs.
//
// One partition pair per device, shared by every handle that opts in (option sm_partition): the handles' decode
// streams are created in the decode partition, their bulk streams in the bulk partition; two tokens (mutexes)
// make the handles alternate — at any time at most one handle decodes and at most one runs a bulk stage.
#include "model.h"
#include <cuda.h>
#include <map>

namespace genie {
namespace {

template <typename F> F driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return reinterpret_cast<F>(p);
}

std::mutex g_part_mu;
std::map<int, std::unique_ptr<DevicePartition>> g_parts;

}  // namespace

DevicePartition* device_partition(int device, int decode_sms) {
  std::lock_guard<std::mutex> lock(g_part_mu);
  auto it = g_parts.find(device);
  if (it != g_parts.end()) {
    GENIE_CHECK(it->second->decode_sms_requested == decode_sms,
                "this device is already partitioned with a different decode SM count");
    return it->second.get();
  }
  using GetRes = CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType);
  using Split = CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  using GenDesc = CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  using Create = CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  using DevGet = CUresult (*)(CUdevice*, int);
  auto get_res = driver_fn<GetRes>("cuDeviceGetDevResource");
  auto split = driver_fn<Split>("cuDevSmResourceSplitByCount");
  auto gen_desc = driver_fn<GenDesc>("cuDevResourceGenerateDesc");
  auto create = driver_fn<Create>("cuGreenCtxCreate");
  auto dev_get = driver_fn<DevGet>("cuDeviceGet");
  GENIE_CHECK(get_res && split && gen_desc && create && dev_get, "green contexts are not available in this driver");
  GENIE_CUDA(cudaSetDevice(device));
  GENIE_CUDA(cudaFree(nullptr));                       // make sure the primary context exists
  CUdevice dev;
  GENIE_CHECK(dev_get(&dev, device) == CUDA_SUCCESS, "cuDeviceGet failed");
  CUdevResource all, dec, rest;
  GENIE_CHECK(get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS, "cuDeviceGetDevResource failed");
  unsigned int groups = 1;
  CUresult rc = split(&dec, &groups, &all, &rest, 0, (unsigned)decode_sms);
  GENIE_CHECK(rc == CUDA_SUCCESS && groups == 1, "cuDevSmResourceSplitByCount failed (" + std::to_string((int)rc) + ")");
  GENIE_CHECK(rest.sm.smCount >= 8, "sm_partition leaves fewer than 8 SMs for the bulk stages");
  std::unique_ptr<DevicePartition> p(new DevicePartition());
  p->device = device; p->decode_sms_requested = decode_sms;
  p->decode_sms = (int)dec.sm.smCount; p->bulk_sms = (int)rest.sm.smCount;
  CUdevResourceDesc d1, d2;
  GENIE_CHECK(gen_desc(&d1, &dec, 1) == CUDA_SUCCESS && gen_desc(&d2, &rest, 1) == CUDA_SUCCESS,
              "cuDevResourceGenerateDesc failed");
  CUgreenCtx g1, g2;
  GENIE_CHECK(create(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS, "cuGreenCtxCreate (decode) failed");
  GENIE_CHECK(create(&g2, d2, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS, "cuGreenCtxCreate (bulk) failed");
  p->green_decode = g1; p->green_bulk = g2;
  DevicePartition* out = p.get();
  g_parts[device] = std::move(p);
  return out;
}

cudaStream_t partition_stream(DevicePartition* p, bool decode, int priority) {
  using StreamCreate = CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int);
  static auto fn = driver_fn<StreamCreate>("cuGreenCtxStreamCreate");
  GENIE_CHECK(fn != nullptr, "cuGreenCtxStreamCreate is not available");
  CUstream s = nullptr;
  CUresult rc = fn(&s, reinterpret_cast<CUgreenCtx>(decode ? p->green_decode : p->green_bulk), CU_STREAM_NON_BLOCKING,
                   priority);
  GENIE_CHECK(rc == CUDA_SUCCESS, "cuGreenCtxStreamCreate failed (" + std::to_string((int)rc) + ")");
  return reinterpret_cast<cudaStream_t>(s);
}

}  // namespace genie
