// nvl_exchange.cu — NVLink peer-memory primitives of the context-parallel KV / dKV exchange.
//
// Replaces, for ranks of one NVSwitch domain, the reference's a2av group-collective engine
// (magi_attention/comm/primitive/grpcoll/_a2av_grpcoll_impl.py:69-340 -> all_to_all_single + range_gather /
// range_reduce passes) and its NCCL kernels: bulk rows move with COPY ENGINES over peer-mapped memory
// (cudaMemcpyAsync on IPC-opened pointers), so no communication CTA ever competes with the attention grid for an SM;
// cross-rank ordering is a handful of 32-bit epoch flags written straight into the peer's memory by one-thread kernels
// (st.release.sys) and polled by one-warp kernels with a timeout (ld.acquire.sys), all stream-ordered.
//
//   b200_nvl_alloc / free          arena allocation (cudaMalloc'd, zero-filled: IPC-exportable, never from a pool)
//   b200_nvl_export / import / release   cudaIpc{Get,Open,Close}MemHandle
//   b200_nvl_copy_spans            n device-to-device span copies on a stream (copy engines, local or peer)
//   b200_nvl_signal                flags[i] <- value on up to 64 (peer) addresses, after everything earlier on the stream
//   b200_nvl_wait                  block the stream until every *flags[i] >= value (wrap-safe), or time out -> *err = 1
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../../include/b200_ffa.h"

namespace {

constexpr int kMaxFlags = 64;
struct FlagList {
  uint32_t* p[kMaxFlags];
};

__global__ void nvl_signal_kernel(FlagList f, int n, uint32_t value) {
  const int i = threadIdx.x;
  if (i < n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[i]), "r"(value) : "memory");
  }
}

__global__ void nvl_wait_kernel(FlagList f, int n, uint32_t value, uint32_t* err, unsigned long long timeout_ns) {
  const int i = threadIdx.x;
  bool ok = true;
  if (i < n) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t v;
    int spins = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.p[i]) : "memory");
      if (int32_t(v - value) >= 0) break;
      __nanosleep(200);
      if ((++spins & 1023) == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) { ok = false; break; }
      }
    }
  }
  if (!ok && err) {
    *reinterpret_cast<volatile uint32_t*>(err) = 1u;
    __threadfence_system();
  }
}

}  // namespace

extern "C" {

int b200_nvl_alloc(void** ptr, size_t bytes) {
  if (!ptr || !bytes) return B200_ERR_INVALID;
  if (cudaMalloc(ptr, bytes) != cudaSuccess) { cudaGetLastError(); return B200_ERR_CUDA; }
  if (cudaMemset(*ptr, 0, bytes) != cudaSuccess) { cudaGetLastError(); return B200_ERR_CUDA; }
  return cudaDeviceSynchronize() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

int b200_nvl_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? B200_OK : B200_ERR_CUDA; }

int b200_nvl_export(const void* ptr, uint8_t handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)) != cudaSuccess) { cudaGetLastError(); return B200_ERR_CUDA; }
  for (int i = 0; i < 64; ++i) handle[i] = reinterpret_cast<uint8_t*>(&h)[i];
  return B200_OK;
}

int b200_nvl_import(const uint8_t handle[64], void** ptr) {
  cudaIpcMemHandle_t h;
  for (int i = 0; i < 64; ++i) reinterpret_cast<uint8_t*>(&h)[i] = handle[i];
  if (cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return B200_ERR_CUDA; }
  return B200_OK;
}

int b200_nvl_release(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? B200_OK : B200_ERR_CUDA; }

int b200_nvl_copy_spans(int32_t n, void* const* dst, const void* const* src, const size_t* bytes, b200_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (int i = 0; i < n; ++i) {
    if (!bytes[i]) continue;
    if (cudaMemcpyAsync(dst[i], src[i], bytes[i], cudaMemcpyDeviceToDevice, stream) != cudaSuccess) {
      cudaGetLastError();
      return B200_ERR_CUDA;
    }
  }
  return B200_OK;
}

int b200_nvl_signal(void* const* flags, int32_t n, uint32_t value, b200_stream_t stream_) {
  if (n <= 0) return B200_OK;
  if (n > kMaxFlags || !flags) return B200_ERR_INVALID;
  FlagList f;
  for (int i = 0; i < n; ++i) f.p[i] = static_cast<uint32_t*>(flags[i]);
  nvl_signal_kernel<<<1, kMaxFlags, 0, static_cast<cudaStream_t>(stream_)>>>(f, n, value);
  return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

int b200_nvl_wait(void* const* flags, int32_t n, uint32_t value, uint32_t* err, uint64_t timeout_ns, b200_stream_t stream_) {
  if (n <= 0) return B200_OK;
  if (n > kMaxFlags || !flags) return B200_ERR_INVALID;
  FlagList f;
  for (int i = 0; i < n; ++i) f.p[i] = static_cast<uint32_t*>(flags[i]);
  nvl_wait_kernel<<<1, kMaxFlags, 0, static_cast<cudaStream_t>(stream_)>>>(f, n, value, err, timeout_ns);
  return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_CUDA;
}

}  // extern "C"
