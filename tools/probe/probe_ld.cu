// probe_ld.cu — register layout of tcgen05.ld shape 16x256b (and its cost next to 32x32b): fill 64 TMEM columns with
// value = lane*1000 + column through the known 32x32b store, read them back with 16x256b.x8 at lane bases 0 and 16, print
// which (lane, column) every (thread, register) received.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../magiattention_b200/csrc/sm100.cuh"
using namespace sm100;

__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe_kernel(uint32_t* out) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (wid == 0) { tmem_alloc(smem_u32(&slot), 128); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_off = uint32_t(wid * 32) << 16;
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
    for (int i = 0; i < 32; ++i) v[i] = uint32_t(tid * 1000 + half * 32 + i);
    tmem_st32(tmem + lane_off + half * 32, v);
  }
  tc_wait_st(); tc_fence_before(); __syncthreads(); tc_fence_after();
  for (int base = 0; base < 2; ++base) {  // lane base 0 / 16 inside this warp's quadrant
    uint32_t r[32];
    ld_16x256b_x8(tmem + lane_off + (uint32_t(base * 16) << 16), r);
    tc_wait_ld();
    for (int i = 0; i < 32; ++i) out[((wid * 2 + base) * 32 + lane) * 32 + i] = r[i];
  }
  tc_fence_before(); __syncthreads();
  if (wid == 0) tmem_dealloc(tmem, 128);
}

int main() {
  uint32_t* d; cudaMalloc(&d, 4 * 2 * 32 * 32 * 4);
  probe_kernel<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe_ld: %s\n", cudaGetErrorString(e));
  static uint32_t h[4 * 2 * 32 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int w = 0; w < 2; ++w) for (int base = 0; base < 2; ++base) {
    printf("warp %d lane-base %d: thread -> regs as (lane:col)\n", w, base * 16);
    for (int t = 0; t < 32; t += (t < 8 ? 1 : 8)) {
      printf("  t%02d:", t);
      for (int i = 0; i < 12; ++i) { uint32_t x = h[((w * 2 + base) * 32 + t) * 32 + i]; printf(" %u:%u", x / 1000, x % 1000); }
      printf(" ... r28-31:");
      for (int i = 28; i < 32; ++i) { uint32_t x = h[((w * 2 + base) * 32 + t) * 32 + i]; printf(" %u:%u", x / 1000, x % 1000); }
      printf("\n");
    }
  }
  // formula check: reg 4j+e of thread t  ==  lane (base + t/4 + 8*(e/2)), column 8j + 2*(t%4) + (e%2)
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int base = 0; base < 2; ++base) for (int t = 0; t < 32; ++t) for (int i = 0; i < 32; ++i) {
    int j = i / 4, e = i % 4;
    uint32_t want = uint32_t((w * 32 + base * 16 + t / 4 + 8 * (e / 2)) * 1000 + 8 * j + 2 * (t % 4) + (e % 2));
    if (h[((w * 2 + base) * 32 + t) * 32 + i] != want) ++bad;
  }
  printf("formula mismatches: %d\n", bad);
  return 0;
}
