// probe.cu — on-device unit checks of the sm_100a primitives in csrc/sm100.cuh.
// Each case runs one 128x128x128 bf16 MMA (8 tcgen05.mma K-steps) with a given operand
// placement and compares D against a host fp32 reference.  Build: see tools/probe/build.sh.
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../../magiattention_b200/csrc/sm100.cuh"

using namespace sm100;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

typedef CUresult (*EncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled_t get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) {
    printf("no cuTensorMapEncodeTiled\n");
    exit(2);
  }
  return (EncodeTiled_t)fn;
}

struct Args {
  int a_src;          // 0 smem, 2 tmem
  uint64_t a_hi;      // descriptor high bits for A
  uint32_t a_kstep;   // byte advance per K step inside a box
  uint32_t a_kbox;    // K steps per box (then jump a_boxjump)
  uint32_t a_boxjump;
  uint64_t b_hi;
  uint32_t b_kstep, b_kbox, b_boxjump;
  uint32_t idesc;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __nv_bfloat16* __restrict__ Ag, float* __restrict__ D, Args args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 2 boxes x 16 KB
  uint8_t* sB = smem + 32768;    // 2 boxes x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 65536 + 64);
  const uint32_t bar_tma = smem_u32(&bars[0]);
  const uint32_t bar_mma = smem_u32(&bars[1]);
  const int tid = threadIdx.x, wid = tid >> 5;

  if (tid == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (wid == 0) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    uint32_t bytes = 32768 + (args.a_src == 0 ? 32768 : 0);
    mbar_arrive_expect_tx(bar_tma, bytes);
    if (args.a_src == 0) {
      tma_load_2d(smem_u32(sA), &tmA, bar_tma, 0, 0);
      tma_load_2d(smem_u32(sA + 16384), &tmA, bar_tma, 64, 0);
    }
    tma_load_2d(smem_u32(sB), &tmB, bar_tma, 0, 0);
    tma_load_2d(smem_u32(sB + 16384), &tmB, bar_tma, 64, 0);
  }
  if (args.a_src == 2) {
    // A (row = tid) -> TMEM columns [128,192) as packed bf16 pairs
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(Ag + size_t(tid) * 128);
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = arow[c * 32 + i];
      tmem_st32(tmem + (uint32_t(wid * 32) << 16) + 128 + c * 32, r);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (wid == 0) {
    mbar_wait(bar_tma, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll 1
      for (uint32_t k = 0; k < 8; ++k) {
        uint32_t boff = (k / args.b_kbox) * args.b_boxjump + (k % args.b_kbox) * args.b_kstep;
        uint64_t bd = smem_desc(args.b_hi, smem_u32(sB) + boff);
        if (args.a_src == 0) {
          uint32_t aoff = (k / args.a_kbox) * args.a_boxjump + (k % args.a_kbox) * args.a_kstep;
          uint64_t ad = smem_desc(args.a_hi, smem_u32(sA) + aoff);
          mma_ss(tmem, ad, bd, args.idesc, k > 0);
        } else {
          mma_ts(tmem, tmem + 128 + k * 8, bd, args.idesc, k > 0);
        }
      }
      tc_commit(bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld32(tmem + (uint32_t(wid * 32) << 16) + c * 32, r);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) D[size_t(tid) * 128 + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc(tmem, 256);
}

// red.global.add.v4.f32 check (used by the backward dQ reduction)
__global__ void red_probe(float* out) {
  float* p = out + threadIdx.x * 4;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f)
               : "memory");
}

static float bf2f(__nv_bfloat16 x) { return __bfloat162float(x); }

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d SMs=%d smem/block optin=%zu\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  EncodeTiled_t encode = get_encode();

  const int N = 128;
  std::vector<__nv_bfloat16> hA(N * N), hB(N * N);
  srand(1234);
  for (int i = 0; i < N * N; ++i) {
    hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
  }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, N * N * 2));
  CK(cudaMalloc(&dB, N * N * 2));
  CK(cudaMalloc(&dD, N * N * 4));
  CK(cudaMemcpy(dA, hA.data(), N * N * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), N * N * 2, cudaMemcpyHostToDevice));

  auto make_map = [&](void* ptr) {
    CUtensorMap m;
    cuuint64_t dims[2] = {128, 128};
    cuuint64_t strides[1] = {128 * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("encode failed %d\n", (int)r);
      exit(2);
    }
    return m;
  };
  CUtensorMap tmA = make_map(dA), tmB = make_map(dB);

  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024));

  struct Case {
    const char* name;
    int a_src;   // 0 smem, 2 tmem
    int a_mn;    // A is MN-major in smem
    int b_mn;
    int variant; // 0 = expected LBO/SBO, 1 = swapped LBO/SBO for MN-major operands
  };
  Case cases[] = {
      {"SS  A=K-major  B=K-major  (S=QK^T)", 0, 0, 0, 0},
      {"SS  A=K-major  B=MN-major (O=P V, smem P)", 0, 0, 1, 0},
      {"SS  A=K-major  B=MN-major [swapped LBO/SBO]", 0, 0, 1, 1},
      {"TS  A=TMEM     B=MN-major (O=P V)", 2, 0, 1, 0},
      {"TS  A=TMEM     B=K-major", 2, 0, 0, 0},
      {"SS  A=MN-major B=MN-major (dQ=dS K)", 0, 1, 1, 0},
      {"SS  A=MN-major B=K-major", 0, 1, 0, 0},
  };
  int n_fail = 0;
  for (const Case& c : cases) {
    Args a{};
    a.a_src = c.a_src;
    uint64_t mn_hi = c.variant == 0 ? desc_mnmajor(16384) : make_smem_desc_hi(1024, 16384);
    if (c.a_mn) { a.a_hi = mn_hi; a.a_kstep = 2048; a.a_kbox = 8; a.a_boxjump = 0; }
    else        { a.a_hi = DESC_KMAJOR; a.a_kstep = 32; a.a_kbox = 4; a.a_boxjump = 16384; }
    if (c.b_mn) { a.b_hi = mn_hi; a.b_kstep = 2048; a.b_kbox = 8; a.b_boxjump = 0; }
    else        { a.b_hi = DESC_KMAJOR; a.b_kstep = 32; a.b_kbox = 4; a.b_boxjump = 16384; }
    a.idesc = make_idesc(true, c.a_mn, c.b_mn, 128, 128);
    CK(cudaMemset(dD, 0xff, N * N * 4));
    probe_kernel<<<1, 128, 70 * 1024>>>(tmA, tmB, dA, dD, a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("[FAIL] %-48s : launch error %s\n", c.name, cudaGetErrorString(e));
      return 3;  // context is dead
    }
    std::vector<float> hD(N * N);
    CK(cudaMemcpy(hD.data(), dD, N * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < N; ++m)
      for (int n = 0; n < N; ++n) {
        float acc = 0;
        for (int k = 0; k < N; ++k) {
          float av = c.a_mn ? bf2f(hA[k * N + m]) : bf2f(hA[m * N + k]);
          float bv = c.b_mn ? bf2f(hB[k * N + n]) : bf2f(hB[n * N + k]);
          acc += av * bv;
        }
        double err = fabs(double(acc) - double(hD[m * N + n]));
        if (!(err <= 1e30)) err = 1e30;
        maxerr = fmax(maxerr, err);
        maxref = fmax(maxref, fabs(acc));
      }
    bool ok = maxerr < 1e-2;
    if (!ok && c.variant == 0) n_fail++;
    printf("[%s] %-48s : max|err| = %.3e (max|ref| = %.2f)\n", ok ? "PASS" : "FAIL", c.name, maxerr, maxref);
  }
  {
    float* dR;
    CK(cudaMalloc(&dR, 128 * 4 * 4));
    CK(cudaMemset(dR, 0, 128 * 4 * 4));
    red_probe<<<1, 128>>>(dR);
    red_probe<<<1, 128>>>(dR);
    CK(cudaDeviceSynchronize());
    float h[8];
    CK(cudaMemcpy(h, dR, sizeof(h), cudaMemcpyDeviceToHost));
    bool ok = h[0] == 2.f && h[1] == 4.f && h[2] == 6.f && h[3] == 8.f && h[4] == 2.f;
    printf("[%s] red.global.add.v4.f32\n", ok ? "PASS" : "FAIL");
    if (!ok) n_fail++;
  }
  printf("probe done: %d failing expected-variant cases\n", n_fail);
  return n_fail ? 1 : 0;
}
