// tcgen05 self-test: Y[128 x 32] = X[128 x K] . W[32 x K]^T with the 3xTF32 split, exercising
// exactly the primitives (operand layout, descriptors, TMEM alloc/ld, mbarrier commit) that the
// fused decoder kernels use.  One CTA of 128 threads; thread t owns row t.
#include "pn_common.cuh"
#include "pn_umma.cuh"

namespace pn {
namespace {

constexpr int kTcMaxK = 128;

__global__ void __launch_bounds__(128, 1) k_tc_selftest(const float* __restrict__ X, const float* __restrict__ W,
                                                       float* __restrict__ Y, int K) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  const uint32_t lbo = 128, a_sbo = (uint32_t)(K / 4) * 128, b_sbo = a_sbo;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + 16 * a_sbo;  // 128 rows = 16 groups
  unsigned char* b_hi = a_lo + 16 * a_sbo;
  unsigned char* b_lo = b_hi + 4 * b_sbo;   // 32 rows = 4 groups
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 32);
  if (t == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
  // A: thread t writes its row
  for (int k = 0; k < K; ++k) {
    float hi, lo;
    umma::split_tf32(X[t * K + k], hi, lo);
    const uint32_t off = umma::kmajor_off(t, k, lbo, a_sbo);
    *reinterpret_cast<float*>(a_hi + off) = hi;
    *reinterpret_cast<float*>(a_lo + off) = lo;
  }
  for (int i = t; i < 32 * K; i += 128) {
    const int n = i / K, k = i % K;
    float hi, lo;
    umma::split_tf32(W[i], hi, lo);
    const uint32_t off = umma::kmajor_off(n, k, lbo, b_sbo);
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = lo;
  }
  umma::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    umma::mma_3xtf32(tmem, umma::smem_u32(a_hi), umma::smem_u32(a_lo), lbo, a_sbo, umma::smem_u32(b_hi),
                     umma::smem_u32(b_lo), lbo, b_sbo, K, umma::instr_desc_tf32(128, 32), 0u);
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();
  float v[32];
  umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
  for (int j = 0; j < 32; ++j) Y[t * 32 + j] = v[j];
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 32);
}

// MN-major variant: Y[128 x 32] = Xt^T . Wt with Xt (K x 128), Wt (K x 32) sample-major, i.e. both
// operands "MN-major" (reduction index = row).  smem: element (mn, k) at (mn/4)*SBO + k*16 + (mn%4)*4.
__global__ void __launch_bounds__(128, 1) k_tc_selftest_mn(const float* __restrict__ Xt, const float* __restrict__ Wt,
                                                          float* __restrict__ Y, int K, int swap) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  const uint32_t row = (uint32_t)K * 16;           // bytes of one quad row
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + 32 * row;
  unsigned char* b_hi = a_lo + 32 * row;
  unsigned char* b_lo = b_hi + 8 * row;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 32);
  if (t == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
  for (int i = t; i < K * 128; i += 128) {
    const int k = i / 128, m = i % 128;
    float hi, lo;
    umma::split_tf32(Xt[i], hi, lo);
    const uint32_t off = (uint32_t)(m >> 2) * row + (uint32_t)k * 16u + (uint32_t)(m & 3) * 4u;
    *reinterpret_cast<float*>(a_hi + off) = hi;
    *reinterpret_cast<float*>(a_lo + off) = lo;
  }
  for (int i = t; i < K * 32; i += 128) {
    const int k = i / 32, n = i % 32;
    float hi, lo;
    umma::split_tf32(Wt[i], hi, lo);
    const uint32_t off = (uint32_t)(n >> 2) * row + (uint32_t)k * 16u + (uint32_t)(n & 3) * 4u;
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = lo;
  }
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    const uint32_t idesc = umma::instr_desc_tf32(128, 32) | (1u << 15) | (1u << 16);
    const uint32_t lbo = swap ? row : 128u, sbo = swap ? 128u : row;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t o = (uint32_t)ks * 128u;
      const uint64_t ah = umma::smem_desc(umma::smem_u32(a_hi) + o, lbo, sbo), al = umma::smem_desc(umma::smem_u32(a_lo) + o, lbo, sbo);
      const uint64_t bh = umma::smem_desc(umma::smem_u32(b_hi) + o, lbo, sbo), bl = umma::smem_desc(umma::smem_u32(b_lo) + o, lbo, sbo);
      umma::mma_tf32(tmem, al, bh, idesc, ks > 0);
      umma::mma_tf32(tmem, ah, bl, idesc, 1u);
      umma::mma_tf32(tmem, ah, bh, idesc, 1u);
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after();
  float v[32];
  umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
  for (int j = 0; j < 32; ++j) Y[t * 32 + j] = v[j];
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 32);
}

}  // namespace
}  // namespace pn

using namespace pn;

extern "C" int pn_tc_selftest_mn(const float* Xt, const float* Wt, float* Y, int K, int swap, void* stream) {
  if (!Xt || !Wt || !Y || K < 8 || K > kTcMaxK || (K % 8) != 0) { set_error("pn_tc_selftest_mn: K must be a multiple of 8 in [8,128]"); return 1; }
  const int smem = 2 * (32 + 8) * K * 16;
  cudaFuncSetAttribute(k_tc_selftest_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_tc_selftest_mn<<<1, 128, smem, (cudaStream_t)stream>>>(Xt, Wt, Y, K, swap);
  return launch_status("k_tc_selftest_mn");
}

extern "C" int pn_tc_selftest(const float* X, const float* W, float* Y, int K, void* stream) {
  if (!X || !W || !Y || K < 8 || K > kTcMaxK || (K % 8) != 0) { set_error("pn_tc_selftest: K must be a multiple of 8 in [8,128]"); return 1; }
  const int smem = 2 * (16 + 4) * (K / 4) * 128;
  cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_tc_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(X, W, Y, K);
  return launch_status("k_tc_selftest");
}
