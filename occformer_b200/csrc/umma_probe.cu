// Development probe (not on the product path): D[128x32] = A[128x64] * V[64x32] through tcgen05.mma kind::tf32 in
// four operand configurations, to pin down descriptor / TMEM-operand conventions on real hardware.
//   mode 0: A smem (K-major SW128),  B smem MN-major SW128 (V rows as stored: [k][n], n contiguous)
//   mode 1: A TMEM (tcgen05.st),     B smem MN-major SW128
//   mode 2: A TMEM,                  B smem K-major SW128 (V^T: [n][k], k contiguous, 32-float atoms)
//   mode 3: A smem,                  B smem K-major SW128
#include "occ_common.cuh"
#include "occ_ptx.cuh"

namespace occ {

__global__ void __launch_bounds__(128)
umma_probe_kernel(const float* __restrict__ A /*128x64*/, const float* __restrict__ V /*64x32*/, float* __restrict__ D,
                  int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 2 tiles of 128 rows x 128 B (k 0..31, k 32..63)
  uint8_t* sB = smem + 32768;         // MN-major: 64 rows (k) x 128 B;  K-major: 2 tiles of 32 rows (n) x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<128>(tmem_ptr);
  // stage A: row r, chunk c (16 B) of k-tile t
  for (int i = tid; i < 128 * 16; i += 128) {
    const int r = i / 16, cc = i % 16, t = cc / 8, c = cc % 8;
    *reinterpret_cast<float4*>(sA + t * 16384 + r * 128 + ((c ^ (r & 7)) << 4)) =
        *reinterpret_cast<const float4*>(A + r * 64 + t * 32 + c * 4);
  }
  const bool b_mn = (mode == 0 || mode == 1 || mode >= 4);
  const bool b32 = mode >= 4;  // 4: SS, 5: TS with the SWIZZLE_128B_BASE32B MN-major layout; 6/7: same with LBO=1024
  if (b32) {
    for (int i = tid; i < 64 * 8; i += 128) {
      const int r = i / 8, c = i % 8;  // row = k, 16-byte chunk c; 32-byte chunk (c>>1) swizzled with (r & 3)
      *reinterpret_cast<float4*>(sB + r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4))) =
          *reinterpret_cast<const float4*>(V + r * 32 + c * 4);
    }
  } else if (b_mn) {
    for (int i = tid; i < 64 * 8; i += 128) {
      const int r = i / 8, c = i % 8;  // row = k, 32 n contiguous
      *reinterpret_cast<float4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const float4*>(V + r * 32 + c * 4);
    }
  } else {
    for (int i = tid; i < 32 * 64; i += 128) {
      const int n = i / 64, k = i % 64, t = k / 32, kk = k % 32;
      const int c = kk / 4, w = kk % 4;
      *reinterpret_cast<float*>(sB + t * 4096 + n * 128 + ((c ^ (n & 7)) << 4) + w * 4) = V[k * 32 + n];
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const bool a_tmem = (mode == 1 || mode == 2 || mode == 5 || mode == 7);
  if (a_tmem) {
    // thread i = row i: write its 64 A values to TMEM columns 32..95
    uint32_t r0[32], r1[32];
    for (int j = 0; j < 32; ++j) { r0[j] = __float_as_uint(A[tid * 64 + j]); r1[j] = __float_as_uint(A[tid * 64 + 32 + j]); }
    const uint32_t base = tmem_base + ((uint32_t)(warp * 32) << 16) + 32;
    tmem_st_32x32(base, r0);
    tmem_st_32x32(base + 32, r1);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, 32, 0, b_mn ? 1 : 0);
    for (int kk = 0; kk < 8; ++kk) {
      uint64_t bdesc;
      if (b32) bdesc = make_sw128b32_mn_desc(smem_u32(sB) + kk * 1024, 512, mode >= 6 ? 1024 : 512);
      else if (b_mn) bdesc = make_sw128_desc(smem_u32(sB) + kk * 1024, 1024, 1024);
      else bdesc = make_sw128_desc(smem_u32(sB) + (kk / 4) * 4096 + (kk % 4) * 32, 1024, 16);
      if (a_tmem) {
        mma_tf32_ts(tmem_base, tmem_base + 32 + kk * 8, bdesc, idesc, kk != 0);
      } else {
        const uint64_t adesc = make_sw128_desc(smem_u32(sA) + (kk / 4) * 16384 + (kk % 4) * 32, 1024, 16);
        mma_tf32_ss(tmem_base, adesc, bdesc, idesc, kk != 0);
      }
    }
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) D[tid * 32 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<128>(tmem_base); }
}

}  // namespace occ

extern "C" int occ_debug_umma_probe(const float* A, const float* V, float* D, int mode, cudaStream_t stream) {
  OCC_REQUIRE(A && V && D && mode >= 0 && mode <= 7);
  const size_t smem = 32768 + 8192 + 64 + 1024;
  static bool configured = false;
  if (!configured) {
    OCC_CUDA(cudaFuncSetAttribute(occ::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  occ::umma_probe_kernel<<<1, 128, smem, stream>>>(A, V, D, mode);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}
