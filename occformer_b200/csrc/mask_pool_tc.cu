// Mask einsum + adaptive max pool of the Mask2Former-3D decoder, "query-stationary" on the tcgen05 tensor cores.
//
// Reference: projects/mmdet3d_plugin/occformer/mask2former/mask2former_nusc_occ.py:457 (mask_pred = einsum('bqc,bcxyz->
// bqxyz', mask_embed, mask_features)), :463-466 (attn_mask = adaptive_max_pool3d(mask_pred, level size).sigmoid() < .5)
// and :652-653 (rows that block every key attend everywhere).  The (B, Q, X, Y, Z) logits of the intermediate layers
// are only ever consumed through the pooled mask, so this kernel never writes them: 491 MB of mask features are read
// once per layer and ~1 MB of pooled logits is written.
//
// Both operands arrive in the S32 split format; every 32-k block is three bf16 tensor-core passes (occ_ptx.cuh).
// D[128 queries x 128 voxels] = membed[b] (A, resident in smem, K-major) x mf tile^T (B operand: the voxel rows of a
// (bx, by, bz) box, K-major, TMA 5-D box loads).  With the queries on the TMEM lanes every epilogue thread owns one
// query and sees the 128 voxels of the tile as its own registers: the window maximum is a register reduction with
// compile-time indices (no shuffles, no shared memory, no atomics when the box holds whole pooling cells), and a
// cell's Q maxima are written by adjacent threads (coalesced).
//   warp 0: TMA producer (ring of 16 KB k-blocks)   warp 1: MMA issuer   warp 2: TMEM owner (4 accumulators)
//   warps 4-7 / 8-11: two epilogue warpgroups (even / odd tiles)
#include "occ_common.cuh"
#include "occ_ptx.cuh"

namespace occ {

constexpr int MP_THREADS = 384;
constexpr int MP_KB_BYTES = 128 * 32 * 4;  // one k-block of A or B: 128 rows x 32 floats

struct MaskPoolParams {
  const float* membed;  // (B, Q, E) S32
  int* pooled;          // (B, Xo*Yo*Zo, Q) ordered ints
  int* flag;            // (B, Q)
  int Q, E, KB, stages;
  int tiles_y, tiles_z, n_tiles;  // tiles of one sample
  int Xo, Yo, Zo;
  int passes;  // tensor-core passes per 32-k block (occ_common.cuh: mma_passes)
};

__device__ __forceinline__ int ordered_int(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}

// P: cubic pooling window; (BX, BY, BZ): voxel box of a tile (BX*BY*BZ == 128, z fastest = B-operand row order)
template <int P, int BX, int BY, int BZ>
__global__ void __launch_bounds__(MP_THREADS, 1)
mask_pool_tc_kernel(const __grid_constant__ CUtensorMap tmB, const MaskPoolParams p) {
  static_assert(BX * BY * BZ == 128, "tile = 128 voxels");
  constexpr int CWX = P < BX ? P : BX, CWY = P < BY ? P : BY, CWZ = P < BZ ? P : BZ;  // cell extent inside the box
  constexpr int NCX = BX / CWX, NCY = BY / CWY, NCZ = BZ / CWZ, NC = NCX * NCY * NCZ;
  constexpr bool COMPLETE = CWX == P && CWY == P && CWZ == P;
  static_assert(NC <= 16, "cell maxima live in registers");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sa = smem;                                   // A: KB k-blocks
  uint8_t* ring = sa + (size_t)p.KB * MP_KB_BYTES;       // B ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * MP_KB_BYTES);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* acc_full = empty_bar + 8;
  uint64_t* acc_empty = acc_full + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmB);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  {  // A = membed[b]: 128 rows (rows >= Q zero) x E, K-major SWIZZLE_128B k-blocks
    const int E4 = p.E >> 2;
    const float* src = p.membed + (size_t)b * p.Q * p.E;
    for (int i = threadIdx.x; i < 128 * E4; i += MP_THREADS) {
      const int r = i / E4, c4 = i - r * E4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < p.Q) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * p.E) + c4);  // S32 words, copied verbatim
      const int kb = c4 >> 3, ch = c4 & 7;
      *reinterpret_cast<float4*>(sa + (size_t)kb * MP_KB_BYTES + r * 128 + ((ch ^ (r & 7)) << 4)) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int it = 0;
      for (int i = 0; i < n_my; ++i) {
        const int t = blockIdx.x + i * gridDim.x;
        const int tz = t % p.tiles_z, ty = (t / p.tiles_z) % p.tiles_y, tx = t / (p.tiles_z * p.tiles_y);
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % p.stages;
          mbar_wait(&empty_bar[s], (uint32_t)(((it / p.stages) & 1) ^ 1));
          mbar_expect_tx(&full_bar[s], MP_KB_BYTES);
          tma_load_5d(ring + (size_t)s * MP_KB_BYTES, &tmB, &full_bar[s], kb * 32, tz * BZ, ty * BY, tx * BX, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc_bf16(128, 128, 0, 0);
      int it = 0;
      for (int i = 0; i < n_my; ++i) {
        const int buf = i & 3;
        mbar_wait(&acc_empty[buf], (uint32_t)(((i >> 2) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 128;
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % p.stages;
          mbar_wait(&full_bar[s], (uint32_t)((it / p.stages) & 1));
          tc_fence_after();
          const uint64_t adesc = make_sw128_desc(smem_u32(sa + (size_t)kb * MP_KB_BYTES), 1024, 16);
          const uint64_t bdesc = make_sw128_desc(smem_u32(ring + (size_t)s * MP_KB_BYTES), 1024, 16);
          mma_bf16x3_ss(d_tmem, adesc, bdesc, IDESC, kb != 0, p.passes);
          mma_commit(&empty_bar[s]);
        }
        mma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue warpgroups
    const int wg = (warp - 4) >> 2;
    const int q = ((warp & 3) << 5) + lane;  // query = TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    bool anypos = false;
    int* pooled_b = p.pooled + (size_t)b * p.Xo * p.Yo * p.Zo * p.Q;
    for (int i = wg; i < n_my; i += 2) {
      const int buf = i & 3;
      const int t = blockIdx.x + i * gridDim.x;
      const int tz = t % p.tiles_z, ty = (t / p.tiles_z) % p.tiles_y, tx = t / (p.tiles_z * p.tiles_y);
      mbar_wait(&acc_full[buf], (uint32_t)((i >> 2) & 1));
      tc_fence_after();
      float m[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) m[c] = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32(lane_base + buf * 128 + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = cc * 32 + j;  // voxel of the box, z fastest
          const int x = n / (BY * BZ), y = (n / BZ) % BY, z = n % BZ;
          const int cell = ((x / CWX) * NCY + y / CWY) * NCZ + z / CWZ;
          m[cell] = fmaxf(m[cell], __uint_as_float(r[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (q < p.Q) {
        const int gx0 = tx * BX / P, gy0 = ty * BY / P, gz0 = tz * BZ / P;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int cx = c / (NCY * NCZ), cy = (c / NCZ) % NCY, cz = c % NCZ;
          int* dst = pooled_b + ((size_t)((gx0 + cx) * p.Yo + gy0 + cy) * p.Zo + gz0 + cz) * p.Q + q;
          const int o = ordered_int(m[c]);
          anypos |= o >= 0;
          if (COMPLETE) *dst = o;
          else atomicMax(dst, o);
        }
      }
    }
    if (q < p.Q && anypos) p.flag[(size_t)b * p.Q + q] = 1;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int P, int BX, int BY, int BZ>
static int launch_mask_pool(const float* mf, const MaskPoolParams& p0, int B, int X, int Y, int Z, cudaStream_t stream) {
  MaskPoolParams p = p0;
  p.tiles_y = Y / BY;
  p.tiles_z = Z / BZ;
  p.n_tiles = (X / BX) * p.tiles_y * p.tiles_z;
  CUtensorMap tmB;
  uint64_t dims[5] = {(uint64_t)p.E, (uint64_t)Z, (uint64_t)Y, (uint64_t)X, (uint64_t)B};
  uint64_t strides[4] = {(uint64_t)p.E * 4, (uint64_t)Z * p.E * 4, (uint64_t)Y * Z * p.E * 4, (uint64_t)X * Y * Z * p.E * 4};
  uint32_t box[5] = {32u, (uint32_t)BZ, (uint32_t)BY, (uint32_t)BX, 1u};
  int rc = make_tmap_f32(&tmB, mf, 5, dims, strides, box, nullptr);
  if (rc) return rc;
  const size_t smem = (size_t)(p.KB + p.stages) * MP_KB_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  auto kern = mask_pool_tc_kernel<P, BX, BY, BZ>;
  OCC_ENSURE_SMEM(kern, smem);
  int gx = sm_count() / B;
  if (gx < 1) gx = 1;
  if (gx > p.n_tiles) gx = p.n_tiles;
  kern<<<dim3(gx, B), MP_THREADS, smem, stream>>>(tmB, p);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}

// Returns OCC_OK when the query-stationary kernel handled the call, 1 when the shape is not covered (the caller falls
// back to the voxel-stationary epilogue pooling in gemm_bf16x3.cu), < 0 / cudaError on failure.
int mask_pool_query_stationary(const float* mf, const float* membed, int* pooled, int* flag, int B, int X, int Y, int Z,
                               int E, int Q, int Xo, int Yo, int Zo, cudaStream_t stream) {
  const int wx = X / Xo, wy = Y / Yo, wz = Z / Zo;
  if (wx != wy || wy != wz || E % 32 != 0 || Q > 128 || B > 65535) return 1;
  MaskPoolParams p{};
  p.passes = mma_passes();
  p.membed = membed; p.pooled = pooled; p.flag = flag;
  p.Q = Q; p.E = E; p.KB = E / 32; p.Xo = Xo; p.Yo = Yo; p.Zo = Zo;
  const int budget = 227 * 1024 - 1024 - 256 - p.KB * MP_KB_BYTES;
  p.stages = budget / MP_KB_BYTES;
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 3) return 1;
  OCC_CUDA(cudaMemsetAsync(flag, 0, (size_t)B * Q * sizeof(int), stream));
  if (wx == 2 && X % 2 == 0 && Y % 4 == 0 && Z % 16 == 0) return launch_mask_pool<2, 2, 4, 16>(mf, p, B, X, Y, Z, stream);
  if (wx == 4 && X % 4 == 0 && Y % 4 == 0 && Z % 8 == 0) return launch_mask_pool<4, 4, 4, 8>(mf, p, B, X, Y, Z, stream);
  if (wx == 8 && X % 8 == 0 && Y % 8 == 0 && Z % 8 == 0) {
    // a box holds a quarter of a cell: the partial maxima meet through atomicMax on the ordered ints
    OCC_CUDA(cudaMemsetAsync(pooled, 0x80, (size_t)B * Xo * Yo * Zo * Q * sizeof(int), stream));
    return launch_mask_pool<8, 8, 2, 8>(mf, p, B, X, Y, Z, stream);
  }
  return 1;
}

}  // namespace occ
