// Masked cross-attention of the Mask2Former-3D decoder on the tcgen05 tensor cores (flash-decoding over key chunks).
//
// Reference: mmcv BaseTransformerLayer 'cross_attn' -> MultiheadAttention -> nn.MultiheadAttention with a bool
// attn_mask, called from projects/mmdet3d_plugin/occformer/mask2former/mask2former_nusc_occ.py
// :657-667 with attn_mask = (adaptive_max_pool3d(mask_pred) .sigmoid() < 0.5) and the all-blocked-row reset (:652-653).
//
// One CTA = (key chunk, head, sample); M = 128 query rows (Q <= 128 real), N = 128 keys per tile, head dim 32.
// Split-bf16 operands, three tensor-core passes per contraction (occ_ptx.cuh): Kp / Vp arrive in S32 (one 128-byte chunk
// per (key, head)), the Q tile is split while it is staged, P is split by the softmax threads.
//   warp 0   : TMA producer -- K tile (K-major, SWIZZLE_128B) and V tile (MN-major operand, N = 64 = hi | lo of the head
//              dim, SWIZZLE_128B) are plain 2-D boxes of the projected key / value matrices
//   warp 1   : MMA issuer   -- S = Q K^T (6 x 128x128x16, Q staged once per CTA), O' = [P_hi; P_lo] [V_hi | V_lo]
//              (16 x 128x64x16, P in TMEM); O = O'[:, :32] + O'[:, 32:] when the tile is merged
//   (mask)   : a small pre-pass (mask_bits_kernel) turns the pooled mask logits (ordered ints, blocked <=> negative)
//              into one "blocked" bit per (key, query), 32 keys per word, once per layer instead of once per head; the
//              softmax threads read their 4 words per tile straight from global (L2-resident, 1 MB at S = 80 000)
//   warps 4-7 / 8-11: two softmax warpgroups (even / odd tiles): two passes over the S row in TMEM (max, then
//              exp2 / sum / P store), PV, then the tile's O row is merged into a running (max, sum, acc[32]) in registers.
// Each warpgroup writes one partial per (chunk, head, query): (m, l, acc[32]) in the natural-log domain, the format
// consumed by cross_merge_kernel (head_ops.cu).
#include "occ_common.cuh"
#include "occ_ptx.cuh"

namespace occ {

constexpr int XT_HD = 32;
constexpr int XT_KEYS = 128;
constexpr int XT_STAGES = 4;
constexpr int XT_TILE = XT_KEYS * XT_HD * 4;           // 16 KB
constexpr int XT_STAGE_BYTES = 2 * XT_TILE;             // K tile + V tile
constexpr int XT_THREADS = 384;  // warps: 0 TMA, 1 MMA, 2 TMEM owner, 3 idle, 4-11 softmax

__global__ void __launch_bounds__(XT_THREADS, 1)
cross_attn_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const float* __restrict__ qh, int koff, int voff, const uint32_t* __restrict__ bits /*(B,NW,Q)*/,
                     const int* __restrict__ row_flag, float* __restrict__ part, int S, int Q, int E, int H,
                     int tiles_per_chunk, int nchunk, int passes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sq = smem + XT_STAGES * XT_STAGE_BYTES;  // Q tile, 16 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sq + XT_TILE);
  uint64_t* empty_bar = full_bar + XT_STAGES;
  uint64_t* s_ready = empty_bar + XT_STAGES;
  uint64_t* p_ready = s_ready + 2;
  uint64_t* o_ready = p_ready + 2;
  uint64_t* o_free = o_ready + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int ntiles_total = (S + XT_KEYS - 1) / XT_KEYS;
  const int tile0 = c * tiles_per_chunk;
  const int n_tiles = min(tiles_per_chunk, ntiles_total - tile0);  // >= 1 by construction of nchunk

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < XT_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // TMA producer (expect_tx)
      mbar_init(&empty_bar[i], 1);  // tcgen05.commit after PV
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_ready[i], 1);
      mbar_init(&o_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  // Q tile of this (sample, head): 128 rows x one S32 chunk (32 head-dim values split into hi | lo), K-major
  // SWIZZLE_128B; rows >= Q are zero.  Thread (r, ch) converts 4 values: hi pair words -> 16-byte chunk ch >> 1 (half
  // ch & 1), lo pair words -> chunk 4 + (ch >> 1).
  for (int i = threadIdx.x; i < 128 * 8; i += XT_THREADS) {
    const int r = i >> 3, ch = i & 7;
    uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
    if (r < Q) {
      const float4 v = *reinterpret_cast<const float4*>(qh + ((size_t)b * Q + r) * E + h * XT_HD + ch * 4);
      split_pair(v.x, v.y, hi.x, lo.x);
      split_pair(v.z, v.w, hi.y, lo.y);
    }
    uint8_t* row = sq + r * 128;
    *reinterpret_cast<uint2*>(row + ((((ch >> 1)) ^ (r & 7)) << 4) + ((ch & 1) << 3)) = hi;
    *reinterpret_cast<uint2*>(row + (((4 + (ch >> 1)) ^ (r & 7)) << 4) + ((ch & 1) << 3)) = lo;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % XT_STAGES;
        mbar_wait(&empty_bar[s], (uint32_t)(((t / XT_STAGES) & 1) ^ 1));
        uint8_t* st = smem + (size_t)s * XT_STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], 2 * XT_TILE);
        const int row0 = b * S + (tile0 + t) * XT_KEYS;
        tma_load_2d(st, &tmK, &full_bar[s], koff + h * XT_HD, row0);
        tma_load_2d(st + XT_TILE, &tmV, &full_bar[s], voff + h * XT_HD, row0);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC_QK = make_idesc_bf16(128, XT_KEYS, 0, 0);
      constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 2 * XT_HD, 0, 1);
      const uint64_t qdesc = make_sw128_desc(smem_u32(sq), 1024, 16);
      int nq = 0, np = 0;
      while (np < n_tiles) {
        if (nq < n_tiles && nq - np < 2 && mbar_test(&full_bar[nq % XT_STAGES], (uint32_t)((nq / XT_STAGES) & 1))) {
          tc_fence_after();
          const uint64_t kdesc = make_sw128_desc(smem_u32(smem + (size_t)(nq % XT_STAGES) * XT_STAGE_BYTES), 1024, 16);
          const uint32_t s_tmem = tmem_base + (nq & 1) * 128;
          mma_bf16x3_ss(s_tmem, qdesc, kdesc, IDESC_QK, 0u, passes);
          mma_commit(&s_ready[nq & 1]);
          ++nq;
        }
        if (np < nq) {
          const int tb = np & 1;
          const uint32_t k = (uint32_t)(np >> 1);
          if (mbar_test(&p_ready[tb], k & 1) && mbar_test(&o_free[tb], (k & 1) ^ 1)) {
            tc_fence_after();
            const int s = np % XT_STAGES;
            const uint64_t vdesc = make_sw128_desc(smem_u32(smem + (size_t)s * XT_STAGE_BYTES + XT_TILE), 1024, 1024);
            const uint32_t p_tmem = tmem_base + tb * 128, o_tmem = tmem_base + 256 + tb * 2 * XT_HD;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {  // 16 keys per MMA = two 8-key (1024 B) atoms of V rows
              const uint32_t pc = p_tmem + (kk >> 1) * 32 + (kk & 1) * 8;
              mma_bf16_ts(o_tmem, pc, vdesc + (uint64_t)(kk * 128), IDESC_PV, kk != 0);  // P_hi [V_hi | V_lo]
              if (passes == 3) mma_bf16_ts(o_tmem, pc + 16, vdesc + (uint64_t)(kk * 128), IDESC_PV, 1u);  // P_lo [V_hi | V_lo]
            }
            mma_commit(&o_ready[tb]);
            mma_commit(&empty_bar[s]);
            ++np;
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================================================== softmax / merge warpgroups
    const int wg = (warp - 4) >> 2;
    const int i = ((warp & 3) << 5) + lane;  // query row = TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const bool use_mask = (i < Q) && row_flag[b * Q + i] != 0;
    const int NW = ((S + XT_KEYS - 1) / XT_KEYS) * 4;
    const int tb = wg;
    const float L2E = 1.4426950408889634f;
    float M = -INFINITY, L = 0.f, acc[XT_HD];
#pragma unroll
    for (int d = 0; d < XT_HD; ++d) acc[d] = 0.f;
    for (int t = wg; t < n_tiles; t += 2) {
      const uint32_t k = (uint32_t)(t >> 1);
      mbar_wait(&s_ready[tb], k & 1);
      tc_fence_after();
      uint32_t blk[4];
      {
        const int key_base = (tile0 + t) * XT_KEYS;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int nvalid = min(max(S - (key_base + 32 * w), 0), 32);
          const uint32_t tailw = nvalid >= 32 ? 0u : (0xFFFFFFFFu << nvalid);  // keys beyond S are always blocked
          blk[w] = use_mask ? __ldg(bits + ((size_t)b * NW + (size_t)(tile0 + t) * 4 + w) * Q + i) : tailw;
        }
      }
      const uint32_t s_col = lane_base + tb * 128;
      uint32_t r[32];
      float m = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {  // pass A: row maximum over the un-blocked keys
        tmem_ld_32x32(s_col + cc * 32, r);
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (!((blk[cc] >> j) & 1u)) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[j]));
        m = fmaxf(m, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      }
      const float ml2 = m * L2E;
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {  // pass B: P = exp(s - m), row sum, P -> TMEM (in place of S), split into hi | lo
        tmem_ld_32x32(s_col + cc * 32, r);
        tmem_ld_wait();
        const uint32_t bw = cc == 0 ? blk[0] : cc == 1 ? blk[1] : cc == 2 ? blk[2] : blk[3];
        uint32_t lo16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {  // keys 2j, 2j+1 -> packed hi word j (in place: r[j] was consumed by pair j/2) + lo word
          float p0 = 0.f, p1 = 0.f;
          if (!((bw >> (2 * j)) & 1u)) p0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), L2E, -ml2));
          if (!((bw >> (2 * j + 1)) & 1u)) p1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), L2E, -ml2));
          l4[j & 3] += p0 + p1;
          uint32_t hi, lo;
          split_pair(p0, p1, hi, lo);
          r[j] = hi;
          lo16[j] = lo;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) r[16 + j] = lo16[j];
        tmem_st_32x32(s_col + cc * 32, r);
      }
      const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[tb]);
      mbar_wait(&o_ready[tb], k & 1);
      tc_fence_after();
      // merge the tile into the running softmax state of this row: O = O'[:, :32] + O'[:, 32:] (the V_hi and V_lo halves),
      // one half at a time so that only 32 accumulator words are live next to acc[]
      const bool upd = m > -INFINITY;
      const float Mn = fmaxf(M, m);
      const float a = upd ? ex2_approx((M - Mn) * L2E) : 1.f, bsc = upd ? ex2_approx((m - Mn) * L2E) : 0.f;
      tmem_ld_32x32(lane_base + 256 + tb * 2 * XT_HD, r);  // P [V_hi]
      tmem_ld_wait();
      if (upd) {
#pragma unroll
        for (int d = 0; d < XT_HD; ++d) acc[d] = acc[d] * a + __uint_as_float(r[d]) * bsc;
      }
      tmem_ld_32x32(lane_base + 256 + tb * 2 * XT_HD + XT_HD, r);  // P [V_lo]
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[tb]);
      if (upd) {
#pragma unroll
        for (int d = 0; d < XT_HD; ++d) acc[d] = fmaf(__uint_as_float(r[d]), bsc, acc[d]);
        L = L * a + l * bsc;
        M = Mn;
      }
    }
    if (i < Q) {
      float* dst = part + ((((size_t)b * H + h) * (2 * nchunk) + 2 * c + wg) * Q + i) * (XT_HD + 2);
      dst[0] = M;
      dst[1] = L;
#pragma unroll
      for (int d = 0; d < XT_HD; ++d) dst[2 + d] = acc[d];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// pooled (B,S,Q) ordered ints -> bits (B, NW, Q): bit k of word w = key 32w+k is blocked for query q (negative
// pooled logit, or key >= S).  One thread per (word, query): 32 independent coalesced loads.
__global__ void __launch_bounds__(128)
mask_bits_kernel(const int* __restrict__ pooled, uint32_t* __restrict__ bits, int S, int Q, int NW) {
  const int q = threadIdx.x, w = blockIdx.x, b = blockIdx.y;
  if (q >= Q) return;
  uint32_t word = 0;
  int vals[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int key = 32 * w + k;
    vals[k] = key < S ? __ldg(pooled + ((size_t)b * S + key) * Q + q) : -1;
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) word |= (uint32_t)(vals[k] < 0) << k;
  bits[((size_t)b * NW + w) * Q + q] = word;
}

}  // namespace occ

using namespace occ;

extern "C" int occ_mask_bits(const int* pooled, unsigned* bits, int B, int S, int Q, cudaStream_t stream) {
  OCC_REQUIRE(pooled && bits && B > 0 && S > 0 && Q > 0 && Q <= 128 && B <= 65535);
  const int NW = ((S + XT_KEYS - 1) / XT_KEYS) * 4;
  mask_bits_kernel<<<dim3(NW, B), 128, 0, stream>>>(pooled, bits, S, Q, NW);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}

// number of partials per (sample, head) that occ_cross_attn_tc writes for S keys (= 2 * key chunks)
extern "C" int occ_cross_attn_tc_partials(int S) {
  const int tiles = (S + XT_KEYS - 1) / XT_KEYS;
  const int tpc = (tiles + 47) / 48;  // <= 48 chunks -> <= 96 partials
  return 2 * ((tiles + tpc - 1) / tpc);
}

// qh (B,Q,E) scaled projected queries (fp32); Kp / Vp (B*S, ld) projected keys / values in S32, this layer's slice at
// column koff / voff; bits (B, NW = 4*ceil(S/128), Q) blocked-key words from occ_mask_bits; row_flag (B*Q);
// part (B, H, npart, Q, 34).
extern "C" int occ_cross_attn_tc(const float* qh, const float* Kp, const float* Vp, int ld, int koff, int voff,
                                 const unsigned* bits, const int* row_flag, float* part, int B, int S, int Q, int E,
                                 int H, cudaStream_t stream) {
  OCC_REQUIRE(qh && Kp && Vp && bits && row_flag && part);
  OCC_REQUIRE(B > 0 && S > 0 && Q > 0 && Q <= 128 && Q % 4 == 0 && H > 0 && E == H * XT_HD && B <= 65535 && H <= 65535);
  OCC_REQUIRE(ld % 4 == 0 && koff % 4 == 0 && voff % 4 == 0 && koff + E <= ld && voff + E <= ld);
  OCC_REQUIRE((reinterpret_cast<uintptr_t>(qh) & 15) == 0 && (reinterpret_cast<uintptr_t>(Kp) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(Vp) & 15) == 0);
  const int tiles = (S + XT_KEYS - 1) / XT_KEYS;
  const int tpc = (tiles + 47) / 48;
  const int nchunk = (tiles + tpc - 1) / tpc;
  CUtensorMap tmK, tmV;
  uint64_t dims[2] = {(uint64_t)ld, (uint64_t)B * S};
  uint64_t strides[1] = {(uint64_t)ld * 4};
  uint32_t box[2] = {(uint32_t)XT_HD, (uint32_t)XT_KEYS};
  int rc = make_tmap_f32(&tmK, Kp, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_f32(&tmV, Vp, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const size_t smem = (size_t)XT_STAGES * XT_STAGE_BYTES + XT_TILE + 1024 /*align*/ + 256 /*barriers*/;
  OCC_ENSURE_SMEM(cross_attn_tc_kernel, smem);
  dim3 grid(nchunk, H, B);
  cross_attn_tc_kernel<<<grid, XT_THREADS, smem, stream>>>(tmK, tmV, qh, koff, voff, bits, row_flag, part, S, Q, E, H, tpc,
                                                           nchunk, mma_passes());
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}
