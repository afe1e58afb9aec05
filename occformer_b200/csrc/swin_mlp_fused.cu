// Swin block tail for C = 128 tokens in one kernel: attention output projection + residual, LayerNorm2, FFN
// (Linear - GELU - Linear) + residual.  Three chained 128x128x128 GEMMs per 128-token tile on tcgen05 (split-bf16
// operands, three passes = fp32-faithful, occ_ptx.cuh); the intermediate activations never leave the SM (TMEM / registers).
//
// Reference: mmdet SwinBlock.forward (swin.py) as used by projects/mmdet3d_plugin/occformer/backbones/
// dualpath_block.py:57-70 (shared_transformer on every Z slice and on the BEV slice):
//     x = identity + attn.proj(att)          (window_attention.py WindowMSA.proj, ShiftWindowMSA residual)
//     x = x + ffn(norm2(x))                  (FFN: Linear(C, C) -> GELU(erf) -> Linear(C, C))
// Unfused this is proj GEMM (R att, R tok, W y1) + LayerNorm (R y1, W y1n) + FFN1 (R y1n, W h) + FFN2 (R h, R y1, W y2):
// 2.4 GB of HBM traffic for the 680k tokens of a 200x200x16 grid; fused: R att, R tok, W y2 = 1.04 GB.
//
//   warp 0  : producer -- 'att' k-blocks (HBM) into a 4-slot ring, weight k-blocks (L2) into a 4-slot ring, both in the
//             exact order the MMA warp consumes them (non-blocking polling, so neither ring stalls the other)
//   warp 1  : MMA issuer + TMEM owner, static order  M1(0) | M2(i) M1(i+1) M3(i) | ...  so that the next tile's
//             projection GEMM runs under the current tile's GELU epilogue.  TMEM (4 x 128 columns): T0 / T2 = Racc of
//             even / odd tiles (D1, then y1 + b2 written back by epilogue 1, then M3 accumulates the FFN output ONTO it:
//             the residual add is done by the tensor core and y1 never has to stay in registers), T1 = A operand of the
//             next GEMM (LN2(y1), then GELU(h)), T3 = D2.
//   warps 2-17: 16 epilogue warps; thread = (token row, 32-column chunk).  The LayerNorm row statistics of the four
//             chunks meet in shared memory (two-pass: mean, then centred sum of squares).  The identity rows of the
//             next tile are prefetched into registers while the current tile waits for its GEMMs.
#include "occ_common.cuh"
#include "occ_ptx.cuh"

namespace occ {

constexpr int SMF_THREADS = 576;
constexpr int SMF_EPI_THREADS = 512;
constexpr int SMF_C = 128;
constexpr int SMF_SLOT = 128 * 32 * 4;  // 16 KB k-block: 128 rows x 32 floats, K-major SWIZZLE_128B
constexpr int SMF_ATT_SLOTS = 4;
constexpr int SMF_W_SLOTS = 4;
constexpr float SMF_EPS = 1e-5f;

struct SwinMlpParams {
  const float* tok;  // (M, 128) identity of the attention residual
  const float *bp, *lnw, *lnb, *b1, *b2;
  float* out;        // (M, 128)
  long long M;
  int n_tiles;
  int passes;  // tensor-core passes per 32-k block (occ_common.cuh: mma_passes)
};

// GELU(x) = x * Phi(x) with erfc(z) = P(t) exp(-z^2), t = 1 / (1 + p z) (Abramowitz-Stegun 7.1.26, |err| <= 1.5e-7):
// branch-free, 2 MUFU + ~14 FP32 ops instead of erff's ~30.
__device__ __forceinline__ float smf_gelu(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ex2_approx(-1.4426950408889634f * z * z);
  const float half_q = 0.5f * poly * t * e;               // 0.5 * erfc(z)
  const float phi = 0.5f + copysignf(0.5f - half_q, x);   // x >= 0: 1 - q/2 ; x < 0: q/2
  return x * phi;
}

__global__ void __launch_bounds__(SMF_THREADS, 1)
swin_mlp_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWp,
                      const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                      const __grid_constant__ CUtensorMap tmOut, const SwinMlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* att_ring = smem;
  uint8_t* w_ring = att_ring + SMF_ATT_SLOTS * SMF_SLOT;
  uint8_t* out_stage = w_ring + SMF_W_SLOTS * SMF_SLOT;  // 4 k-blocks (one per column chunk), SWIZZLE_128B, TMA-stored
  float* vecs = reinterpret_cast<float*>(out_stage + 4 * SMF_SLOT);  // bp | lnw | lnb | b1 | b2
  float* stat = vecs + 5 * SMF_C;                                           // [2][4][128] row partials (sum, centred sq)
  uint64_t* att_full = reinterpret_cast<uint64_t*>(stat + 2 * 4 * 128);
  uint64_t* att_empty = att_full + SMF_ATT_SLOTS;
  uint64_t* w_full = att_empty + SMF_ATT_SLOTS;
  uint64_t* w_empty = w_full + SMF_W_SLOTS;
  uint64_t* d1_ready = w_empty + SMF_W_SLOTS;  // [2] M1 of tile slot g complete
  uint64_t* d23_ready = d1_ready + 2;          // M2 / M3 complete (in tile order)
  uint64_t* a_ready = d23_ready + 1;           // epilogue -> MMA: the A operand in R1 is ready (16 warps arrive)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmWp);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < SMF_ATT_SLOTS; ++i) {
      mbar_init(&att_full[i], 1);
      mbar_init(&att_empty[i], 1);
    }
    for (int i = 0; i < SMF_W_SLOTS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(&d1_ready[0], 1);
    mbar_init(&d1_ready[1], 1);
    mbar_init(d23_ready, 1);
    mbar_init(a_ready, SMF_EPI_THREADS / 32);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  for (int i = threadIdx.x; i < 5 * SMF_C; i += SMF_THREADS) {
    const int which = i / SMF_C, c = i % SMF_C;
    const float* src = which == 0 ? p.bp : which == 1 ? p.lnw : which == 2 ? p.lnb : which == 3 ? p.b1 : p.b2;
    vecs[i] = __ldg(src + c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================================== producer (two rings, consumption order)
    if (lane == 0) {
      // att stream: tile i, 4 k-blocks each.  weight stream (4 k-blocks per matrix): Wp(0), then per tile i:
      // W1(i), Wp(i+1) if there is a next tile, W2(i).
      const int att_total = n_my * 4;
      const int w_total = 12 * n_my;
      int ia = 0, iw = 0;
      while (ia < att_total || iw < w_total) {
        if (ia < att_total) {
          const int s = ia % SMF_ATT_SLOTS;
          if (mbar_test(&att_empty[s], (uint32_t)(((ia / SMF_ATT_SLOTS) & 1) ^ 1))) {
            const int tile = blockIdx.x + (ia >> 2) * gridDim.x;
            mbar_expect_tx(&att_full[s], SMF_SLOT);
            tma_load_2d(att_ring + (size_t)s * SMF_SLOT, &tmA, &att_full[s], (ia & 3) * 32, tile * 128);
            ++ia;
          }
        }
        if (iw < w_total) {
          const int s = iw % SMF_W_SLOTS;
          if (mbar_test(&w_empty[s], (uint32_t)(((iw / SMF_W_SLOTS) & 1) ^ 1))) {
            // matrix blocks in stream order: 0: Wp(0); then for tile i: 1+3i: W1, 2+3i: Wp (next tile) or -- for the last
            // tile, which has no successor -- W2; 3+3i: W2 (absent for the last tile)
            const int blk = iw >> 2;
            int mat;  // 0 Wp, 1 W1, 2 W2
            if (blk == 0) mat = 0;
            else {
              const int i = (blk - 1) / 3, r = (blk - 1) % 3;
              const bool last = i == n_my - 1;
              mat = r == 0 ? 1 : (r == 1 ? (last ? 2 : 0) : 2);
            }
            const CUtensorMap* tm = mat == 0 ? &tmWp : mat == 1 ? &tmW1 : &tmW2;
            mbar_expect_tx(&w_full[s], SMF_SLOT);
            tma_load_2d(w_ring + (size_t)s * SMF_SLOT, tm, &w_full[s], (iw & 3) * 32, 0);
            ++iw;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = make_idesc_bf16(128, 128, 0, 0);
      int ia = 0, iw = 0;
      uint32_t na = 0;  // waits consumed on a_ready
      auto issue_m1 = [&](int i) {  // D1 = att x Wp^T into Racc of tile slot i & 1
        const uint32_t d_tmem = tmem_base + (i & 1) * 256;
        for (int kb = 0; kb < 4; ++kb, ++ia, ++iw) {
          const int sa = ia % SMF_ATT_SLOTS, sw = iw % SMF_W_SLOTS;
          mbar_wait(&att_full[sa], (uint32_t)((ia / SMF_ATT_SLOTS) & 1));
          mbar_wait(&w_full[sw], (uint32_t)((iw / SMF_W_SLOTS) & 1));
          tc_fence_after();
          const uint64_t adesc = make_sw128_desc(smem_u32(att_ring + (size_t)sa * SMF_SLOT), 1024, 16);
          const uint64_t bdesc = make_sw128_desc(smem_u32(w_ring + (size_t)sw * SMF_SLOT), 1024, 16);
          mma_bf16x3_ss(d_tmem, adesc, bdesc, IDESC, kb != 0, p.passes);
          mma_commit(&att_empty[sa]);
          mma_commit(&w_empty[sw]);
        }
        mma_commit(&d1_ready[i & 1]);
      };
      auto issue_m23 = [&](int i, bool third) {  // M2: T3 = T1 x W1^T ;  M3: Racc += T1 x W2^T  (A operand from TMEM)
        mbar_wait(a_ready, na & 1u);
        ++na;
        tc_fence_after();
        const uint32_t d_tmem = third ? tmem_base + (i & 1) * 256 : tmem_base + 384, a_tmem = tmem_base + 128;
        for (int kb = 0; kb < 4; ++kb, ++iw) {
          const int sw = iw % SMF_W_SLOTS;
          mbar_wait(&w_full[sw], (uint32_t)((iw / SMF_W_SLOTS) & 1));
          tc_fence_after();
          const uint64_t bdesc = make_sw128_desc(smem_u32(w_ring + (size_t)sw * SMF_SLOT), 1024, 16);
          // A block kb in TMEM: 32 columns = [16 packed hi | 16 packed lo] of k = 32 kb .. 32 kb + 31
          mma_bf16x3_ts(d_tmem, a_tmem + kb * 32, bdesc, IDESC, (third || kb != 0) ? 1u : 0u, p.passes);
          mma_commit(&w_empty[sw]);
        }
        mma_commit(d23_ready);
      };
      issue_m1(0);
      for (int i = 0; i < n_my; ++i) {
        issue_m23(i, false);  // M2(i) after epilogue 1 of tile i
        // Racc of the other tile slot was last read by epilogue 3 of tile i-1, which every epilogue warp finished
        // before it arrived for epilogue 1 of tile i
        if (i + 1 < n_my) issue_m1(i + 1);
        issue_m23(i, true);  // M3(i) after epilogue 2 of tile i
      }
    }
  } else {
    // ===================================================================== epilogue: thread = (row q, 32-column chunk cc)
    const int e = warp - 2;
    const int cc = e >> 2;
    const int q = ((warp & 3) << 5) + lane;  // TMEM lane (a warp may only touch the lanes of quadrant warp % 4)
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + cc * 32;
    const float *s_bp = vecs + cc * 32, *s_lnw = vecs + SMF_C + cc * 32, *s_lnb = vecs + 2 * SMF_C + cc * 32,
                *s_b1 = vecs + 3 * SMF_C + cc * 32, *s_b2 = vecs + 4 * SMF_C + cc * 32;
    float* st_sum = stat;
    float* st_sq = stat + 4 * 128;
    uint32_t n23 = 0;
    const uint32_t t1 = lane_base + 128, t3 = lane_base + 384;
    auto load_tok = [&](int i, float (&t)[32]) {
      const long long row = (long long)(blockIdx.x + i * gridDim.x) * 128 + q;
      const float4* tp = reinterpret_cast<const float4*>(p.tok + (row < p.M ? row : 0) * SMF_C + cc * 32);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 v = __ldg(tp + c4);
        t[4 * c4 + 0] = v.x; t[4 * c4 + 1] = v.y; t[4 * c4 + 2] = v.z; t[4 * c4 + 3] = v.w;
      }
    };
    const bool issuer = (e & 3) == 0 && lane == 0;  // one thread per column chunk issues that chunk's TMA stores
    uint8_t* stage_row = out_stage + (size_t)cc * SMF_SLOT + q * 128;
    float y[32];
    load_tok(0, y);
    for (int i = 0; i < n_my; ++i) {
      const uint32_t racc = lane_base + (i & 1) * 256;
      const long long row = (long long)(blockIdx.x + i * gridDim.x) * 128 + q;
      uint32_t r[32];
      // ---- epilogue 1: y1 = tok + D1 + bp ; LN2(y1) -> T1 ; y1 + b2 -> Racc (M3 accumulates onto it)
      mbar_wait(&d1_ready[i & 1], (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      tmem_ld_32x32(racc, r);
      tmem_ld_wait();
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        y[j] += __uint_as_float(r[j]) + s_bp[j];
        sum += y[j];
      }
      st_sum[cc * 128 + q] = sum;
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(y[j] + s_b2[j]);
      tmem_st_32x32(racc, r);
      named_bar_sync(1, SMF_EPI_THREADS);
      const float mean = ((st_sum[q] + st_sum[128 + q]) + (st_sum[256 + q] + st_sum[384 + q])) * (1.0f / SMF_C);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        y[j] -= mean;
        sq += y[j] * y[j];
      }
      st_sq[cc * 128 + q] = sq;
      named_bar_sync(1, SMF_EPI_THREADS);
      const float var = ((st_sq[q] + st_sq[128 + q]) + (st_sq[256 + q] + st_sq[384 + q])) * (1.0f / SMF_C);
      const float rstd = rsqrtf(var + SMF_EPS);
      {
        float a[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = y[j] * rstd * s_lnw[j] + s_lnb[j];
        split_chunk32(a, r);  // this thread's 32 k values of the next GEMM's A operand: 16 hi + 16 lo packed columns
      }
      tmem_st_32x32(t1, r);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      if (i + 1 < n_my) load_tok(i + 1, y);  // in flight while this tile waits for M2 / M3
      // ---- epilogue 2: h = GELU(D2 + b1) -> T1
      mbar_wait(d23_ready, n23 & 1u);
      ++n23;
      tc_fence_after();
      tmem_ld_32x32(t3, r);
      tmem_ld_wait();
      {
        float a[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = smf_gelu(__uint_as_float(r[j]) + s_b1[j]);
        split_chunk32(a, r);
      }
      tmem_st_32x32(t1, r);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      // ---- epilogue 3: Racc = y1 + b2 + h W2^T
      mbar_wait(d23_ready, n23 & 1u);
      ++n23;
      tc_fence_after();
      tmem_ld_32x32(racc, r);
      tmem_ld_wait();
      // rows go through a swizzled staging k-block and one TMA store per chunk: a direct store would touch 32 different
      // 128-byte lines per instruction (thread = row)
      if (issuer) tma_store_wait_read<0>();  // the previous tile's store has finished reading the staging block
      named_bar_sync(2 + cc, 128);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        *reinterpret_cast<float4*>(stage_row + ((c4 ^ (q & 7)) << 4)) =
            make_float4(__uint_as_float(r[4 * c4 + 0]), __uint_as_float(r[4 * c4 + 1]), __uint_as_float(r[4 * c4 + 2]),
                        __uint_as_float(r[4 * c4 + 3]));
      fence_proxy_async_smem();
      named_bar_sync(2 + cc, 128);
      if (issuer) {
        tma_store_2d(&tmOut, out_stage + (size_t)cc * SMF_SLOT, cc * 32, (int)(row - q));  // rows >= M are clipped
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace occ

using namespace occ;

// out[M,128] = y1 + W2 gelu(W1 LN(y1) + b1) + b2,  y1 = tok + att Wp^T + bp.   att, Wp, W1, W2 in S32; weights
// (out, in) row-major as in nn.Linear.
extern "C" int occ_swin_proj_ffn(const float* att, const float* tok, const float* wp, const float* bp, const float* ln_w,
                                 const float* ln_b, const float* w1, const float* b1, const float* w2, const float* b2,
                                 float* out, long long M, int C, cudaStream_t stream) {
  OCC_REQUIRE(att && tok && wp && bp && ln_w && ln_b && w1 && b1 && w2 && b2 && out);
  OCC_REQUIRE(M > 0 && M < (1ll << 31) - 128);
  if (C != SMF_C) return OCC_EUNSUPPORTED;
  OCC_REQUIRE((reinterpret_cast<uintptr_t>(att) & 15) == 0 && (reinterpret_cast<uintptr_t>(tok) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(wp) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(w1) & 15) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0);
  CUtensorMap tmA, tmWp, tmW1, tmW2;
  uint32_t box[2] = {32u, 128u};
  uint64_t strides[1] = {(uint64_t)SMF_C * 4};
  {
    uint64_t dims[2] = {(uint64_t)SMF_C, (uint64_t)M};
    int rc = make_tmap_f32(&tmA, att, 2, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  uint64_t wdims[2] = {(uint64_t)SMF_C, (uint64_t)SMF_C};
  int rc = make_tmap_f32(&tmWp, wp, 2, wdims, strides, box, nullptr);
  if (rc) return rc;
  rc = make_tmap_f32(&tmW1, w1, 2, wdims, strides, box, nullptr);
  if (rc) return rc;
  rc = make_tmap_f32(&tmW2, w2, 2, wdims, strides, box, nullptr);
  if (rc) return rc;
  CUtensorMap tmOut;
  {
    uint64_t dims[2] = {(uint64_t)SMF_C, (uint64_t)M};
    rc = make_tmap_f32(&tmOut, out, 2, dims, strides, box, nullptr);
    if (rc) return rc;
  }
  SwinMlpParams p{};
  p.passes = mma_passes();
  p.tok = tok; p.bp = bp; p.lnw = ln_w; p.lnb = ln_b; p.b1 = b1; p.b2 = b2; p.out = out; p.M = M;
  p.n_tiles = (int)((M + 127) / 128);
  const size_t smem = (size_t)(SMF_ATT_SLOTS + SMF_W_SLOTS + 4) * SMF_SLOT + (5 * SMF_C + 8 * 128) * sizeof(float) + 1024 /*align*/ + 512;
  OCC_ENSURE_SMEM(swin_mlp_fused_kernel, smem);
  int grid = sm_count();
  if (grid > p.n_tiles) grid = p.n_tiles;
  swin_mlp_fused_kernel<<<grid, SMF_THREADS, smem, stream>>>(tmA, tmWp, tmW1, tmW2, tmOut, p);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}
