// Shifted-window multi-head self-attention core over the B*(Z+1) X-Y images of the dual-path encoder, on the
// 5th-generation tensor cores (tcgen05.mma on split-bf16 operands, three passes per contraction = fp32-faithful, see
// occ_ptx.cuh; accumulators and the probability operand in TMEM).  qkv arrives in the S32 split format (one 128-byte
// chunk per (token, head, q|k|v)), the output is written in S32 (A operand of the projection GEMM).
//
// Replaces ShiftWindowMSA.forward's pad / roll / mask build / window_partition / window_reverse / roll back /
// crop and WindowMSA.forward's score pipeline (q*scale, QK^T, + relative-position bias, + shift mask, softmax,
// PV) -- projects/mmdet3d_plugin/occformer/backbones/modules/window_attention.py:168-242 and :69-107 -- with
// index arithmetic: the kernel gathers the 49 tokens of a (possibly shifted, possibly padded) window straight
// from the token-ordered qkv tensor and scatters the result back to token order.  No (nW,49,49) mask tensor,
// no score tensor in HBM.
//
// Semantics kept:
//   * padding to a multiple of 7 happens AFTER LayerNorm: pad tokens are zeros, so their q/k/v equal the qkv
//     bias and they take part in the softmax as real keys; their outputs are cropped.
//   * shifted blocks: roll(-3,-3); region ids from slices (0,-7),(-7,-3),(-3,None) on the padded extent;
//     additive mask -100.0 (not -inf) where ids differ.
//   * relative-position bias added after the q*scale product.
//
// Work unit = (pair of windows, head): M = 128 query rows (window A in rows 0..63, window B in rows 64..127, 49
// real rows each).  Persistent CTA, 512 threads:
//   warps 12-15  loaders : cp.async gather of the Q / K / V head slices (128 B per token) into 128B-swizzled
//                          K-major tiles, 3-stage ring, + the head's relative-position bias and the token metadata
//   warp 0       MMA     : S = Q K^T (6 x tcgen05.mma 128x128x16: hi*hi + lo*hi + hi*lo over the 32-wide head dim), then
//                          O' = [P_hi; P_lo] [V_hi | V_lo] (16 x 128x64x16 with P read from TMEM and the V chunk rows as an
//                          MN-major SWIZZLE_128B operand, N = 64 = hi | lo halves of the head dim -- no transpose of V
//                          anywhere); O = O'[:, :32] + O'[:, 32:] in the epilogue
//   warps 4-7 / 8-11     : two softmax warpgroups (even / odd units): tcgen05.ld S row -> scale, bias, shift mask,
//                          softmax in registers -> tcgen05.st P (block-diagonal: the other window's columns are zero)
//                          -> after PV: tcgen05.ld O, normalise, scatter 128 B per (token, head) to global.
#include "window_geom.cuh"

namespace occ {

__global__ void __launch_bounds__(WA_THREADS, 1)
window_attn_tc_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                      const float* __restrict__ bias_pad /*(heads, 2404)*/, float* __restrict__ out, const WinGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  float* sb = reinterpret_cast<float*>(smem + WA_STAGES * WA_STAGE_BYTES);  // bias of head h, resident
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WA_STAGES * WA_STAGE_BYTES + WA_BIAS_BYTES);
  uint64_t* empty_bar = full_bar + WA_STAGES;
  uint64_t* s_ready = empty_bar + WA_STAGES;  // [2]
  uint64_t* p_ready = s_ready + 2;            // [2]
  uint64_t* o_ready = p_ready + 2;            // [2]
  uint64_t* o_free = o_ready + 2;             // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long npairs = (g.nwin + 1) / 2;
  const int H = g.heads;
  // One head per CTA (gridDim.x is a multiple of H): CTA c works on head c % H of the window pairs c / H,
  // + gridDim.x / H, ...  The H CTAs of a group walk the same pairs at the same time, so the 12 128-byte segments of
  // every qkv token row (one DRAM page) are fetched together, and the head's bias table stays resident in smem.
  const int h = blockIdx.x % H;
  // head-major qkv: the head's q, k, v slices are one contiguous 384-byte run per token (three adjacent lines of one
  // DRAM page); reference order: three 128-byte slices C floats apart
  const int h_off = g.head_major ? h * 3 * HD : h * HD;
  const int kv_step = g.head_major ? HD : g.C;
  const int pair0 = blockIdx.x / H, pair_stride = gridDim.x / H;
  const long long n_units = (npairs > pair0) ? (npairs - pair0 + pair_stride - 1) / pair_stride : 0;
  for (int i = threadIdx.x; i < WT * WA_BIAS_LD; i += WA_THREADS) {  // scores live in the log2 domain (exp2 softmax)
    const int r = i / WA_BIAS_LD, c = i % WA_BIAS_LD;
    sb[i] = c < WT ? bias_pad[(size_t)h * WA_BIAS_FLOATS + r * WT + c] * 1.4426950408889634f : 0.f;
  }

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < WA_STAGES; ++i) {
      mbar_init(&full_bar[i], 256); // per loader thread: one asynchronous cp.async arrival + one release arrival
      mbar_init(&empty_bar[i], 1);  // tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_ready[i], 1);
      mbar_init(&o_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<WA_TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= 12) {
    // ===================================================================== loaders
    const int l = threadIdx.x - 12 * 32;  // 0..127
    const int C = g.C;
    auto issue = [&](long long u) {
      const int s = (int)(u % WA_STAGES);
      const long long pair = pair0 + u * pair_stride;
      uint8_t* st = smem + (size_t)s * WA_STAGE_BYTES;
      long long* rows = reinterpret_cast<long long*>(st + WA_OFF_ROWS);
      int* region = reinterpret_cast<int*>(st + WA_OFF_REGION);
      mbar_wait(&empty_bar[s], (uint32_t)(((u / WA_STAGES) & 1) ^ 1));
      {  // metadata of row l: window = 2*pair + l/64, token = l%64
        const long long win = 2 * pair + (l >> 6);
        const int t = l & 63;
        long long r = -2;  // -2: MMA padding row (zero), -1: window pad token (qkv = bias)
        int reg = 0;
        if (t < WT && win < g.nwin) {
          long long w = win;
          const int wy = (int)(w % g.nWy); w /= g.nWy;
          const int wx = (int)(w % g.nWx); w /= g.nWx;
          r = window_token_row(g, (int)w, wx, wy, t, &reg);
        }
        rows[l] = r;
        region[l] = reg;
        // global address of this token's q head-slice (k / v follow at +kv_step / +2 kv_step floats); pad tokens read
        // the qkv bias
        reinterpret_cast<const float**>(st + WA_OFF_SRC)[l] =
            r >= 0 ? qkv + r * 3 * C + h_off : (r == -1 ? qkv_bias + h_off : nullptr);
      }
      {  // shift-mask bookkeeping: same[half][r] = 64-bit set of the window's keys that lie in region r (two warp ballots)
        const int t = l & 63, wl = l >> 5;  // loader warp wl covers tokens 32*(wl&1) .. +31 of window half wl>>1
        const bool tok = t < WT;
        uint32_t* same32 = reinterpret_cast<uint32_t*>(st + WA_OFF_SAME);
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const uint32_t bal = __ballot_sync(0xffffffffu, tok && region[l] == r);
          if (lane == 0) same32[((wl >> 1) * 9 + r) * 2 + (wl & 1)] = bal;
        }
      }
      named_bar_sync(2, 128);
      // thread l copies 16-byte chunk c = l & 7 of rows (l >> 3) + 16*rr, rr = 0..7, for the three tiles.  (r & 7) and
      // (r & 3) do not depend on rr, so the swizzled chunk offsets are per-thread constants:
      //   Q, K: K-major SWIZZLE_128B (16-byte chunk c ^ (r & 7)); V: MN-major bf16 operand, N = 64 (hi | lo of the head
      //   dim) = one 128-byte row per key, 8-key atoms -- the same SWIZZLE_128B chunk pattern
      {
        const int c = l & 7, rb = l >> 3;
        const int off_qk = (c ^ (rb & 7)) << 4;
        const int off_v = off_qk;
        const float* const* srcp = reinterpret_cast<const float* const*>(st + WA_OFF_SRC);
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = rb + 16 * rr;
          uint8_t* drow = st + r * 128;
          const float* src = srcp[r];
          if (src == nullptr) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(drow + off_qk) = z;
            *reinterpret_cast<float4*>(drow + WA_TILE + off_qk) = z;
            *reinterpret_cast<float4*>(drow + 2 * WA_TILE + off_v) = z;
          } else {
            cp_async_16(drow + off_qk, src + c * 4);
            cp_async_16(drow + WA_TILE + off_qk, src + kv_step + c * 4);
            cp_async_16(drow + 2 * WA_TILE + off_v, src + 2 * kv_step + c * 4);
          }
        }
      }
      cp_async_mbar_arrive_noinc(&full_bar[s]);  // fires when this thread's copies of the unit have landed
      mbar_arrive(&full_bar[s]);                 // release: zero-fill stores + metadata
    };
    // the loaders never wait for their own copies: up to WA_STAGES units of gathers are in flight per CTA
    for (long long u = 0; u < n_units; ++u) issue(u);
    cp_async_wait<0>();
  } else if (warp == 0 || warp == 3) {
    // ===================================================================== MMA issuers: QK (warp 0), PV (warp 3)
    // One issuing lane per kind, each blocking on its own operands only, 32-bit counters advanced incrementally: PV(u) must
    // not queue behind the gathers of unit u+1, and QK(u+1) must not queue behind the softmax of unit u.  (One lane
    // polling for both kinds with 64-bit counters spent more time in its own instruction stream than the MMAs take; see
    // swin_attn_fused.cu.)  tcgen05.commit tracks the issuing thread's own MMAs and every dependency between the two kinds
    // is a completion barrier: S/P buffer (u & 1) is free for QK(u) once PV(u - 2) has COMPLETED (o_ready).
    constexpr uint32_t IDESC_QK = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 2 * HD, 0, 1);  // B = V tile, MN-major (hi | lo of d contiguous)
    const uint32_t nu = (uint32_t)n_units;
    if (lane == 0 && warp == 0) {
      uint32_t s = 0, par = 0;  // stage of unit u, parity of its "full" phase
      for (uint32_t u = 0; u < nu; ++u) {
        const uint32_t tb = u & 1, k = u >> 1;
        if (u >= 2) mbar_wait(&o_ready[tb], (k - 1) & 1);
        mbar_wait(&full_bar[s], par);
        fence_proxy_async_smem();  // cp.async / st.shared (generic proxy) -> tcgen05.mma operand reads (async proxy)
        tc_fence_after();
        const uint32_t qaddr = smem_u32(smem + (size_t)s * WA_STAGE_BYTES);
        const uint64_t qdesc = make_sw128_desc(qaddr, 1024, 16);
        const uint64_t kdesc = make_sw128_desc(qaddr + WA_TILE, 1024, 16);
        mma_bf16x3_ss(tmem_base + tb * 128, qdesc, kdesc, IDESC_QK, 0u, g.passes);
        mma_commit(&s_ready[tb]);
        if (++s == WA_STAGES) { s = 0; par ^= 1; }
      }
    } else if (lane == 0 && warp == 3) {
      uint32_t s = 0;
      for (uint32_t u = 0; u < nu; ++u) {
        const uint32_t tb = u & 1, k = u >> 1;
        mbar_wait(&p_ready[tb], k & 1);
        mbar_wait(&o_free[tb], (k & 1) ^ 1);
        tc_fence_after();
        const uint32_t vaddr = smem_u32(smem + (size_t)s * WA_STAGE_BYTES + 2 * WA_TILE);
        const uint64_t vdesc = make_sw128_desc(vaddr, 1024, 1024);  // MN-major: SBO = stride between 8-key atoms
        const uint32_t p_tmem = tmem_base + tb * 128;
        const uint32_t o_tmem = tmem_base + 256 + tb * 2 * HD;
        // P: four 32-key chunks of 32 columns = [16 packed hi | 16 packed lo]; 16 keys per MMA = two 8-key (1024 B) atoms
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t pc = p_tmem + (kk >> 1) * 32 + (kk & 1) * 8;
          mma_bf16_ts(o_tmem, pc, vdesc + (uint64_t)(kk * 128), IDESC_PV, kk != 0);       // P_hi [V_hi | V_lo]
          if (g.passes == 3) mma_bf16_ts(o_tmem, pc + 16, vdesc + (uint64_t)(kk * 128), IDESC_PV, 1u);  // P_lo [V_hi | V_lo]
        }
        mma_commit(&o_ready[tb]);
        mma_commit(&empty_bar[s]);
        if (++s == WA_STAGES) s = 0;
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================================================== softmax / epilogue warpgroups
    const int wg = (warp - 4) >> 2;                 // 0: even units, 1: odd units
    const int i = ((warp & 3) << 5) + lane;         // query row = TMEM lane 0..127
    const int half = i >> 6, t = i & 63;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float scale = 0.17677669529663687f;  // 32^-0.5
    const int tb = wg;
    for (long long u = wg; u < n_units; u += 2) {
      const int s = (int)(u % WA_STAGES);
      const uint32_t k = (uint32_t)(u >> 1);
      const uint8_t* st = smem + (size_t)s * WA_STAGE_BYTES;
      const long long* rows = reinterpret_cast<const long long*>(st + WA_OFF_ROWS);
      const int* region = reinterpret_cast<const int*>(st + WA_OFF_REGION);
      mbar_wait(&s_ready[tb], k & 1);
      tc_fence_after();
      const long long my_row = rows[i];
      const int my_reg = region[i];
      // keys that are NOT in this query's region get the additive -100 of the reference's shift mask
      const uint2 same = reinterpret_cast<const uint2*>(st + WA_OFF_SAME)[half * 9 + my_reg];
      const uint32_t diff_lo = ~same.x, diff_hi = ~same.y & 0x1FFFFu;  // 49 keys: 32 + 17
      const bool uniform = (diff_lo | diff_hi) == 0u;  // true for every un-shifted block and for interior windows
      const float4* brow4 = reinterpret_cast<const float4*>(sb + (t < WT ? t : 0) * WA_BIAS_LD);
      const uint32_t s_col = lane_base + tb * 128 + half * 64;
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(s_col, ra);
      tmem_ld_32x32(s_col + 32, rb);
      tmem_ld_wait();
      // z = (s * scale + bias [- 100]) * log2(e): one FFMA per score (bias pre-multiplied), four independent max chains
      const float sl2 = scale * 1.4426950408889634f, neg = -100.0f * 1.4426950408889634f;
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 bv = brow4[j4];
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * j4 + e;
          float z = fmaf(__uint_as_float(ra[j]), sl2, bb[e]);
          if (!uniform) z += ((diff_lo >> j) & 1u) ? neg : 0.f;
          ra[j] = __float_as_uint(z);
          mx[e] = fmaxf(mx[e], z);
        }
      }
#pragma unroll
      for (int j4 = 0; j4 < 5; ++j4) {
        const float4 bv = brow4[8 + j4];
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * j4 + e;
          if (32 + j < WT) {
            float z = fmaf(__uint_as_float(rb[j]), sl2, bb[e]);
            if (!uniform) z += ((diff_hi >> j) & 1u) ? neg : 0.f;
            rb[j] = __float_as_uint(z);
            mx[e] = fmaxf(mx[e], z);
          }
        }
      }
      const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      float sm4[4] = {0.f, 0.f, 0.f, 0.f};
      float pa[32], pb[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float p = ex2_approx(__uint_as_float(ra[j]) - m);
        sm4[j & 3] += p;
        pa[j] = p;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float p = 0.f;
        if (j < WT - 32) {
          p = ex2_approx(__uint_as_float(rb[j]) - m);
          sm4[j & 3] += p;
        }
        pb[j] = p;
      }
      const float sum = (sm4[0] + sm4[1]) + (sm4[2] + sm4[3]);
      split_chunk32(pa, ra);  // P as a TMEM A operand: per 32-key chunk 16 packed hi columns | 16 packed lo columns
      split_chunk32(pb, rb);
      // P row: own window's 64 key columns, zeros in the other window's 64 columns (block-diagonal)
      tmem_st_32x32(s_col, ra);
      tmem_st_32x32(s_col + 32, rb);
#pragma unroll
      for (int j = 0; j < 32; ++j) ra[j] = 0u;
      const uint32_t o_col = lane_base + tb * 128 + (half ^ 1) * 64;
      tmem_st_32x32(o_col, ra);
      tmem_st_32x32(o_col + 32, ra);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[tb]);
      // ---- epilogue of the same unit once PV has landed
      mbar_wait(&o_ready[tb], k & 1);
      tc_fence_after();
      tmem_ld_32x32(lane_base + 256 + tb * 2 * HD, ra);       // P [V_hi]
      tmem_ld_32x32(lane_base + 256 + tb * 2 * HD + HD, rb);  // P [V_lo]
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[tb]);
      if (t < WT && my_row >= 0) {
        const float inv = 1.0f / sum;
        float o[32];
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = (__uint_as_float(ra[d]) + __uint_as_float(rb[d])) * inv;
        split_chunk32(o, ra);  // the head's 32 output channels are one S32 chunk of the row
        uint4* dst = reinterpret_cast<uint4*>(out + my_row * g.C + h * HD);
#pragma unroll
        for (int d = 0; d < 8; ++d) dst[d] = make_uint4(ra[4 * d], ra[4 * d + 1], ra[4 * d + 2], ra[4 * d + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<WA_TMEM_COLS>(tmem_base);
  }
}

}  // namespace occ

using namespace occ;

// qkv (rows, 3C) with rows = B*X*Y*(Z+1) token-ordered (voxel tokens then BEV tokens); out (rows, C) in S32 (A operand of the projection GEMM); qkv and qkv_bias in S32.
// bias_pad = relative_position_bias_table[relative_position_index] arranged (heads, 49*49 padded to 2404 floats).
extern "C" int occ_window_attention(const float* qkv, const float* qkv_bias, const float* bias_pad, float* out, int B,
                                    int X, int Y, int Z, int C, int heads, int shift, int qkv_head_major,
                                    cudaStream_t stream) {
  OCC_REQUIRE(qkv && qkv_bias && bias_pad && out);
  OCC_REQUIRE(B > 0 && X > 0 && Y > 0 && Z > 0 && heads > 0 && C == heads * HD);
  OCC_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(qkv_bias) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(bias_pad) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  WinGeom g = make_win_geom(B, X, Y, Z, C, heads, shift);
  g.head_major = qkv_head_major ? 1 : 0;
  OCC_REQUIRE(g.nwin < (1ll << 31));
  const size_t smem = (size_t)WA_STAGES * WA_STAGE_BYTES + WA_BIAS_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  OCC_ENSURE_SMEM(window_attn_tc_kernel, smem);
  const long long npairs = (g.nwin + 1) / 2;
  OCC_REQUIRE(heads <= sm_count());
  long long groups = sm_count() / heads;  // CTA groups of `heads` CTAs, one head each
  if (groups > npairs) groups = npairs;
  const int grid = (int)(groups * heads);
  window_attn_tc_kernel<<<grid, WA_THREADS, smem, stream>>>(qkv, qkv_bias, bias_pad, out, g);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}
