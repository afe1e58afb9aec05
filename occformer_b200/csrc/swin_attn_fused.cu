// Shifted-window attention with the QKV projection fused in (C = 128): the kernel reads the LayerNorm'ed tokens once
// (S32, 512 B per token) instead of a materialised q/k/v tensor (1.5 KB per token written by a GEMM and read back).
//
//   WindowMSA.forward: qkv = Linear(C -> 3C)(x) ... attn = softmax(q k^T * scale + bias [+ mask]) v
//   (projects/mmdet3d_plugin/occformer/backbones/modules/window_attention.py:69-107), ShiftWindowMSA pad / roll / mask /
//   partition / reverse (:168-242) as index arithmetic, exactly as in window_attn.cu (same geometry header).
//
// Work unit = (pair of windows, head), one head per CTA (gridDim.x is a multiple of the head count): the head's 96 x 128
// slice of the qkv weight (rows [q | k | v] x 32, head-major) stays resident in shared memory for the whole kernel.
// Per unit:
//   loader   (warp 1)       one TMA box (32 channels x 128 rows) per k-block: the tokens arrive in WINDOW LAYOUT
//                           (window_geom.cuh: the LayerNorm producer writes token (img, x, y) to row window * 64 + t),
//                           so a pair of windows is a contiguous tile.  (Per-row gathers -- cp.async or tile::gather4,
//                           ~110 cycles per 512-byte gather4 -- could not feed the MMAs: r02_ncu_swin_qkv_attn_v3/v4.)
//   MMA      (warps 0/2/3)  M1: D[128 x 96] = A W_h^T            (4 k-blocks x 6 tcgen05.mma 128x96x16: three bf16 passes)
//                           QK: S = Q K^T, PV: O' = [P_hi; P_lo][V_hi | V_lo]   (as in window_attn.cu); one issuer each
//   three warpgroups (warps 4-7 / 8-11 / 12-15, unit u -> warpgroup u % 3), each running its unit start to end:
//       metadata: token row / shift-mask region of the thread's tile row, key-region bitmasks by two ballots per window
//                 half (computed while M1 of the unit is in flight)
//       convert : tcgen05.ld D -> + bias -> split -> V operand tile (one per warpgroup), then Q / K operand tiles (single:
//                 free again as soon as QK^T of the previous unit has been read)
//       softmax : S row -> scale, relative-position bias, shift mask, exp2 -> P (TMEM, split)      (unchanged)
//       epilogue: O -> normalise -> S32 scatter to the token-ordered output (A operand of the projection GEMM)
// A warpgroup's unit is a serial chain of TMEM round trips (IPC ~0.2 per warp), so the kernel's rate is (warpgroups in
// flight) / chain: three units are
// in flight, each with its own TMEM slot of 128 columns that holds D, then S, then P of the unit (D is dead once
// converted, S once exponentiated), plus two O buffers.  The token tiles arrive through a ring of five 16 KB k-block
// slots.  Measured (profiles/r02_ncu_swin_qkv_attn_v7.md): 0.48 ms per launch at 200x200x(16+1), tensor pipe 48 % active.
// HBM: tokens in (window layout: 64 rows per window) + attention out (rows*C*4); the 4 head-CTAs of a group walk the
// same window pairs at the same time, so three of the four token reads are L2 hits.
#include "window_geom.cuh"

namespace occ {

constexpr int SF_C = 128;
constexpr int SF_KB = SF_C / 32;                      // k-blocks of the projection
constexpr int SF_W_BYTES = SF_KB * 96 * 128;          // 49152: W_h as 4 K-major SWIZZLE_128B tiles of 96 rows
constexpr int SF_SLOT = 128 * 128;                    // 16384: one k-block of the token tile (128 rows x 128 B)
constexpr int SF_NSLOT = 5;                           // ring of k-block slots = 1.25 token tiles
constexpr int SF_NWG = 3;                             // converting warpgroups = units in flight
constexpr int SF_SAME_BYTES = 192;                    // per warpgroup: 2 window halves x 9 regions x 64-bit key masks (144), padded
constexpr int SF_OFF_A = SF_W_BYTES;                                    // 49152
constexpr int SF_OFF_Q = SF_OFF_A + SF_NSLOT * SF_SLOT;                 // 131072: Q tile, then K tile (single buffered)
constexpr int SF_OFF_V = SF_OFF_Q + 2 * WA_TILE;                        // 163840: V tiles (one per warpgroup)
constexpr int SF_OFF_BIAS = SF_OFF_V + SF_NWG * WA_TILE;                     // 212992: relative-position bias (49, 52) * log2 e
constexpr int SF_OFF_QB = SF_OFF_BIAS + WA_BIAS_BYTES;                  // 223232: this head's 96 qkv bias values
constexpr int SF_OFF_SAME = SF_OFF_QB + 384;                            // 223616
constexpr int SF_OFF_BAR = SF_OFF_SAME + SF_NWG * SF_SAME_BYTES;        // 224192
constexpr int SF_SMEM = SF_OFF_BAR + 384 + 1024;                        // 225600 <= 232448 (227 KB)
constexpr uint32_t SF_TMEM_U = 0, SF_TMEM_O = 384;  // unit slots 3 x 128 (D 96 cols -> S -> P), O: 2 x 64

__global__ void __launch_bounds__(WA_THREADS, 1)
swin_qkv_attn_kernel(const __grid_constant__ CUtensorMap tmap_tok /*window-layout tokens (npairs*128, 128) S32, box 32 x 128*/,
                     const float* __restrict__ wqkv /*(384, 128) S32,
                     head-major rows*/, const float* __restrict__ bqkv /*(384) fp32, head-major*/,
                     const float* __restrict__ bias_pad /*(heads, 2404)*/, float* __restrict__ out, const WinGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sw = smem;
  uint8_t* sa = smem + SF_OFF_A;
  float* sb = reinterpret_cast<float*>(smem + SF_OFF_BIAS);
  float* sqb = reinterpret_cast<float*>(smem + SF_OFF_QB);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + SF_OFF_BAR);  // [SF_NSLOT] k-block slot filled
  uint64_t* a_empty = a_full + SF_NSLOT;                              // [SF_NSLOT] k-block slot consumed by M1
  // every barrier below has ONE waiter: the warpgroup (or the MMA thread) that owns index u % 3 -- a waiter that first
  // waits for an odd phase of a shared barrier cannot tell "phase 1 done" from "nothing done yet"
  uint64_t* d_ready = a_empty + SF_NSLOT;  // [3] M1 of unit u landed in TMEM slot u % 3
  uint64_t* qkv_full = d_ready + SF_NWG;   // [3] Q, K, V[u % 3] of the unit are in shared memory
  uint64_t* qk_free = qkv_full + SF_NWG;   // [3] QK^T of unit u has been read out of the Q / K tiles (waiter: unit u + 1)
  uint64_t* v_empty = qk_free + SF_NWG;    // [3] PV has been read out of V[u % 3]
  uint64_t* s_ready = v_empty + SF_NWG;    // [3]
  uint64_t* p_ready = s_ready + SF_NWG;    // [3]
  uint64_t* o_ready = p_ready + SF_NWG;    // [3]
  uint64_t* o_free = o_ready + SF_NWG;     // [2] O buffer u & 1 has been read by the epilogue (waiter: MMA thread)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 2);
  static_assert((2 * SF_NSLOT + 7 * SF_NWG + 2) * 8 + 4 <= 384, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long npairs = (g.nwin + 1) / 2;
  const int H = g.heads;
  const int h = blockIdx.x % H;
  const int pair0 = blockIdx.x / H, pair_stride = gridDim.x / H;
  const long long n_units = (npairs > pair0) ? (npairs - pair0 + pair_stride - 1) / pair_stride : 0;

  // resident operands of this head: W_h (96 rows of the head-major qkv weight) as 4 K-major swizzled k-blocks, its bias,
  // the relative-position bias table (log2 domain)
  for (int i = threadIdx.x; i < 96 * 32; i += WA_THREADS) {  // 16-byte chunks: row r, chunk cc of the 512-byte row
    const int r = i >> 5, cc = i & 31;
    const int kb = cc >> 3, c = cc & 7;
    const float4 v = __ldg(reinterpret_cast<const float4*>(wqkv + ((size_t)h * 96 + r) * SF_C) + cc);
    *reinterpret_cast<float4*>(sw + kb * (96 * 128) + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  if (threadIdx.x < 96) sqb[threadIdx.x] = bqkv[h * 96 + threadIdx.x];
  for (int i = threadIdx.x; i < WT * WA_BIAS_LD; i += WA_THREADS) {
    const int r = i / WA_BIAS_LD, c = i % WA_BIAS_LD;
    sb[i] = c < WT ? bias_pad[(size_t)h * WA_BIAS_FLOATS + r * WT + c] * 1.4426950408889634f : 0.f;
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SF_NSLOT; ++i) {
      mbar_init(&a_full[i], 1);               // expect_tx by the loader; the 32 row gathers complete the bytes
      mbar_init(&a_empty[i], 1);              // tcgen05.commit after the k-block's MMAs
    }
    for (int i = 0; i < SF_NWG; ++i) {
      mbar_init(&d_ready[i], 1);
      mbar_init(&qkv_full[i], 4);  // the four warps of the converting warpgroup
      mbar_init(&qk_free[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_ready[i], 1);
    }
    mbar_init(&o_free[0], 4);
    mbar_init(&o_free[1], 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  fence_proxy_async_smem();  // W_h was written with generic stores; the MMAs read it through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 1) {
    // ===================================================================== loader: one TMA box per k-block
    if (lane == 0) {
      uint32_t slot = 0, par = 1;  // parity of the "slot is empty" phase to wait for (fresh barriers pass parity 1)
      for (uint32_t u = 0; u < (uint32_t)n_units; ++u) {
        const int row0 = (int)((pair0 + (long long)u * pair_stride) * 128);
#pragma unroll 1
        for (int kb = 0; kb < SF_KB; ++kb) {
          mbar_wait(&a_empty[slot], par);
          mbar_expect_tx(&a_full[slot], SF_SLOT);
          tma_load_2d(sa + (size_t)slot * SF_SLOT, &tmap_tok, &a_full[slot], kb * 32, row0);
          if (++slot == SF_NSLOT) { slot = 0; par ^= 1; }
        }
      }
    }
  } else if (warp == 0 || warp == 2 || warp == 3) {
    // ===================================================================== MMA issuers: M1 (warp 0), QK (warp 2), PV (warp 3)
    // One issuing lane per kind, each blocking on its own operands only, 32-bit counters advanced incrementally.  (A single
    // lane polling for all three kinds -- with 64-bit counters, % and / in the first version -- needed ~1 k cycles per
    // poll round, i.e. per M1 k-block: the issuer's own instruction stream, not the tensor pipe, set the pace:
    // profiles/r02_ncu_swin_qkv_attn_v5/v6.md.)  tcgen05.commit tracks the issuing thread's own MMAs, and every
    // dependency between the three kinds is a completion barrier, so nothing relies on cross-thread issue order.
    constexpr uint32_t IDESC_M1 = make_idesc_bf16(128, 96, 0, 0);
    constexpr uint32_t IDESC_QK = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 2 * HD, 0, 1);
    const uint32_t nu = (uint32_t)n_units;
    if (lane == 0 && warp == 0) {
      const uint64_t adesc0 = make_sw128_desc(smem_u32(sa), 1024, 16);
      const uint64_t bdesc0 = make_sw128_desc(smem_u32(sw), 1024, 16);
      uint32_t slot = 0, par = 0;  // ring slot of the next k-block, parity of its "full" phase
      uint32_t s1 = 0, p1 = 1;     // u % 3; parity of the v_empty phase that frees TMEM slot s1 (first use: no wait)
      for (uint32_t u = 0; u < nu; ++u) {
        // TMEM slot s1 still holds the P of unit u - 3 until its PV has completed
        if (u >= (uint32_t)SF_NWG) mbar_wait(&v_empty[s1], p1);
#pragma unroll 1
        for (uint32_t kb = 0; kb < (uint32_t)SF_KB; ++kb) {
          mbar_wait(&a_full[slot], par);
          tc_fence_after();
          mma_bf16x3_ss(tmem_base + SF_TMEM_U + s1 * 128, adesc0 + (uint64_t)(slot * (SF_SLOT >> 4)),
                        bdesc0 + (uint64_t)(kb * ((96 * 128) >> 4)), IDESC_M1, kb != 0, g.passes);
          mma_commit(&a_empty[slot]);
          // (holding the projection to one k-block in flight, so that QK^T / PV never queue behind 24 MMAs, was measured
          // slower: 0.54 vs 0.48 ms)
          if (++slot == SF_NSLOT) { slot = 0; par ^= 1; }
        }
        mma_commit(&d_ready[s1]);
        if (++s1 == SF_NWG) { s1 = 0; p1 ^= 1; }
      }
    } else if (lane == 0 && warp == 2) {
      const uint64_t qdesc = make_sw128_desc(smem_u32(smem + SF_OFF_Q), 1024, 16);
      const uint64_t kdesc = make_sw128_desc(smem_u32(smem + SF_OFF_Q) + WA_TILE, 1024, 16);
      uint32_t sq = 0, pq = 0;
      for (uint32_t u = 0; u < nu; ++u) {
        mbar_wait(&qkv_full[sq], pq);  // Q / K / V tiles converted (D of the unit, in the same TMEM slot, is dead by then)
        fence_proxy_async_smem();
        tc_fence_after();
        mma_bf16x3_ss(tmem_base + SF_TMEM_U + sq * 128, qdesc, kdesc, IDESC_QK, 0u, g.passes);
        mma_commit(&s_ready[sq]);
        mma_commit(&qk_free[sq]);  // the single Q / K tiles may be overwritten by the next unit's conversion
        if (++sq == SF_NWG) { sq = 0; pq ^= 1; }
      }
    } else if (lane == 0 && warp == 3) {
      const uint64_t vdesc0 = make_sw128_desc(smem_u32(smem + SF_OFF_V), 1024, 1024);
      uint32_t sp = 0, pp = 0;
      for (uint32_t u = 0; u < nu; ++u) {
        mbar_wait(&p_ready[sp], pp);                          // P of the unit is in TMEM
        mbar_wait(&o_free[u & 1], ((u >> 1) & 1) ^ 1);        // O buffer u & 1 read by the epilogue of unit u - 2
        tc_fence_after();
        const uint64_t vdesc = vdesc0 + (uint64_t)(sp * (WA_TILE >> 4));
        const uint32_t p_tmem = tmem_base + SF_TMEM_U + sp * 128;
        const uint32_t o_tmem = tmem_base + SF_TMEM_O + (u & 1) * 2 * HD;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t pc = p_tmem + (kk >> 1) * 32 + (kk & 1) * 8;
          mma_bf16_ts(o_tmem, pc, vdesc + (uint64_t)(kk * 128), IDESC_PV, kk != 0);
          if (g.passes == 3) mma_bf16_ts(o_tmem, pc + 16, vdesc + (uint64_t)(kk * 128), IDESC_PV, 1u);
        }
        mma_commit(&o_ready[sp]);
        mma_commit(&v_empty[sp]);
        if (++sp == SF_NWG) { sp = 0; pp ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== conversion / softmax / epilogue warpgroups
    const int wg = (warp - 4) >> 2;
    const int i = ((warp & 3) << 5) + lane;  // token row of the tile = TMEM lane
    const int half = i >> 6, t = i & 63;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float scale = 0.17677669529663687f;  // 32^-0.5
    const int tb = wg;  // TMEM slot, V tile and barrier index of this warpgroup's units (u % 3 == wg)
    for (long long u = wg; u < n_units; u += SF_NWG) {
      const uint32_t k = (uint32_t)(u / SF_NWG);
      const int ob = (int)(u & 1);
      // ---- metadata of this thread's tile row (computed while M1 of the unit is still in flight): global token row for
      // the output scatter, shift-mask region, and -- two ballots per window half, exchanged through shared memory --
      // the 64-bit sets of the window's keys that lie in each of the 9 regions
      long long my_row = -2;  // -2: MMA padding row, -1: window pad token
      int my_reg = 0;
      {
        const int win = (int)(2 * (pair0 + u * pair_stride)) + half;
        if (t < WT && win < (int)g.nwin) {
          const int nimg = g.B * (g.Z + 1);  // window index = (wx * nWy + wy) * n_images + img (window_geom.cuh)
          const int img = win % nimg, w2 = win / nimg;
          my_row = window_token_row(g, img, w2 / g.nWy, w2 % g.nWy, t, &my_reg);
        }
      }
      uint2 same = make_uint2(0xffffffffu, 0x1ffffu);  // un-shifted partition: one region, every key of the window
      if (g.shift) {
        uint32_t* same32 = reinterpret_cast<uint32_t*>(smem + SF_OFF_SAME + wg * SF_SAME_BYTES);
        const int wq = warp & 3;  // window half wq >> 1, tokens (wq & 1) * 32 + lane
#pragma unroll
        for (int rg = 0; rg < 9; ++rg) {
          const uint32_t bal = __ballot_sync(0xffffffffu, t < WT && my_reg == rg);
          if (lane == 0) same32[((wq >> 1) * 9 + rg) * 2 + (wq & 1)] = bal;
        }
        named_bar_sync(3 + wg, 128);
        same = reinterpret_cast<const uint2*>(same32)[half * 9 + my_reg];
      }
      // ---- conversion: D row (q | k | v of this head) + bias -> S32 rows of the Q / K / V operand tiles
      mbar_wait(&d_ready[tb], k & 1);
      tc_fence_after();
      uint32_t ra[32], rb[32];
      {
        const int sw7 = i & 7;
#pragma unroll 1
        for (int step = 0; step < 3; ++step) {  // v first (its buffer is free early), then q, k (wait for QK of unit u - 1)
          const int which = step == 0 ? 2 : step - 1;
          if (step == 0) mbar_wait(&v_empty[tb], (k & 1) ^ 1);
          if (step == 1 && u > 0) mbar_wait(&qk_free[(u - 1) % SF_NWG], (uint32_t)(((u - 1) / SF_NWG) & 1));
          tmem_ld_32x32(lane_base + SF_TMEM_U + tb * 128 + which * 32, ra);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]) + sqb[which * 32 + j];
          split_chunk32(v, rb);
          uint8_t* drow = smem + (which == 2 ? SF_OFF_V + tb * WA_TILE : SF_OFF_Q + which * WA_TILE) + i * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(drow + ((c ^ sw7) << 4)) = make_uint4(rb[4 * c], rb[4 * c + 1], rb[4 * c + 2], rb[4 * c + 3]);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&qkv_full[tb]);
      }
      // ---- softmax (as window_attn.cu)
      mbar_wait(&s_ready[tb], k & 1);
      tc_fence_after();
      const uint32_t diff_lo = ~same.x, diff_hi = ~same.y & 0x1FFFFu;
      const bool uniform = (diff_lo | diff_hi) == 0u;
      const float4* brow4 = reinterpret_cast<const float4*>(sb + (t < WT ? t : 0) * WA_BIAS_LD);
      const uint32_t s_col = lane_base + SF_TMEM_U + tb * 128 + half * 64;
      tmem_ld_32x32(s_col, ra);
      tmem_ld_32x32(s_col + 32, rb);
      tmem_ld_wait();
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
      split_chunk32(pa, ra);
      split_chunk32(pb, rb);
      tmem_st_32x32(s_col, ra);
      tmem_st_32x32(s_col + 32, rb);
#pragma unroll
      for (int j = 0; j < 32; ++j) ra[j] = 0u;
      const uint32_t o_col = lane_base + SF_TMEM_U + tb * 128 + (half ^ 1) * 64;
      tmem_st_32x32(o_col, ra);
      tmem_st_32x32(o_col + 32, ra);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[tb]);
      // ---- epilogue
      mbar_wait(&o_ready[tb], k & 1);
      tc_fence_after();
      tmem_ld_32x32(lane_base + SF_TMEM_O + ob * 2 * HD, ra);
      tmem_ld_32x32(lane_base + SF_TMEM_O + ob * 2 * HD + HD, rb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[ob]);
      if (t < WT && my_row >= 0) {
        const float inv = 1.0f / sum;
        float o[32];
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = (__uint_as_float(ra[d]) + __uint_as_float(rb[d])) * inv;
        split_chunk32(o, ra);
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
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace occ

using namespace occ;

// tokn (occ_window_layout_rows(B,X,Y,Z), 128) S32 = LayerNorm1'ed tokens in the WINDOW LAYOUT of this block's partition
// (shift), zero where no token lands (occ_gn_relu_zmean_ln with win_shift = shift writes it); wqkv (384, 128) S32 and
// bqkv (384) fp32 with HEAD-MAJOR rows [head][q|k|v][32]; bias_pad as occ_window_attention; out (rows, 128) S32.
// Returns -2 (unsupported) for C != 128: the caller then runs occ_gemm_bf16x3 + occ_window_attention.
extern "C" int occ_swin_qkv_attention(const float* tokn, const float* wqkv, const float* bqkv, const float* bias_pad,
                                      float* out, int B, int X, int Y, int Z, int C, int heads, int shift,
                                      cudaStream_t stream) {
  OCC_REQUIRE(tokn && wqkv && bqkv && bias_pad && out);
  OCC_REQUIRE(B > 0 && X > 0 && Y > 0 && Z > 0 && heads > 0 && C == heads * HD);
  if (C != SF_C) return OCC_EUNSUPPORTED;
  OCC_REQUIRE((reinterpret_cast<uintptr_t>(tokn) & 15) == 0 && (reinterpret_cast<uintptr_t>(wqkv) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(bias_pad) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const WinGeom g = make_win_geom(B, X, Y, Z, C, heads, shift);
  OCC_REQUIRE(g.nwin < (1ll << 31));
  static_assert(SF_SMEM <= 227 * 1024, "shared memory budget");
  OCC_ENSURE_SMEM(swin_qkv_attn_kernel, SF_SMEM);
  const long long npairs = (g.nwin + 1) / 2;
  long long groups = sm_count() / heads;
  if (groups > npairs) groups = npairs;
  if (groups < 1) groups = 1;
  CUtensorMap tm;
  {
    const long long rows = (g.nwin + 1) / 2 * 128;  // occ_window_layout_rows
    const uint64_t dims[2] = {(uint64_t)SF_C, (uint64_t)rows}, strides[1] = {(uint64_t)SF_C * 4};
    const uint32_t box[2] = {32, 128};
    OCC_REQUIRE(make_tmap_f32(&tm, tokn, 2, dims, strides, box, nullptr) == OCC_OK);
  }
  swin_qkv_attn_kernel<<<(int)(groups * heads), WA_THREADS, SF_SMEM, stream>>>(tm, wqkv, bqkv, bias_pad, out, g);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}
