// Persistent warp-specialised fp32-faithful GEMM / implicit-GEMM convolution for sm_100a ("bf16x3").
//
//   D[M,N] = epilogue( A[M,K] * W[N,K]^T )          (plain mode: A row-major, W = torch Linear weight)
//   D[vox,Cout] = epilogue( im2col(x)[vox, taps*Cin] * W2[Cout, taps*Cin]^T )   (conv mode)
//
// Operands are fp32 values stored in the S32 split format (occ_ptx.cuh: every 32-column chunk = 32 bf16 hi | 32 bf16
// lo); each 32-k block is contracted by three bf16 tensor-core passes (hi*hi + lo*hi + hi*lo, fp32 accumulate in
// TMEM): ~1e-5 relative error, the bytes, TMA boxes and shared-memory tiles of a plain fp32 operand.
//
// One CTA per SM, 384 threads: warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle), warp 1 =
// tcgen05.mma issuer (kind::f16 / bf16, M=128, N=BN, K=16 per instruction, fp32 accumulators in TMEM, two
// accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1), warp 2 = TMEM
// allocator, warps 4..7 = epilogue (tcgen05.ld 32x32b -> registers -> bias / activation / residual /
// GroupNorm statistics -> global).  Conv mode never materialises im2col: every filter tap is a 5-D TMA
// box {32 ch, bz, by, bx, 1} with a shifted origin; out-of-bounds = zero fill = the conv padding;
// stride-2 convs use the tensor map's elementStrides.
//
// Replaces on the reference path: cuDNN Conv3d + cuBLAS Linear calls inside DualpathTransformerBlock
// (projects/mmdet3d_plugin/occformer/backbones/dualpath_block.py:36-48,79), SwinBlock / WindowMSA
// linears (backbones/modules/window_attention.py:65-67,336-344), BottleNeckASPP convs
// (backbones/modules/aspp.py:49-172) and the decoder's K/V projections + mask einsum
// (mask2former/mask2former_nusc_occ.py:446-457).
#include "occ_common.cuh"
#include "occ_ptx.cuh"

namespace occ {

constexpr int BM = 128;
constexpr int BK = 32;  // 32 fp32 = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 4;
constexpr int EPI_BUF_BYTES = 32 * 128;                 // one 32-row x 32-column fp32 chunk, 128-byte swizzled rows
constexpr int EPI_BYTES = 8 /*warps*/ * EPI_BUF_BYTES;
constexpr int CTRL_BYTES = 2048;   // mbarriers + TMEM pointer + fp64 GroupNorm partial sums of up to 96 groups; a multiple
                                   // of 1024 so that the swizzled epilogue staging buffers behind it stay 1024-byte aligned
constexpr int MAX_STAT_GROUPS = 96;  // 2 * 96 doubles = 1536 B of CTRL_BYTES (barriers + TMEM pointer < 256 B)
constexpr int MAX_TAPS = 28;

struct GemmParams {
  int M, N, K, num_k_blocks;
  float* out;
  int ldo;
  const float* bias;
  const float* residual;
  int ldr;
  int splits;     // split-K: > 1 => every tile's k-blocks are divided among `splits` CTAs, partial sums meet in `out`
                  // (zero-initialised by the host) through TMA reduce-add; no bias / act / stats in the kernel
  int act;        // 0 none, 1 relu, 2 gelu(erf)
  int split_out;  // write the result in the S32 split format (the consumer is another tensor-core contraction)
  int passes;     // tensor-core passes per 32-k block (occ_common.cuh: mma_passes)
  // conv mode
  int conv;
  int Cin, KX, KY, KZ, dil, stride;
  int padx, pady, padz;
  int B, Xo, Yo, Zo;
  int bx, by, bz, tiles_x, tiles_y, tiles_z;
  int use_tma_store;  // epilogue writes through a TMA store (output rows 16-byte aligned)
  // fused adaptive max pooling of the output (decoder mask logits -> attention-mask logits), conv mode only:
  // pool_out[(cell), n] = max over the pw^3 voxels of the cell, as order-preserving ints (see enc_ordered)
  int* pool_out;
  int* pool_flag;  // [N]: set to 1 when some pooled value of column n is >= 0
  int pwx, pwy, pwz, pXo, pYo, pZo;
  int pool_nsteps, pool_stride[5];               // lane xor-strides of the in-warp butterfly (z, then y, then x bits)
  int pool_cwx, pool_cwy, pool_cwz, pool_ncy, pool_ncz;  // cell extent inside a tile, cells per tile along y / z
  int pool_complete;                             // every cell lies inside one tile: plain stores, no global atomics
  int store_out;   // 0: the GEMM output itself is not written (only pooled)
  // GroupNorm statistics (sum, sumsq per (batch, group)), accumulated in fp64
  double* gn_stats;
  int cpg;
  int rows_per_batch;  // plain mode: rows per batch sample (for gn_stats), else 0
  int* splitk_sem;     // [SPLITK_SEMS], see above (required when splits > 1)
  // explicit tap table (conv mode, stride 1): tap t reads the input at voxel + (tdx, tdy, tdz)[t] (dilation and padding
  // folded in); 0 = the regular KX x KY x KZ stencil.  Lets several dilated branches share one launch (ASPP).
  int ntaps;
  signed char tdx[MAX_TAPS], tdy[MAX_TAPS], tdz[MAX_TAPS];
};

// Split-K ordering: sem[(m, n) tile] counts the splits that have added their partial sum.  Split s adds after split
// s - 1, so the fp32 reduce-add order -- and with it every bit of the result -- is fixed.  The counters live in a
// caller-provided workspace (occ_conv_workspace_bytes(): 1024 ints, zero before the first call; the last split of
// every tile resets its counter, so one workspace serves every conv of a stream; concurrent streams pass their own).
constexpr int SPLITK_SEMS = 1024;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// order-preserving float <-> int map (signed compare of the ints == compare of the floats; sign is kept)
__device__ __forceinline__ int enc_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
constexpr int ENC_NEG = (int)0x80808080;  // below every encoded finite float; also the byte pattern of the memset

template <int CPG>
__device__ __forceinline__ void accum_stats(const float (&v)[32], float (&sv)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    sv[2 * (j / CPG)] += v[j];
    sv[2 * (j / CPG) + 1] += v[j] * v[j];
  }
}

// Sum each of 32 per-lane values over the 32 lanes of the warp with 31 shuffles; afterwards lane L holds
// the total of value index L in v[0].
__device__ __forceinline__ void warp_butterfly32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = upper ? v[i + n / 2] : v[i];
      const float send = upper ? v[i] : v[i + n / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// MT = M-tiles per CTA tile.  MT = 2 (stage-0 convs, N = 128): two 128-row accumulators share every B k-block, so the
// TMA engine writes 48 KB instead of 64 KB of shared memory per two tiles -- the 128x128 tile is bound by the
// shared-memory port (TMA writes + MMA operand reads), not by the tensor pipe.
template <int BN, int STAGES, int MT, bool LEAN>
__global__ void __launch_bounds__(384, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  // LEAN = plain GEMM, whole 32-column chunks, TMA store, no conv addressing / pooling / GroupNorm statistics / split-K:
  // those branches compile away.  (The Linears of the neck and of the C >= 256 Swin stages have 6-24 k-blocks per tile:
  // their pace is set by the epilogue warps' instruction stream -- 2.1 k instructions per warp and tile in the generic
  // kernel, profiles/r02_ncu_neck_ffn_gemms.md -- not by the 4.6 k tensor cycles of the tile.)
  const bool f_conv = !LEAN && p.conv != 0;
  const bool f_pool = !LEAN && p.pool_out != nullptr;
  const bool f_stats = !LEAN && p.gn_stats != nullptr;
  const int f_splits = LEAN ? 1 : p.splits;
  const bool f_tma = LEAN || p.use_tma_store != 0;
  const bool f_store = LEAN || p.store_out != 0;

  constexpr int B_STAGE_BYTES = BN * BK * 4;
  constexpr int A_BYTES = MT * A_STAGE_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
  constexpr int ACC_COLS = BN * MT;  // accumulator columns of one tile (MT sub-tiles side by side)
  static_assert(2 * ACC_COLS <= 512, "two accumulator buffers must fit TMEM");
  constexpr uint32_t TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128
                                 : (2 * ACC_COLS <= 256) ? 256 : 512;
  constexpr uint32_t IDESC = make_idesc_bf16(BM, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  double* stat_acc = reinterpret_cast<double*>(tmem_ptr + 2);  // (sum, sumsq) per group, <= MAX_STAT_GROUPS groups
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES + CTRL_BYTES;  // 8 warps x 4 KB, 1024-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_n_tiles = (p.N + BN - 1) / BN;
  const int num_m_tiles = f_conv ? p.B * p.tiles_x * p.tiles_y * p.tiles_z : (p.M + BM - 1) / BM;
  const int num_m_groups = (num_m_tiles + MT - 1) / MT;
  const int num_tiles = num_m_groups * num_n_tiles * f_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_ptr);
  if (threadIdx.x >= 128) {
    for (int i = threadIdx.x - 128; i < 2 * MAX_STAT_GROUPS; i += 256) stat_acc[i] = 0.0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int cblocks = f_conv ? (p.Cin / BK) : 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile % f_splits;
        const int n_tile = (tile / f_splits) % num_n_tiles;
        const int m_group = tile / (f_splits * num_n_tiles);
        const int kb0 = (int)(((long long)split * p.num_k_blocks) / f_splits);
        const int kb1 = (int)(((long long)(split + 1) * p.num_k_blocks) / f_splits);
        int cb[MT], cx0[MT], cy0[MT], cz0[MT];
#pragma unroll
        for (int sub = 0; sub < MT; ++sub) {
          cb[sub] = cx0[sub] = cy0[sub] = cz0[sub] = 0;
          if (f_conv) {
            int t = m_group * MT + sub;  // past the last tile: cb == B, the TMA box is out of bounds and zero-filled
            const int tz = t % p.tiles_z; t /= p.tiles_z;
            const int ty = t % p.tiles_y; t /= p.tiles_y;
            const int tx = t % p.tiles_x; t /= p.tiles_x;
            cb[sub] = t;
            cx0[sub] = tx * p.bx * p.stride - p.padx;
            cy0[sub] = ty * p.by * p.stride - p.pady;
            cz0[sub] = tz * p.bz * p.stride - p.padz;
          }
        }
        int tap = 0, kc = 0, tx_ = 0, ty_ = 0, tz_ = 0;
        if (f_conv && kb0 > 0) {  // split-K: start in the middle of the tap sequence
          tap = kb0 / cblocks;
          kc = kb0 % cblocks;
          tz_ = tap % p.KZ;
          ty_ = (tap / p.KZ) % p.KY;
          tx_ = tap / (p.KZ * p.KY);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (f_conv) {
            // input offset of this tap: regular stencil (dilated, padding folded into c?0) or the explicit table
            const int ox = p.ntaps ? p.tdx[tap] : tx_ * p.dil, oy = p.ntaps ? p.tdy[tap] : ty_ * p.dil,
                      oz = p.ntaps ? p.tdz[tap] : tz_ * p.dil;
#pragma unroll
            for (int sub = 0; sub < MT; ++sub)
              tma_load_5d(sa + sub * A_STAGE_BYTES, &tmA, &full_bar[stage], kc * BK, cz0[sub] + oz, cy0[sub] + oy,
                          cx0[sub] + ox, cb[sub]);
            if (++kc == cblocks) {
              kc = 0;
              ++tap;
              if (++tz_ == p.KZ) {
                tz_ = 0;
                if (++ty_ == p.KY) { ty_ = 0; ++tx_; }
              }
            }
          } else {
#pragma unroll
            for (int sub = 0; sub < MT; ++sub)
              tma_load_2d(sa + sub * A_STAGE_BYTES, &tmA, &full_bar[stage], kb * BK, (m_group * MT + sub) * BM);
          }
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_tile * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        (void)tap;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (single thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&tmem_empty[buf], (((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
        const int split = tile % f_splits;
        const int kb0 = (int)(((long long)split * p.num_k_blocks) / f_splits);
        const int kb1 = (int)(((long long)(split + 1) * p.num_k_blocks) / f_splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t bdesc = make_sw128_desc(sb, 1024, 16);
#pragma unroll
          for (int sub = 0; sub < MT; ++sub) {
            const uint64_t adesc = make_sw128_desc(sa + sub * A_STAGE_BYTES, 1024, 16);
            // one S32 row block = 32 k values: hi*hi + lo*hi + hi*lo, six K=16 MMAs stepping 32 bytes inside the row
            mma_bf16x3_ss(d_tmem + sub * BN, adesc, bdesc, IDESC, kb != kb0, p.passes);
          }
          mma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit(&tmem_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------- epilogue warps
    // eight epilogue warps: two warpgroups share the four TMEM lane quadrants (warp w may only touch lanes
    // 32*(w%4)..+31); warpgroup 0 drains the even 32-column chunks of every tile, warpgroup 1 the odd ones
    const int ew = (warp - 4) & 3;   // TMEM lane quadrant = 32-row slab of the tile
    const int wg = (warp - 4) >> 2;  // 0 / 1
    const int et = ew * 32 + lane;   // thread index inside the warpgroup
    const int row = ew * 32 + lane;
    uint8_t* stage_buf = epi_smem + (warp - 4) * EPI_BUF_BYTES;
    if (f_pool) {  // pooled mode never stores through the staging buffers: they hold the cell maxima
      for (int i = threadIdx.x - 128; i < 8192; i += 256) reinterpret_cast<int*>(epi_smem)[i] = ENC_NEG;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const bool row_vec = ((reinterpret_cast<uintptr_t>(p.residual) | reinterpret_cast<uintptr_t>(p.out)) & 15) == 0 &&
                         p.ldr % 4 == 0 && p.ldo % 4 == 0;
    int it = 0;
    int cur_b = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int n_tile = (tile / f_splits) % num_n_tiles;
      const int m_group = tile / (f_splits * num_n_tiles);
      const int n0 = n_tile * BN;
      mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
      const int m_tile = m_group * MT + sub;
      if (m_tile >= num_m_tiles) continue;  // odd tail of a paired tile (warp-uniform)
      // ---- output row of this thread
      long long m = -1;
      int tile_b = 0, tile_x0 = 0, tile_y0 = 0, tile_z0 = 0;
      if (f_conv) {
        int t = m_tile;
        const int tz = t % p.tiles_z; t /= p.tiles_z;
        const int ty = t % p.tiles_y; t /= p.tiles_y;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        tile_b = t;
        tile_x0 = tx * p.bx; tile_y0 = ty * p.by; tile_z0 = tz * p.bz;
        const int dz = row % p.bz;
        const int dy = (row / p.bz) % p.by;
        const int dx = row / (p.bz * p.by);
        const int x = tx * p.bx + dx, y = ty * p.by + dy, z = tz * p.bz + dz;
        if (x < p.Xo && y < p.Yo && z < p.Zo)
          m = (((long long)tile_b * p.Xo + x) * p.Yo + y) * p.Zo + z;
      } else {
        const long long mm = (long long)m_tile * BM + row;
        if (mm < p.M) m = mm;
        if (p.rows_per_batch > 0) tile_b = (int)(((long long)m_tile * BM) / p.rows_per_batch);
      }
      const bool valid = m >= 0;
      // pooled mode: this row's cell inside the tile and the (cell, column) slots this thread drains, once per tile
      int pool_cl = 0, pool_n = 0;
      long long pool_dst[8];
      if (f_pool) {
        const int dz = row % p.bz, dy = (row / p.bz) % p.by, dx = row / (p.bz * p.by);
        pool_cl = ((dx / p.pool_cwx) * p.pool_ncy + dy / p.pool_cwy) * p.pool_ncz + dz / p.pool_cwz;
        const int ncell = (p.bx / p.pool_cwx) * p.pool_ncy * p.pool_ncz;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pool_dst[e] = -1;
          const int i = et + e * 128;
          if (i < ncell * 32) {
            pool_n = e + 1;
            const int cl = i >> 5;
            const int cz = cl % p.pool_ncz, cy = (cl / p.pool_ncz) % p.pool_ncy, cx = cl / (p.pool_ncz * p.pool_ncy);
            const int vx = tile_x0 + cx * p.pool_cwx, vy = tile_y0 + cy * p.pool_cwy, vz = tile_z0 + cz * p.pool_cwz;
            if (vx < p.Xo && vy < p.Yo && vz < p.Zo)
              pool_dst[e] = ((((long long)tile_b * p.pXo + vx / p.pwx) * p.pYo + vy / p.pwy) * p.pZo + vz / p.pwz) * p.N;
          }
        }
      }

      if (f_stats && tile_b != cur_b) {
        // flush the per-CTA fp64 partial sums of the previous batch sample
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (cur_b >= 0) {
          const int ngroups2 = 2 * (p.N / p.cpg);
          for (int i = threadIdx.x - 128; i < ngroups2; i += 256) {
            const double v = stat_acc[i];
            if (v != 0.0) atomicAdd(&p.gn_stats[(size_t)cur_b * ngroups2 + i], v);
            stat_acc[i] = 0.0;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_b = tile_b;
      }

      if (f_splits > 1) {  // wait until the lower splits of this tile have added their partial sums
        if (threadIdx.x == 128) {
          const int want = tile % f_splits;
          volatile int* sem = p.splitk_sem + ((tile / f_splits) & (SPLITK_SEMS - 1));
          for (int spin = 0; *sem != want && spin < (1 << 16); ++spin) __nanosleep(64);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + buf * ACC_COLS + sub * BN;
      // software-pipelined TMEM reads: the load of this warp's next chunk (c+2) is in flight while chunk c is processed
      uint32_t rnext[32];
      if (n0 + wg * 32 < p.N && wg < BN / 32) tmem_ld_32x32(t_row + wg * 32, rnext);
#pragma unroll 1
      for (int c = wg; c < BN / 32; c += 2) {
        const int nc = n0 + c * 32;
        if (nc >= p.N) break;
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rnext[j]);
        if (c + 2 < BN / 32 && nc + 64 < p.N) tmem_ld_32x32(t_row + (c + 2) * 32, rnext);
        if (f_stats) {
          // per-group sum / sumsq of the raw conv output, butterfly-reduced over the 32 rows of the warp
          const int cpg = p.cpg;  // power of two in [1, 32]
          if (cpg == 1) {
            // one group per channel: 32 sums and 32 sums of squares -> two butterflies
            float ss[32], sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              ss[j] = valid ? v[j] : 0.f;
              sq[j] = valid ? v[j] * v[j] : 0.f;
            }
            warp_butterfly32(ss, lane);
            warp_butterfly32(sq, lane);
            if (nc + lane < p.N) {
              atomicAdd(&stat_acc[2 * (nc + lane)], (double)ss[0]);
              atomicAdd(&stat_acc[2 * (nc + lane) + 1], (double)sq[0]);
            }
          } else {
            float sv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sv[j] = 0.f;
            if (valid) {
              if (cpg == 2) accum_stats<2>(v, sv);
              else if (cpg == 4) accum_stats<4>(v, sv);
              else if (cpg == 8) accum_stats<8>(v, sv);
              else if (cpg == 16) accum_stats<16>(v, sv);
              else accum_stats<32>(v, sv);
            }
            warp_butterfly32(sv, lane);
            // lane L now holds the warp total of value L (value = 2*local_group + {0: sum, 1: sumsq})
            if (lane < 2 * (32 / cpg)) atomicAdd(&stat_acc[2 * (nc / cpg) + lane], (double)sv[0]);
          }
        }
        // ---- epilogue math in the thread = row domain (32 consecutive output columns in registers)
        const bool full_chunk = LEAN || (nc + 32 <= p.N);
        if (p.bias) {
          if (full_chunk && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bq = __ldg(reinterpret_cast<const float4*>(p.bias + nc + j));
              v[j] += bq.x; v[j + 1] += bq.y; v[j + 2] += bq.z; v[j + 3] += bq.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.N) v[j] += __ldg(p.bias + nc + j);
          }
        }
        if (p.residual && valid) {
          const float* rrow = p.residual + m * p.ldr + nc;
          if (full_chunk && row_vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 rr = __ldg(reinterpret_cast<const float4*>(rrow + j));
              v[j] += rr.x; v[j + 1] += rr.y; v[j + 2] += rr.z; v[j + 3] += rr.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.N) v[j] += rrow[j];
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (p.split_out) {  // the 32 values of this (row, chunk) become the 32 words of their S32 chunk
          uint32_t w[32];
          split_chunk32(v, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(w[j]);
        }
        if (f_pool) {
          // ---- fused adaptive max pool (windows are powers of two >= 2 that divide the grid).  The G lanes of this
          // warp's (ex, ey, ez) voxel sub-box that share a pooling cell reduce their 32 columns with a halving
          // butterfly (lane keeps 32/G column maxima: 32 - 32/G shuffles instead of 32 log2 G), the partial maxima of
          // the four warps meet through smem atomicMax, then one global atomicMax per (cell of the tile, column).
          // Smem buffers alternate per chunk: one named barrier per chunk.
          int* ps = reinterpret_cast<int*>(epi_smem) + (wg * 2 + ((c >> 1) & 1)) * 2048;  // <= 32 cells x 32 columns
          {
            float w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = valid ? v[j] : -INFINITY;
            int qoff = 0;  // first column (within the chunk) of the values this lane ends up holding
            int n = 32;
#pragma unroll
            for (int st = 0; st < 5; ++st) {
              if (st < p.pool_nsteps) {  // warp-uniform
                const int o = p.pool_stride[st];
                const bool upper = (lane & o) != 0;
                const int half = 32 >> (st + 1);
#pragma unroll
                for (int i = 0; i < half; ++i) {
                  const float keep = upper ? w[i + half] : w[i];
                  const float send = upper ? w[i] : w[i + half];
                  w[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, o));
                }
                qoff += upper ? half : 0;
                n = half;
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)  // G >= 4 lanes per cell in every supported configuration -> n <= 8
              if (i < n && nc + qoff + i < p.N && w[i] > -INFINITY)
                atomicMax(&ps[pool_cl * 32 + qoff + i], enc_ordered(w[i]));
          }
          asm volatile("bar.sync %0, 128;" ::"r"(2 + wg) : "memory");  // the four warps of this warpgroup
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (e < pool_n) {
              const int i = et + e * 128;
              const int val = ps[i];
              ps[i] = ENC_NEG;
              const int q = nc + (i & 31);
              if (val != ENC_NEG && q < p.N && pool_dst[e] >= 0) {
                int* dst = p.pool_out + pool_dst[e] + q;
                if (p.pool_complete) *dst = val; else atomicMax(dst, val);
                if (val >= 0) p.pool_flag[(size_t)tile_b * p.N + q] = 1;
              }
            }
          }
        }
        if (!f_store) {
          // nothing else to write
        } else if (f_tma) {
          // registers -> 128B-swizzled smem chunk (conflict-free) -> one TMA store per warp and chunk; the TMA unit
          // generates the row addresses and clips rows >= M / columns >= N / voxels outside the grid
          uint8_t* sb = stage_buf;
          if (lane == 0) tma_store_wait_read<0>();  // this warp's previous chunk has been read out of the buffer
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(sb + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (f_conv) {
              // rows ew*32 .. ew*32+31 of the tile form a sub-box (ex, ey, ez) of the (bx, by, bz) voxel box
              const int r0 = ew * 32;
              if (f_splits > 1)
                tma_reduce_add_5d(&tmC, sb, nc, tile_z0 + r0 % p.bz, tile_y0 + (r0 / p.bz) % p.by,
                                  tile_x0 + r0 / (p.bz * p.by), tile_b);
              else
                tma_store_5d(&tmC, sb, nc, tile_z0 + r0 % p.bz, tile_y0 + (r0 / p.bz) % p.by,
                             tile_x0 + r0 / (p.bz * p.by), tile_b);
            } else if (f_splits > 1) {
              tma_reduce_add_2d(&tmC, sb, nc, m_tile * BM + ew * 32);
            } else {
              tma_store_2d(&tmC, sb, nc, m_tile * BM + ew * 32);
            }
            tma_store_commit();
          }
        } else if (valid) {
          float* orow = p.out + m * p.ldo + nc;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nc + j < p.N) orow[j] = v[j];
        }
      }
      }  // sub
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      if (f_splits > 1) {  // publish: this split's reduce-adds are complete
        if (lane == 0) tma_store_wait_all();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 128) {
          __threadfence();
          const int split = tile % f_splits;
          atomicExch(p.splitk_sem + ((tile / f_splits) & (SPLITK_SEMS - 1)), split == f_splits - 1 ? 0 : split + 1);
        }
      }
    }
    if (f_stats) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (cur_b >= 0) {
        const int ngroups2 = 2 * (p.N / p.cpg);
        for (int i = threadIdx.x - 128; i < ngroups2; i += 256) {
          const double v = stat_acc[i];
          if (v != 0.0) atomicAdd(&p.gn_stats[(size_t)cur_b * ngroups2 + i], v);
        }
      }
    }
    if (lane == 0) tma_store_wait_all();  // all bulk stores of this warp have been written
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace occ

// =====================================================================================================
// host side
// =====================================================================================================
namespace occ {

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// GroupNorm statistics of a finished (B, rows_per_batch, C) conv output (split-K path): stats[b][g] += (sum, sumsq).
// One CTA = 32 rows; thread t owns the float4 column t (t + 256, ...): eight independent row loads in flight, fp32 sums
// over the 32 rows, fp64 per group in shared memory, one fp64 global atomic per (CTA, group, moment).
constexpr int GS_ROWS = 32;
__global__ void __launch_bounds__(256)
conv_gn_stats_kernel(const float* __restrict__ out, double* __restrict__ stats, int rows_per_batch, int C, int cpg) {
  __shared__ double sg[2 * MAX_STAT_GROUPS];
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * GS_ROWS;
  const int nr = min(GS_ROWS, rows_per_batch - r0);
  const int groups = C / cpg;
  const int C4 = C >> 2;
  if (threadIdx.x < 2 * MAX_STAT_GROUPS) sg[threadIdx.x] = 0.0;
  __syncthreads();
  const float4* base = reinterpret_cast<const float4*>(out + ((size_t)b * rows_per_batch + r0) * C);
  for (int c4 = threadIdx.x; c4 < C4; c4 += 256) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int r = 0; r < nr; ++r) {
      const float4 v = __ldg(base + (size_t)r * C4 + c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
    const int c = c4 << 2;
    if (cpg >= 4) {  // the four columns share a group (cpg is a power of two)
      atomicAdd(&sg[2 * (c / cpg)], (double)((s.x + s.y) + (s.z + s.w)));
      atomicAdd(&sg[2 * (c / cpg) + 1], (double)((q.x + q.y) + (q.z + q.w)));
    } else {
      const float sv[4] = {s.x, s.y, s.z, s.w}, qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&sg[2 * ((c + j) / cpg)], (double)sv[j]);
        atomicAdd(&sg[2 * ((c + j) / cpg) + 1], (double)qv[j]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * groups) atomicAdd(&stats[(size_t)b * 2 * groups + threadIdx.x], sg[threadIdx.x]);
}

template <int BN, int STAGES, int MT = 1, bool LEAN = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                       int num_tiles, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (MT * A_STAGE_BYTES + BN * BK * 4) + 1024 /*align*/ + CTRL_BYTES + EPI_BYTES;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  OCC_ENSURE_SMEM((gemm_bf16x3_kernel<BN, STAGES, MT, LEAN>), smem);
  int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  if (grid < 1) grid = 1;
  gemm_bf16x3_kernel<BN, STAGES, MT, LEAN><<<grid, 384, smem, stream>>>(tmA, tmB, tmC, p);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}

static int dispatch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmC, const void* W, GemmParams& p, int num_m_tiles,
                         cudaStream_t stream) {
  // N-tile: 192-wide tiles for the 192-channel neck / head matrices (N = 192, 576: no padded columns), else 128 / 256
  int BN = p.N <= 32 ? 32 : p.N <= 64 ? 64 : (p.N <= 128 || (p.N % 256 != 0 && p.N % 128 == 0)) ? 128
           : (p.N % 192 == 0 && p.N % 256 != 0) ? 192 : 256;
  CUtensorMap tmB;
  uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
  uint64_t strides[1] = {(uint64_t)p.K * 4};
  uint32_t box[2] = {(uint32_t)BK, (uint32_t)BN};
  int rc = make_tmap_f32(&tmB, W, 2, dims, strides, box, nullptr);
  if (rc) return rc;
  if (p.splits < 1) p.splits = 1;
  const int num_tiles = num_m_tiles * ((p.N + BN - 1) / BN) * p.splits;
  // stage-0 convs (N = 128, thousands of M-tiles): pairs of M-tiles share every weight k-block (MT = 2)
  if (BN == 128 && p.conv && p.splits == 1 && p.pool_out == nullptr && p.use_tma_store &&
      num_m_tiles >= 4 * sm_count())
    return launch_gemm<128, 4, 2>(tmA, tmB, tmC, p, ((num_m_tiles + 1) / 2) * ((p.N + BN - 1) / BN), stream);
  // plain Linear with whole 32-column chunks through the TMA store: the lean instantiation (shorter epilogue)
  const bool lean = !p.conv && p.splits == 1 && p.pool_out == nullptr && p.gn_stats == nullptr && p.use_tma_store &&
                    p.store_out && p.N % 32 == 0 && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0 &&
                    (p.residual == nullptr || ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0 && p.ldr % 4 == 0));
  if (lean) {
    switch (BN) {
      case 32: break;
      case 64: return launch_gemm<64, 8, 1, true>(tmA, tmB, tmC, p, num_tiles, stream);
      case 128: return launch_gemm<128, 5, 1, true>(tmA, tmB, tmC, p, num_tiles, stream);
      case 192: return launch_gemm<192, 4, 1, true>(tmA, tmB, tmC, p, num_tiles, stream);
      default: return launch_gemm<256, 4, 1, true>(tmA, tmB, tmC, p, num_tiles, stream);
    }
  }
  switch (BN) {
    case 32: return launch_gemm<32, 8>(tmA, tmB, tmC, p, num_tiles, stream);
    case 64: return launch_gemm<64, 8>(tmA, tmB, tmC, p, num_tiles, stream);
    case 128: return launch_gemm<128, 5>(tmA, tmB, tmC, p, num_tiles, stream);
    case 192: return launch_gemm<192, 4>(tmA, tmB, tmC, p, num_tiles, stream);
    default: return launch_gemm<256, 4>(tmA, tmB, tmC, p, num_tiles, stream);
  }
}

}  // namespace occ

using namespace occ;

extern "C" int occ_gemm_bf16x3(const float* A, const float* W, float* out, int M, int N, int K, const float* bias,
                               const float* residual, int act, int split_out, double* gn_stats, int cpg,
                               int rows_per_batch, cudaStream_t stream) {
  OCC_REQUIRE(A && W && out);
  OCC_REQUIRE(M > 0 && N > 0 && K > 0);
  OCC_REQUIRE(K % 32 == 0);                 // S32 operand rows: whole 32-k chunks
  OCC_REQUIRE(!split_out || N % 32 == 0);   // S32 output rows
  OCC_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  OCC_REQUIRE(act >= 0 && act <= 2);
  if (gn_stats) OCC_REQUIRE(cpg >= 1 && cpg <= 32 && (cpg & (cpg - 1)) == 0 && N % cpg == 0 && N / cpg <= MAX_STAT_GROUPS &&
                            rows_per_batch > 0 && rows_per_batch % BM == 0);
  GemmParams p{};
  p.passes = mma_passes();
  p.M = M; p.N = N; p.K = K; p.num_k_blocks = (K + BK - 1) / BK;
  p.out = out; p.ldo = N; p.bias = bias; p.residual = residual; p.ldr = N; p.act = act; p.split_out = split_out;
  p.conv = 0; p.gn_stats = gn_stats; p.cpg = cpg; p.rows_per_batch = gn_stats ? rows_per_batch : 0;
  p.pool_out = nullptr; p.pool_flag = nullptr; p.store_out = 1;
  CUtensorMap tmA;
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t strides[1] = {(uint64_t)K * 4};
  uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
  int rc = make_tmap_f32(&tmA, A, 2, dims, strides, box, nullptr);
  if (rc) return rc;
  CUtensorMap tmC = tmA;  // placeholder when the TMA-store path is not usable (N % 4 != 0: scalar row stores)
  p.use_tma_store = (N % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (p.use_tma_store) {
    uint64_t cd[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t cs[1] = {(uint64_t)N * 4};
    uint32_t cb[2] = {32u, 32u};
    rc = make_tmap_f32(&tmC, out, 2, cd, cs, cb, nullptr);
    if (rc) return rc;
  }
  return dispatch_gemm(tmA, tmC, W, p, (M + BM - 1) / BM, stream);
}

// x: (B, X, Y, Z, Cin) channel-last, S32;  w2: (Cout, KX*KY*KZ*Cin) tap-major repacked weights, S32 per tap;
// out: (B, Xo, Yo, Zo, Cout).  pad = dil*(K-1)/2 per axis ("same" for stride 1), Xo = (X + 2p - dil*(K-1) - 1)/s + 1.
extern "C" size_t occ_conv_workspace_bytes(void) { return SPLITK_SEMS * sizeof(int); }

extern "C" int occ_conv_bf16x3(const float* x, const float* w2, float* out, int B, int X, int Y, int Z, int Cin,
                               int Cout, int KX, int KY, int KZ, int stride, int dil, const float* bias,
                               const float* residual, int act, int split_out, double* gn_stats, int cpg,
                               void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  OCC_REQUIRE(x && w2 && out);
  OCC_REQUIRE(!split_out || Cout % 32 == 0);
  OCC_REQUIRE(B > 0 && X > 0 && Y > 0 && Z > 0 && Cin > 0 && Cout > 0);
  OCC_REQUIRE(Cin % BK == 0);
  OCC_REQUIRE((KX == 1 || KX == 3) && (KY == 1 || KY == 3) && (KZ == 1 || KZ == 3));
  OCC_REQUIRE(stride == 1 || stride == 2);
  OCC_REQUIRE(dil >= 1 && act >= 0 && act <= 2);
  OCC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0);
  if (gn_stats) OCC_REQUIRE(cpg >= 1 && cpg <= 32 && (cpg & (cpg - 1)) == 0 && Cout % cpg == 0 && Cout / cpg <= MAX_STAT_GROUPS);
  GemmParams p{};
  p.passes = mma_passes();
  p.conv = 1;
  p.Cin = Cin; p.KX = KX; p.KY = KY; p.KZ = KZ; p.dil = dil; p.stride = stride;
  p.padx = dil * (KX - 1) / 2; p.pady = dil * (KY - 1) / 2; p.padz = dil * (KZ - 1) / 2;
  p.B = B;
  p.Xo = (X + 2 * p.padx - dil * (KX - 1) - 1) / stride + 1;
  p.Yo = (Y + 2 * p.pady - dil * (KY - 1) - 1) / stride + 1;
  p.Zo = (Z + 2 * p.padz - dil * (KZ - 1) - 1) / stride + 1;
  p.bz = next_pow2(p.Zo) < BM ? next_pow2(p.Zo) : BM;
  p.by = next_pow2(p.Yo) < BM / p.bz ? next_pow2(p.Yo) : BM / p.bz;
  p.bx = BM / (p.bz * p.by);
  p.tiles_x = (p.Xo + p.bx - 1) / p.bx;
  p.tiles_y = (p.Yo + p.by - 1) / p.by;
  p.tiles_z = (p.Zo + p.bz - 1) / p.bz;
  p.N = Cout; p.K = KX * KY * KZ * Cin; p.num_k_blocks = p.K / BK;
  p.M = B * p.Xo * p.Yo * p.Zo;
  p.out = out; p.ldo = Cout; p.bias = bias; p.residual = residual; p.ldr = Cout; p.act = act; p.split_out = split_out;
  p.gn_stats = gn_stats; p.cpg = cpg; p.rows_per_batch = 0;
  p.pool_out = nullptr; p.pool_flag = nullptr; p.store_out = 1;
  OCC_REQUIRE(p.bx * stride <= 256 && p.by * stride <= 256 && p.bz * stride <= 256);
  CUtensorMap tmA;
  uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)Z, (uint64_t)Y, (uint64_t)X, (uint64_t)B};
  uint64_t strides[4] = {(uint64_t)Cin * 4, (uint64_t)Z * Cin * 4, (uint64_t)Y * Z * Cin * 4,
                         (uint64_t)X * Y * Z * Cin * 4};
  uint32_t box[5] = {(uint32_t)BK, (uint32_t)(p.bz * stride), (uint32_t)(p.by * stride), (uint32_t)(p.bx * stride), 1};
  uint32_t estr[5] = {1, (uint32_t)stride, (uint32_t)stride, (uint32_t)stride, 1};
  int rc = make_tmap_f32(&tmA, x, 5, dims, strides, box, estr);
  if (rc) return rc;
  CUtensorMap tmC = tmA;
  p.use_tma_store = (Cout % 4 == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (p.use_tma_store) {
    // one epilogue warp owns 32 consecutive rows of the (bx, by, bz) box = a sub-box (ex, ey, ez)
    const int ez = p.bz < 32 ? p.bz : 32;
    const int ey = p.by < 32 / ez ? p.by : 32 / ez;
    const int ex = 32 / (ez * ey);
    uint64_t cd[5] = {(uint64_t)Cout, (uint64_t)p.Zo, (uint64_t)p.Yo, (uint64_t)p.Xo, (uint64_t)B};
    uint64_t cs[4] = {(uint64_t)Cout * 4, (uint64_t)p.Zo * Cout * 4, (uint64_t)p.Yo * p.Zo * Cout * 4,
                      (uint64_t)p.Xo * p.Yo * p.Zo * Cout * 4};
    uint32_t cb[5] = {32u, (uint32_t)ez, (uint32_t)ey, (uint32_t)ex, 1u};
    rc = make_tmap_f32(&tmC, out, 5, cd, cs, cb, nullptr);
    if (rc) return rc;
  }
  const int num_m_tiles = B * p.tiles_x * p.tiles_y * p.tiles_z;
  // Split-K for the deep stages (few output tiles, long K = 27 * Cin): a 512->512 conv on 50x50x4 voxels has 158 tiles
  // for 148 SMs (second wave 7 % full), a 1024->1024 conv on 25x25x2 has 40.  The raw conv output of the encoder has
  // no bias / activation (GroupNorm follows), so the partial sums can meet in `out` through TMA reduce-add; the
  // GroupNorm statistics are then taken from the finished tensor.
  p.splits = 1;
  if (p.use_tma_store && !bias && !residual && act == 0 && !split_out && workspace &&
      workspace_bytes >= SPLITK_SEMS * sizeof(int) && (reinterpret_cast<uintptr_t>(workspace) & 3) == 0) {
    const int bn = Cout <= 32 ? 32 : Cout <= 64 ? 64 : (Cout <= 128 || (Cout % 256 != 0 && Cout % 128 == 0)) ? 128
                   : (Cout % 192 == 0 && Cout % 256 != 0) ? 192 : 256;
    const long long tiles = (long long)num_m_tiles * ((Cout + bn - 1) / bn);
    const int sms = sm_count();
    auto eff = [&](int s) { const long long t = tiles * s; return (double)t / (double)(((t + sms - 1) / sms) * sms); };
    // (not below 8 M-tiles = 1024 output rows: such convs are latency-bound either way, and keeping them on the
    // single-pass path keeps small problems bit-reproducible -- the reduce-add order of the splits is not fixed)
    if (num_m_tiles >= 8 && tiles < 2 * sms && p.num_k_blocks >= 64 && eff(1) < 0.85) {
      int best = 1;
      for (int sp = 2; sp <= 8 && sp * 8 <= p.num_k_blocks; ++sp) {
        if (eff(sp) > eff(best) + 0.02) best = sp;
        if (eff(best) >= 0.85) break;
      }
      p.splits = best;
    }
  }
  if (p.splits > 1) {
    OCC_CUDA(cudaMemsetAsync(out, 0, (size_t)p.M * Cout * sizeof(float), stream));
    p.gn_stats = nullptr;
    p.splitk_sem = static_cast<int*>(workspace);
  }
  rc = dispatch_gemm(tmA, tmC, w2, p, num_m_tiles, stream);
  if (rc) return rc;
  if (p.splits > 1 && gn_stats) {
    const int rows_per_batch = p.Xo * p.Yo * p.Zo;
    dim3 grid((rows_per_batch + GS_ROWS - 1) / GS_ROWS, B);
    conv_gn_stats_kernel<<<grid, 256, 0, stream>>>(out, gn_stats, rows_per_batch, Cout, cpg);
    OCC_LAUNCH_CHECK();
  }
  return OCC_OK;
}


// Several same-input stride-1 convolutions in one launch: an explicit tap table instead of a regular stencil.
// x (B,X,Y,Z,Cin) S32; taps: ntaps x (dx, dy, dz) input offsets (HOST ints, |offset| <= 127); w2 (Cout, ntaps*Cin) S32,
// tap-major in the table's order (a branch that does not use a tap has zero weights there); out (B,X,Y,Z,Cout) raw fp32;
// gn_stats as occ_conv_bf16x3 (up to 64 groups).  Used for the four ASPP branches (1x1 + three dilated 3x3,
// P/occformer/backbones/modules/aspp.py:107-113): one launch with 25 taps instead of four.
extern "C" int occ_conv_taps_bf16x3(const float* x, const float* w2, float* out, int B, int X, int Y, int Z, int Cin,
                                    int Cout, int ntaps, const int* taps, double* gn_stats, int cpg, cudaStream_t stream) {
  OCC_REQUIRE(x && w2 && out && taps);
  OCC_REQUIRE(B > 0 && X > 0 && Y > 0 && Z > 0 && Cin > 0 && Cout > 0 && Cin % BK == 0 && ntaps >= 1 && ntaps <= MAX_TAPS);
  OCC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(out) & 15) == 0 && Cout % 4 == 0);
  if (gn_stats) OCC_REQUIRE(cpg >= 1 && cpg <= 32 && (cpg & (cpg - 1)) == 0 && Cout % cpg == 0 && Cout / cpg <= MAX_STAT_GROUPS);
  GemmParams p{};
  p.passes = mma_passes();
  p.conv = 1;
  p.Cin = Cin; p.KX = p.KY = p.KZ = 1; p.dil = 1; p.stride = 1; p.padx = p.pady = p.padz = 0;
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    OCC_REQUIRE(taps[3 * t] >= -127 && taps[3 * t] <= 127 && taps[3 * t + 1] >= -127 && taps[3 * t + 1] <= 127 &&
                taps[3 * t + 2] >= -127 && taps[3 * t + 2] <= 127);
    p.tdx[t] = (signed char)taps[3 * t]; p.tdy[t] = (signed char)taps[3 * t + 1]; p.tdz[t] = (signed char)taps[3 * t + 2];
  }
  p.B = B; p.Xo = X; p.Yo = Y; p.Zo = Z;
  p.bz = next_pow2(Z) < BM ? next_pow2(Z) : BM;
  p.by = next_pow2(Y) < BM / p.bz ? next_pow2(Y) : BM / p.bz;
  p.bx = BM / (p.bz * p.by);
  p.tiles_x = (X + p.bx - 1) / p.bx; p.tiles_y = (Y + p.by - 1) / p.by; p.tiles_z = (Z + p.bz - 1) / p.bz;
  p.N = Cout; p.K = ntaps * Cin; p.num_k_blocks = p.K / BK; p.M = B * X * Y * Z;
  p.out = out; p.ldo = Cout; p.ldr = Cout; p.act = 0; p.split_out = 0;
  p.gn_stats = gn_stats; p.cpg = cpg; p.rows_per_batch = 0; p.store_out = 1; p.splits = 1;
  CUtensorMap tmA;
  uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)Z, (uint64_t)Y, (uint64_t)X, (uint64_t)B};
  uint64_t strides[4] = {(uint64_t)Cin * 4, (uint64_t)Z * Cin * 4, (uint64_t)Y * Z * Cin * 4, (uint64_t)X * Y * Z * Cin * 4};
  uint32_t box[5] = {(uint32_t)BK, (uint32_t)p.bz, (uint32_t)p.by, (uint32_t)p.bx, 1};
  int rc = make_tmap_f32(&tmA, x, 5, dims, strides, box, nullptr);
  if (rc) return rc;
  const int ez = p.bz < 32 ? p.bz : 32;
  const int ey = p.by < 32 / ez ? p.by : 32 / ez;
  const int ex = 32 / (ez * ey);
  CUtensorMap tmC;
  uint64_t cd[5] = {(uint64_t)Cout, (uint64_t)Z, (uint64_t)Y, (uint64_t)X, (uint64_t)B};
  uint64_t cs[4] = {(uint64_t)Cout * 4, (uint64_t)Z * Cout * 4, (uint64_t)Y * Z * Cout * 4, (uint64_t)X * Y * Z * Cout * 4};
  uint32_t cb[5] = {32u, (uint32_t)ez, (uint32_t)ey, (uint32_t)ex, 1u};
  rc = make_tmap_f32(&tmC, out, 5, cd, cs, cb, nullptr);
  if (rc) return rc;
  p.use_tma_store = 1;
  return dispatch_gemm(tmA, tmC, w2, p, B * p.tiles_x * p.tiles_y * p.tiles_z, stream);
}

// Decoder mask logits with fused attention-mask pooling (mask2former_nusc_occ.py:457-466):
//   mask[b, v, q] = <mask_feature[b, v, :], mask_embed[b, q, :]>   (1x1x1 "conv" over the voxel grid, K = E)
//   pooled[b, cell, q] = max over the (X/Xo, Y/Yo, Z/Zo) voxels of the cell  (adaptive_max_pool3d with divisible,
//   power-of-two windows), written as order-preserving ints (sign == sign of the logit: blocked <=> value < 0);
//   flag[b, q] = 1 iff some pooled value of the row is >= 0.   mask_out may be NULL (intermediate decoder layers).
namespace occ {
int mask_pool_query_stationary(const float* mf, const float* membed, int* pooled, int* flag, int B, int X, int Y, int Z,
                               int E, int Q, int Xo, int Yo, int Zo, cudaStream_t stream);
}

extern "C" int occ_mask_gemm_pool(const float* mf, const float* membed, float* mask_out, int* pooled, int* flag, int B,
                                  int X, int Y, int Z, int E, int Q, int Xo, int Yo, int Zo, cudaStream_t stream) {
  OCC_REQUIRE(mf && membed && pooled && flag);
  OCC_REQUIRE(B > 0 && X > 0 && Y > 0 && Z > 0 && E % BK == 0 && Q > 0 && Q % 4 == 0 && Q <= 128);
  OCC_REQUIRE(Xo > 0 && Yo > 0 && Zo > 0 && X % Xo == 0 && Y % Yo == 0 && Z % Zo == 0);
  const int wx = X / Xo, wy = Y / Yo, wz = Z / Zo;
  auto pow2 = [](int v) { return v >= 2 && (v & (v - 1)) == 0; };
  OCC_REQUIRE(pow2(wx) && pow2(wy) && pow2(wz));
  OCC_REQUIRE((reinterpret_cast<uintptr_t>(mf) & 15) == 0 && (reinterpret_cast<uintptr_t>(membed) & 15) == 0);
  if (!mask_out) {  // only the pooled logits are wanted: query-stationary kernel (mask_pool_tc.cu), register pooling
    const int rc = mask_pool_query_stationary(mf, membed, pooled, flag, B, X, Y, Z, E, Q, Xo, Yo, Zo, stream);
    if (rc != 1) return rc;
  }
  OCC_CUDA(cudaMemsetAsync(pooled, 0x80, (size_t)B * Xo * Yo * Zo * Q * sizeof(int), stream));
  OCC_CUDA(cudaMemsetAsync(flag, 0, (size_t)B * Q * sizeof(int), stream));
  for (int b = 0; b < B; ++b) {
    GemmParams p{};
    p.passes = mma_passes();
    p.conv = 1;
    p.Cin = E; p.KX = p.KY = p.KZ = 1; p.dil = 1; p.stride = 1; p.padx = p.pady = p.padz = 0;
    p.B = 1; p.Xo = X; p.Yo = Y; p.Zo = Z;
    // voxel box of a tile: prefer a box that contains whole pooling cells (their maxima are then complete inside the
    // CTA and are written with plain coalesced stores instead of global atomics)
    p.bz = next_pow2(Z) < BM ? next_pow2(Z) : BM;
    p.by = next_pow2(Y) < BM / p.bz ? next_pow2(Y) : BM / p.bz;
    p.bx = BM / (p.bz * p.by);
    for (int cz = (next_pow2(Z) < 16 ? next_pow2(Z) : 16); cz >= wz; cz >>= 1) {
      if (cz * wy * wx > BM) continue;
      int cy = wy, cx = wx;
      while (cz * cy * cx < BM) {  // grow y / x alternately (powers of two stay multiples of the window)
        if (cy <= cx) cy *= 2; else cx *= 2;
      }
      p.bz = cz; p.by = cy; p.bx = cx;
      break;
    }
    p.pool_complete = (p.bx % wx == 0 && p.by % wy == 0 && p.bz % wz == 0) ? 1 : 0;
    p.tiles_x = (X + p.bx - 1) / p.bx; p.tiles_y = (Y + p.by - 1) / p.by; p.tiles_z = (Z + p.bz - 1) / p.bz;
    p.N = Q; p.K = E; p.num_k_blocks = E / BK; p.M = X * Y * Z;
    float* out_b = mask_out ? mask_out + (size_t)b * X * Y * Z * Q : nullptr;
    p.out = out_b; p.ldo = Q; p.bias = nullptr; p.residual = nullptr; p.ldr = Q; p.act = 0; p.split_out = 0;
    p.gn_stats = nullptr; p.cpg = 0; p.rows_per_batch = 0;
    p.pool_out = pooled + (size_t)b * Xo * Yo * Zo * Q; p.pool_flag = flag + (size_t)b * Q;
    p.pwx = wx; p.pwy = wy; p.pwz = wz; p.pXo = Xo; p.pYo = Yo; p.pZo = Zo;
    p.store_out = out_b != nullptr;
    // cells of one tile: (bx/min(wx,bx)) * (by/min(wy,by)) * (bz/min(wz,bz)) <= 64 (smem pool buffers)
    const int cwx = wx < p.bx ? wx : p.bx, cwy = wy < p.by ? wy : p.by, cwz = wz < p.bz ? wz : p.bz;
    OCC_REQUIRE((p.bx / cwx) * (p.by / cwy) * (p.bz / cwz) <= 32);  // <= 1024 (cell, column) slots: 8 per epilogue thread
    p.pool_cwx = cwx; p.pool_cwy = cwy; p.pool_cwz = cwz; p.pool_ncy = p.by / cwy; p.pool_ncz = p.bz / cwz;
    {  // in-warp butterfly: the warp's 32 rows are the (ex, ey, ez) sub-box, lane = (lx*ey + ly)*ez + lz
      const int ez = p.bz < 32 ? p.bz : 32;
      const int ey = p.by < 32 / ez ? p.by : 32 / ez;
      const int ex = 32 / (ez * ey);
      const int rz = wz < ez ? wz : ez, ry = wy < ey ? wy : ey, rx = wx < ex ? wx : ex;
      int n = 0;
      for (int o = 1; o < rz; o <<= 1) p.pool_stride[n++] = o;
      for (int o = 1; o < ry; o <<= 1) p.pool_stride[n++] = o * ez;
      for (int o = 1; o < rx; o <<= 1) p.pool_stride[n++] = o * ez * ey;
      p.pool_nsteps = n;
      OCC_REQUIRE(n >= 2 && n <= 5);  // >= 4 lanes per cell
    }
    const float* x = mf + (size_t)b * X * Y * Z * E;
    CUtensorMap tmA;
    uint64_t dims[5] = {(uint64_t)E, (uint64_t)Z, (uint64_t)Y, (uint64_t)X, 1};
    uint64_t strides[4] = {(uint64_t)E * 4, (uint64_t)Z * E * 4, (uint64_t)Y * Z * E * 4, (uint64_t)X * Y * Z * E * 4};
    uint32_t box[5] = {(uint32_t)BK, (uint32_t)p.bz, (uint32_t)p.by, (uint32_t)p.bx, 1};
    int rc = make_tmap_f32(&tmA, x, 5, dims, strides, box, nullptr);
    if (rc) return rc;
    CUtensorMap tmC = tmA;
    p.use_tma_store = 0;
    if (p.store_out) {
      OCC_REQUIRE((reinterpret_cast<uintptr_t>(out_b) & 15) == 0);
      // pooled mode keeps the cell maxima in the staging buffers, so the output rows are written directly
      p.use_tma_store = 0;
    }
    rc = dispatch_gemm(tmA, tmC, membed + (size_t)b * Q * E, p, p.tiles_x * p.tiles_y * p.tiles_z, stream);
    if (rc) return rc;
  }
  return OCC_OK;
}
