// LSS lift-splat voxel pooling for sm_100a (HBM-bound; output-stationary, every grid row written once).
//
// Reference path replaced (files under /root/reference):
//   projects/mmdet3d_plugin/occformer/image2bev/ViewTransformerLSSVoxel.py:77-100 (voxel_pooling: index,
//     kept mask, boolean-mask gathers), :110-115 (lift: depth softmax (x) context, 242 MB volume),
//   mmdetection3d/mmdet3d/ops/bev_pool/bev_pool.py:83-97 (rank, argsort, 3 gathers, interval bookkeeping),
//   mmdetection3d/mmdet3d/ops/bev_pool/src/bev_pool_cuda.cu:20-42 (interval-sum kernel) + the
//     (B,Z,X,Y,C)->(B,C,Z,X,Y) transpose copy (bev_pool.py:96).
//
// Pipeline (all on the caller's stream, no allocation, no sync):
//   index   : one thread per frustum point: exact fp32 voxel index ((g - (bx - dx/2)) / dx, trunc toward 0),
//             kept test against float nx, linear voxel id, per-voxel count (integer atomics).
//   scan    : exclusive prefix sum of the counts -> interval starts (3 small kernels).
//   fill    : counting-sort placement of kept point ids into per-voxel segments.
//   pool    : one warp per voxel: sums depth[p] * feat[pixel(p), :] (fused lift: the volume is never
//             materialised) or rows of a materialised feats[n, C] (drop-in bev_pool), and writes the
//             C-float row of out[b, x, y, z, :] exactly once -- empty voxels included, so no zero-fill pass.
// Output layout is channel-last (B, X, Y, Z, C); the Python boundary returns permuted views with the
// reference's shapes.
#include "occ_common.cuh"

namespace occ {

// ------------------------------------------------------------------------------------------------ index
// geom: (P,3) fp32 ego coordinates of frustum point p = ((b*N+n)*D+d)*fH*fW + pix.  vox_id[p] = linear voxel
// id ((b*X+x)*Y+y)*Z+z or -1 when dropped.
__global__ void vp_index_geom_kernel(const float* __restrict__ geom, int P, int points_per_batch, float dx0,
                                     float dx1, float dx2, float bx0, float bx1, float bx2, float nx0, float nx1,
                                     float nx2, int X, int Y, int Z, int* __restrict__ vox_id,
                                     int* __restrict__ counts) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float gx = geom[3 * (size_t)p + 0], gy = geom[3 * (size_t)p + 1], gz = geom[3 * (size_t)p + 2];
  // ViewTransformerLSSVoxel.py:84 -- fp32 sub, fp32 div (IEEE, no contraction), .long() truncation
  const float ox = __fsub_rn(bx0, __fdiv_rn(dx0, 2.0f));
  const float oy = __fsub_rn(bx1, __fdiv_rn(dx1, 2.0f));
  const float oz = __fsub_rn(bx2, __fdiv_rn(dx2, 2.0f));
  const long long ix = (long long)__fdiv_rn(__fsub_rn(gx, ox), dx0);
  const long long iy = (long long)__fdiv_rn(__fsub_rn(gy, oy), dx1);
  const long long iz = (long long)__fdiv_rn(__fsub_rn(gz, oz), dx2);
  // :90-92 -- int64 index compared with the *float* nx, upper bound exclusive
  const bool kept = ix >= 0 && (float)ix < nx0 && iy >= 0 && (float)iy < nx1 && iz >= 0 && (float)iz < nx2 &&
                    ix < X && iy < Y && iz < Z;
  int v = -1;
  if (kept) {
    const int b = p / points_per_batch;
    v = ((b * X + (int)ix) * Y + (int)iy) * Z + (int)iz;
    atomicAdd(&counts[v], 1);
  }
  vox_id[p] = v;
}

// coords: (n,4) int64 (x,y,z,b) as handed to mmdet3d.ops.bev_pool.bev_pool (already range-filtered by the
// caller in the reference; we re-check and drop out-of-range rows instead of writing out of bounds).
__global__ void vp_index_coords_kernel(const long long* __restrict__ coords, int n, int B, int X, int Y, int Z,
                                       int* __restrict__ vox_id, int* __restrict__ counts) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const long long x = coords[4 * (size_t)p + 0], y = coords[4 * (size_t)p + 1], z = coords[4 * (size_t)p + 2],
                  b = coords[4 * (size_t)p + 3];
  int v = -1;
  if (x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z && b >= 0 && b < B) {
    v = (((int)b * X + (int)x) * Y + (int)y) * Z + (int)z;
    atomicAdd(&counts[v], 1);
  }
  vox_id[p] = v;
}

// ------------------------------------------------------------------------------------------------ scan
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  const int warp_off = w == 0 ? 0 : warp_sums[w - 1];
  *total = warp_sums[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return warp_off + inc - v;
}

__global__ void vp_scan_reduce_kernel(const int* __restrict__ counts, int V, int* __restrict__ block_sums) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) s += (base + i < V) ? counts[base + i] : 0;
  int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void vp_scan_blocksums_kernel(int* __restrict__ block_sums, int nb, int* __restrict__ grand_total) {
  int carry = 0;
  for (int base = 0; base < nb; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? block_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, &total);
    if (i < nb) block_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void vp_scan_apply_kernel(const int* __restrict__ counts, int V, const int* __restrict__ block_sums,
                                     int* __restrict__ starts) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int c[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    c[i] = (base + i < V) ? counts[base + i] : 0;
    s += c[i];
  }
  int total;
  int off = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < V) starts[base + i] = off;
    off += c[i];
  }
  // starts[V] (= n_kept) is written by vp_scan_blocksums_kernel through grand_total
}

// ------------------------------------------------------------------------------------------------ fill
// Counting-sort placement; `counts` is consumed (decremented back to zero), so the workspace is clean
// for the next call.  Order inside a voxel follows atomic arrival (the reference's argsort is unstable
// too: bev_pool.py:92), the segment *sets* are deterministic.
__global__ void vp_fill_kernel(const int* __restrict__ vox_id, int P, const int* __restrict__ starts,
                               int* __restrict__ counts, int* __restrict__ order) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int v = vox_id[p];
  if (v < 0) return;
  const int slot = starts[v] + atomicSub(&counts[v], 1) - 1;
  order[slot] = p;
}

// ------------------------------------------------------------------------------------------------ pool
// MODE 0: fused lift  row(p) = depth_prob[p] * feat_cl[pixel(p), :]
// MODE 1: materialised rows feats[p, :]
template <int MODE>
__global__ void __launch_bounds__(256)
vp_pool_kernel(const int* __restrict__ starts, const int* __restrict__ order, int V, int C,
               const float* __restrict__ depth_prob, const float* __restrict__ feat, int D, int HW,
               float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int C4 = C >> 2;
  const int nbatch = (V + 31) >> 5;
  // a warp owns 32 consecutive voxels per iteration: their interval bounds arrive with two coalesced loads (instead of
  // two dependent scalar loads per voxel) and the warp streams 32 x C floats of contiguous output
  for (int batch = blockIdx.x * warps_per_block + (threadIdx.x >> 5); batch < nbatch; batch += gridDim.x * warps_per_block) {
    const int vb = batch << 5;
    const int my_s0 = starts[min(vb + lane, V)], my_s1 = starts[min(vb + lane + 1, V)];
    const int nv = min(32, V - vb);
    // empty voxels (77 % of a nuScenes grid) first, as a plain zero stream; then the occupied ones
    uint32_t occupied = __ballot_sync(0xffffffffu, lane < nv && my_s1 > my_s0);
    {
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int cbase = 0; cbase < C4; cbase += 32) {
        if (cbase + lane < C4) {
#pragma unroll 8
          for (int vi = 0; vi < nv; ++vi)
            if (!((occupied >> vi) & 1u)) __stcs(reinterpret_cast<float4*>(out + (size_t)(vb + vi) * C) + cbase + lane, z4);
        }
      }
    }
    while (occupied) {
    const int vi = __ffs(occupied) - 1;
    occupied &= occupied - 1;
    const int v = vb + vi;
    const int s0 = __shfl_sync(0xffffffffu, my_s0, vi), s1 = __shfl_sync(0xffffffffu, my_s1, vi);
    float4* orow = reinterpret_cast<float4*>(out + (size_t)v * C);
    for (int cbase = 0; cbase < C4; cbase += 32) {
      const int c4 = cbase + lane;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = s0; s < s1; s += 32) {
        const int cnt = min(32, s1 - s);
        int pid = 0;
        float w = 0.f;
        size_t rowoff = 0;
        if (lane < cnt) {
          pid = order[s + lane];
          if (MODE == 0) {
            w = __ldg(depth_prob + pid);
            // p = ((bn*D + d)*HW + pix)  ->  feature row = bn*HW + pix
            const int bn = pid / (D * HW);
            const int pix = pid % HW;
            rowoff = ((size_t)bn * HW + pix) * C;
          } else {
            w = 1.f;
            rowoff = (size_t)pid * C;
          }
        }
        const bool active = c4 < C4;
        int j = 0;
        for (; j + 4 <= cnt; j += 4) {
          float4 f[4];
          float ww[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const size_t ro = __shfl_sync(0xffffffffu, rowoff, j + u);
            ww[u] = __shfl_sync(0xffffffffu, w, j + u);
            f[u] = active ? __ldg(reinterpret_cast<const float4*>(feat + ro) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc.x += ww[u] * f[u].x; acc.y += ww[u] * f[u].y; acc.z += ww[u] * f[u].z; acc.w += ww[u] * f[u].w;
          }
        }
        for (; j < cnt; ++j) {
          const size_t ro = __shfl_sync(0xffffffffu, rowoff, j);
          const float wj = __shfl_sync(0xffffffffu, w, j);
          const float4 f = active ? __ldg(reinterpret_cast<const float4*>(feat + ro) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          acc.x += wj * f.x; acc.y += wj * f.y; acc.z += wj * f.z; acc.w += wj * f.w;
        }
      }
      if (c4 < C4) __stcs(orow + c4, acc);  // streaming store: the grid is consumed by the next kernel from HBM/L2
    }
    }
  }
}

// ------------------------------------------------------------------------------------------------ lift prologue
// depth softmax over D for every (camera, pixel): logits (BN, D, HW) -> prob (BN, D, HW)
__global__ void vp_depth_softmax_kernel(const float* __restrict__ logits, int BN, int D, int HW,
                                        float* __restrict__ prob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BN * HW) return;
  const int bn = i / HW, pix = i % HW;
  const float* src = logits + (size_t)bn * D * HW + pix;
  float* dst = prob + (size_t)bn * D * HW + pix;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, src[(size_t)d * HW]);
  float s = 0.f;
  for (int d = 0; d < D; ++d) s += expf(src[(size_t)d * HW] - m);
  for (int d = 0; d < D; ++d) dst[(size_t)d * HW] = expf(src[(size_t)d * HW] - m) / s;
}

// (BN, C, HW) -> (BN, HW, C)
__global__ void vp_nchw_to_nhwc_kernel(const float* __restrict__ in, int C, int HW, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int bn = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? in[((size_t)bn * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) out[((size_t)bn * HW + p) * C + c] = tile[threadIdx.x][i];
  }
}

static int scan_and_fill(int* counts, int* starts, int* block_sums, int* order, const int* vox_id, int P, int V,
                         cudaStream_t stream) {
  const int nb = (V + SCAN_TILE - 1) / SCAN_TILE;
  vp_scan_reduce_kernel<<<nb, SCAN_THREADS, 0, stream>>>(counts, V, block_sums);
  OCC_LAUNCH_CHECK();
  vp_scan_blocksums_kernel<<<1, SCAN_THREADS, 0, stream>>>(block_sums, nb, starts + V);
  OCC_LAUNCH_CHECK();
  vp_scan_apply_kernel<<<nb, SCAN_THREADS, 0, stream>>>(counts, V, block_sums, starts);
  OCC_LAUNCH_CHECK();
  if (P > 0) {
    vp_fill_kernel<<<(P + 255) / 256, 256, 0, stream>>>(vox_id, P, starts, counts, order);
    OCC_LAUNCH_CHECK();
  }
  return OCC_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct VpWorkspace {
  int *counts, *starts, *block_sums, *order, *vox_id;
};

static size_t vp_layout(void* base, int P, int V, VpWorkspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align_up(bytes, 256);
    return static_cast<int*>(p);
  };
  const int nb = (V + SCAN_TILE - 1) / SCAN_TILE;
  ws->counts = take((size_t)V * 4);
  ws->starts = take((size_t)(V + 1) * 4);
  ws->block_sums = take((size_t)nb * 4);
  ws->order = take((size_t)P * 4);
  ws->vox_id = take((size_t)P * 4);
  return off;
}

}  // namespace occ

using namespace occ;

extern "C" size_t occ_voxel_pool_workspace_bytes(int n_points, int B, int X, int Y, int Z) {
  VpWorkspace ws;
  return vp_layout(nullptr, n_points, B * X * Y * Z, &ws);
}

// Fused lift-splat.  depth_prob (B*N, D, fH*fW) fp32 (already softmaxed), feat_cl (B*N, fH*fW, C) channel-last,
// geom (B*N*D*fH*fW, 3).  out (B, X, Y, Z, C).  Bookkeeping left in the workspace: vox_id[P], starts[V+1].
extern "C" int occ_lift_splat(const float* depth_prob, const float* feat_cl, const float* geom, float* out, int B,
                              int N, int D, int HW, int C, float dx0, float dx1, float dx2, float bx0, float bx1,
                              float bx2, float nx0, float nx1, float nx2, int X, int Y, int Z, void* workspace, size_t workspace_bytes, int counts_are_zero,
                              cudaStream_t stream) {
  OCC_REQUIRE(depth_prob && feat_cl && geom && out && workspace);
  OCC_REQUIRE(B > 0 && N > 0 && D > 0 && HW > 0 && C > 0 && C % 4 == 0 && X > 0 && Y > 0 && Z > 0);
  const long long Pll = (long long)B * N * D * HW, Vll = (long long)B * X * Y * Z;
  OCC_REQUIRE(Pll < (1ll << 31) && Vll < (1ll << 31));
  const int P = (int)Pll, V = (int)Vll;
  VpWorkspace ws;
  OCC_REQUIRE(vp_layout(workspace, P, V, &ws) <= workspace_bytes);
  if (!counts_are_zero) OCC_CUDA(cudaMemsetAsync(ws.counts, 0, (size_t)V * 4, stream));
  vp_index_geom_kernel<<<(P + 255) / 256, 256, 0, stream>>>(geom, P, N * D * HW, dx0, dx1, dx2, bx0, bx1, bx2, nx0,
                                                             nx1, nx2, X, Y, Z, ws.vox_id, ws.counts);
  OCC_LAUNCH_CHECK();
  int rc = scan_and_fill(ws.counts, ws.starts, ws.block_sums, ws.order, ws.vox_id, P, V, stream);
  if (rc) return rc;
  const int blocks = sm_count() * 8;
  vp_pool_kernel<0><<<blocks, 256, 0, stream>>>(ws.starts, ws.order, V, C, depth_prob, feat_cl, D, HW, out);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}

// Drop-in for mmdet3d.ops.bev_pool.bev_pool: feats (n, C), coords (n, 4) int64 (x, y, z, b).
// out (B, X, Y, Z, C)  [reference returns the same values as (B, C, Z, X, Y)].
extern "C" int occ_bev_pool(const float* feats, const long long* coords, float* out, int n, int C, int B, int X,
                            int Y, int Z, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  OCC_REQUIRE(out && workspace);
  OCC_REQUIRE(n >= 0 && C > 0 && C % 4 == 0 && B > 0 && X > 0 && Y > 0 && Z > 0);
  OCC_REQUIRE(n == 0 || (feats && coords));
  const long long Vll = (long long)B * X * Y * Z;
  OCC_REQUIRE(Vll < (1ll << 31));
  const int V = (int)Vll;
  VpWorkspace ws;
  OCC_REQUIRE(vp_layout(workspace, n, V, &ws) <= workspace_bytes);
  OCC_CUDA(cudaMemsetAsync(ws.counts, 0, (size_t)V * 4, stream));
  if (n > 0) {
    vp_index_coords_kernel<<<(n + 255) / 256, 256, 0, stream>>>(coords, n, B, X, Y, Z, ws.vox_id, ws.counts);
    OCC_LAUNCH_CHECK();
  }
  int rc = scan_and_fill(ws.counts, ws.starts, ws.block_sums, ws.order, ws.vox_id, n, V, stream);
  if (rc) return rc;
  vp_pool_kernel<1><<<sm_count() * 8, 256, 0, stream>>>(ws.starts, ws.order, V, C, nullptr, feats, 1, 1, out);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}

// Lift prologue: depth softmax + NCHW->NHWC of the context features (ViewTransformerLSSVoxel.py:108-110).
extern "C" int occ_lift_prologue(const float* depth_logits, const float* img_feat, float* depth_prob,
                                 float* feat_cl, int BN, int D, int C, int HW, cudaStream_t stream) {
  OCC_REQUIRE(depth_logits && img_feat && depth_prob && feat_cl && BN > 0 && D > 0 && C > 0 && HW > 0);
  vp_depth_softmax_kernel<<<(BN * HW + 127) / 128, 128, 0, stream>>>(depth_logits, BN, D, HW, depth_prob);
  OCC_LAUNCH_CHECK();
  dim3 grid((HW + 31) / 32, (C + 31) / 32, BN), block(32, 8);
  vp_nchw_to_nhwc_kernel<<<grid, block, 0, stream>>>(img_feat, C, HW, feat_cl);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}

// Byte offsets of the bookkeeping arrays inside the workspace (for tests / callers that want them).
extern "C" int occ_voxel_pool_workspace_layout(int n_points, int B, int X, int Y, int Z, size_t* off_counts,
                                               size_t* off_starts, size_t* off_order, size_t* off_vox_id) {
  VpWorkspace ws;
  char* base = reinterpret_cast<char*>(0x1000);
  vp_layout(base, n_points, B * X * Y * Z, &ws);
  *off_counts = reinterpret_cast<char*>(ws.counts) - base;
  *off_starts = reinterpret_cast<char*>(ws.starts) - base;
  *off_order = reinterpret_cast<char*>(ws.order) - base;
  *off_vox_id = reinterpret_cast<char*>(ws.vox_id) - base;
  return OCC_OK;
}

// ViewTransformerLiftSplatShootVoxel.voxel_pooling(geom_feats, x) with a MATERIALISED volume
// (ViewTransformerLSSVoxel.py:77-100): feats (P, C) rows, geom (P, 3).  out (B, X, Y, Z, C).
extern "C" int occ_voxel_pool_geom(const float* feats, const float* geom, float* out, int B, int points_per_batch,
                                   int C, float dx0, float dx1, float dx2, float bx0, float bx1, float bx2, float nx0,
                                   float nx1, float nx2, int X, int Y, int Z, void* workspace, size_t workspace_bytes,
                                   cudaStream_t stream) {
  OCC_REQUIRE(feats && geom && out && workspace);
  OCC_REQUIRE(B > 0 && points_per_batch > 0 && C > 0 && C % 4 == 0 && X > 0 && Y > 0 && Z > 0);
  const long long Pll = (long long)B * points_per_batch, Vll = (long long)B * X * Y * Z;
  OCC_REQUIRE(Pll < (1ll << 31) && Vll < (1ll << 31));
  const int P = (int)Pll, V = (int)Vll;
  VpWorkspace ws;
  OCC_REQUIRE(vp_layout(workspace, P, V, &ws) <= workspace_bytes);
  OCC_CUDA(cudaMemsetAsync(ws.counts, 0, (size_t)V * 4, stream));
  vp_index_geom_kernel<<<(P + 255) / 256, 256, 0, stream>>>(geom, P, points_per_batch, dx0, dx1, dx2, bx0, bx1, bx2,
                                                             nx0, nx1, nx2, X, Y, Z, ws.vox_id, ws.counts);
  OCC_LAUNCH_CHECK();
  int rc = scan_and_fill(ws.counts, ws.starts, ws.block_sums, ws.order, ws.vox_id, P, V, stream);
  if (rc) return rc;
  vp_pool_kernel<1><<<sm_count() * 8, 256, 0, stream>>>(ws.starts, ws.order, V, C, nullptr, feats, 1, 1, out);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}

// ------------------------------------------------------------------------------------------------ frustum geometry
// ViewTransformerLSSBEVDepth.get_geometry (projects/mmdet3d_plugin/occformer/image2bev/ViewTransformerLSSBEVDepth.py:117-150)
// as one kernel: undo the image augmentation (post_trans / post_rots), unproject with depth, camera -> ego
// (rots * inv(intrins), + trans), then the BEV augmentation matrix bda.  The reference issues ~25 batched 3x3
// cuBLAS gemv launches over 473k points for this; the arithmetic is 3 small matvecs per point.
// One CTA column (blockIdx.y) per camera (b, n); thread = frustum point.
namespace occ {

__device__ __forceinline__ void inv3x3(const float* m, float* o) {
  // adjugate / determinant in double, rounded once to fp32
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  const double r = 1.0 / det;
  o[0] = (float)(A * r); o[1] = (float)(-(b * i - c * h) * r); o[2] = (float)((b * f - c * e) * r);
  o[3] = (float)(B * r); o[4] = (float)((a * i - c * g) * r);  o[5] = (float)(-(a * f - c * d) * r);
  o[6] = (float)(C * r); o[7] = (float)(-(a * h - b * g) * r); o[8] = (float)((a * e - b * d) * r);
}

__device__ __forceinline__ void mv3(const float* m, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = fmaf(m[2], z, fmaf(m[1], y, m[0] * x));
  oy = fmaf(m[5], z, fmaf(m[4], y, m[3] * x));
  oz = fmaf(m[8], z, fmaf(m[7], y, m[6] * x));
}

__global__ void __launch_bounds__(256)
lss_geometry_kernel(const float* __restrict__ frustum /*(P,3)*/, int P, const float* __restrict__ rots,
                    const float* __restrict__ trans, const float* __restrict__ intrins, int intrin_cols,
                    const float* __restrict__ post_rots, const float* __restrict__ post_trans,
                    const float* __restrict__ bda, int bda_dim, int N, float* __restrict__ geom) {
  __shared__ float s_ipr[9], s_comb[9], s_vec[9], s_bda[16];
  const int bn = blockIdx.y, b = bn / N;
  if (threadIdx.x == 0) {
    inv3x3(post_rots + (size_t)bn * 9, s_ipr);
    float K[9], iK[9];
    const float* I = intrins + (size_t)bn * 3 * intrin_cols;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) K[r * 3 + c] = I[r * intrin_cols + c];
    inv3x3(K, iK);
    const float* R = rots + (size_t)bn * 9;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        s_comb[r * 3 + c] = fmaf(R[r * 3 + 2], iK[6 + c], fmaf(R[r * 3 + 1], iK[3 + c], R[r * 3] * iK[c]));
    for (int k = 0; k < 3; ++k) {
      s_vec[k] = post_trans[(size_t)bn * 3 + k];
      s_vec[3 + k] = trans[(size_t)bn * 3 + k];
      s_vec[6 + k] = intrin_cols == 4 ? I[k * 4 + 3] : 0.f;  // KITTI P2 shift (:134-137)
    }
    for (int k = 0; k < bda_dim * bda_dim; ++k) s_bda[k] = bda[(size_t)b * bda_dim * bda_dim + k];
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float x = frustum[3 * (size_t)p] - s_vec[0], y = frustum[3 * (size_t)p + 1] - s_vec[1],
        z = frustum[3 * (size_t)p + 2] - s_vec[2];
  float u, v, w;
  mv3(s_ipr, x, y, z, u, v, w);
  u = u * w - s_vec[6]; v = v * w - s_vec[7]; w = w - s_vec[8];
  mv3(s_comb, u, v, w, x, y, z);
  x += s_vec[3]; y += s_vec[4]; z += s_vec[5];
  float ox, oy, oz;
  if (bda_dim == 4) {
    ox = fmaf(s_bda[2], z, fmaf(s_bda[1], y, s_bda[0] * x)) + s_bda[3];
    oy = fmaf(s_bda[6], z, fmaf(s_bda[5], y, s_bda[4] * x)) + s_bda[7];
    oz = fmaf(s_bda[10], z, fmaf(s_bda[9], y, s_bda[8] * x)) + s_bda[11];
  } else {
    mv3(s_bda, x, y, z, ox, oy, oz);
  }
  float* o = geom + ((size_t)bn * P + p) * 3;
  o[0] = ox; o[1] = oy; o[2] = oz;
}

}  // namespace occ

extern "C" int occ_lss_geometry(const float* frustum, int P, const float* rots, const float* trans, const float* intrins,
                                int intrin_cols, const float* post_rots, const float* post_trans, const float* bda,
                                int bda_dim, int B, int N, float* geom, cudaStream_t stream) {
  OCC_REQUIRE(frustum && rots && trans && intrins && post_rots && post_trans && bda && geom);
  OCC_REQUIRE(P > 0 && B > 0 && N > 0 && (intrin_cols == 3 || intrin_cols == 4) && (bda_dim == 3 || bda_dim == 4));
  OCC_REQUIRE((long long)B * N <= 65535);
  dim3 grid((P + 255) / 256, B * N);
  occ::lss_geometry_kernel<<<grid, 256, 0, stream>>>(frustum, P, rots, trans, intrins, intrin_cols, post_rots, post_trans,
                                                     bda, bda_dim, N, geom);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}
