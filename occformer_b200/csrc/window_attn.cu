// Shifted-window multi-head self-attention core over the B*(Z+1) X-Y images of the dual-path encoder.
//
// Replaces ShiftWindowMSA.forward's pad / roll / mask build / window_partition / window_reverse / roll back /
// crop and WindowMSA.forward's score pipeline (q*scale, QK^T, + relative-position bias, + shift mask, softmax,
// PV) -- projects/mmdet3d_plugin/occformer/backbones/modules/window_attention.py:168-242 and :69-107 -- with
// index arithmetic: the kernel gathers the 49 tokens of a (possibly shifted, possibly padded) window straight
// from the token-ordered qkv tensor and scatters the result back to token order.  No (nW,49,49) mask tensor,
// no score tensor in HBM.
//
// Semantics kept bit-for-bit in structure:
//   * padding to a multiple of 7 happens AFTER LayerNorm: pad tokens are zeros, so their q/k/v equal the qkv
//     bias and they take part in the softmax as real keys; their outputs are cropped.
//   * shifted blocks: roll(-3,-3); region ids from slices (0,-7),(-7,-3),(-3,None) on the padded extent;
//     additive mask -100.0 (not -inf) where ids differ.
//   * relative-position bias added after the q*scale product.
#include "occ_common.cuh"
#include "occ_ptx.cuh"

namespace occ {

constexpr int WS = 7;
constexpr int WT = WS * WS;  // 49 tokens
constexpr int HD = 32;       // head dim (multihead_base_channel, dualpath_block.py:32)
constexpr int KV_LD = 36;    // smem row pitch (floats)
constexpr int WARP_SMEM = 2 * WT * KV_LD + ((WT * WT + 3) / 4) * 4;  // floats per warp, 16-byte multiple

struct WinGeom {
  int B, X, Y, Z, C, heads, shift;
  int Xp, Yp, nWx, nWy;
  long long vox_rows;  // B*X*Y*Z
};

// token t of window (img, wx, wy) -> global token row (or -1 for a pad token) and shift-mask region id
__device__ __forceinline__ long long window_token_row(const WinGeom& g, int img, int wx, int wy, int t, int* region) {
  const int i = t / WS, j = t % WS;
  const int xs = wx * WS + i, ys = wy * WS + j;
  int x = xs, y = ys;
  if (g.shift) {
    x = xs + 3; if (x >= g.Xp) x -= g.Xp;
    y = ys + 3; if (y >= g.Yp) y -= g.Yp;
    const int rx = xs < g.Xp - WS ? 0 : (xs < g.Xp - 3 ? 1 : 2);
    const int ry = ys < g.Yp - WS ? 0 : (ys < g.Yp - 3 ? 1 : 2);
    *region = rx * 3 + ry;
  } else {
    *region = 0;
  }
  if (x >= g.X || y >= g.Y) return -1;
  if (img < g.B * g.Z) {
    const int b = img / g.Z, z = img % g.Z;
    return (((long long)b * g.X + x) * g.Y + y) * g.Z + z;
  }
  const int b = img - g.B * g.Z;
  return g.vox_rows + ((long long)b * g.X + x) * g.Y + y;
}

// SIMT version: one warp per (window, head); lane = query row (two passes for 49 rows).
__global__ void __launch_bounds__(128)
window_attn_simt_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                        const float* __restrict__ bias_dense /*(heads,49,49)*/, float* __restrict__ out, WinGeom g) {
  extern __shared__ __align__(16) float smem[];
  __shared__ long long s_row[WT];
  __shared__ int s_region[WT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y * 4 + warp;
  int w = blockIdx.x;
  const int wy = w % g.nWy; w /= g.nWy;
  const int wx = w % g.nWx; w /= g.nWx;
  const int img = w;
  if (threadIdx.x < WT) {
    int region;
    s_row[threadIdx.x] = window_token_row(g, img, wx, wy, threadIdx.x, &region);
    s_region[threadIdx.x] = region;
  }
  __syncthreads();
  if (head >= g.heads) return;
  float* sk = smem + (size_t)warp * WARP_SMEM;
  float* sv = sk + WT * KV_LD;
  float* sb = sv + WT * KV_LD;
  const int C = g.C;
  const int qoff = head * HD, koff = C + head * HD, voff = 2 * C + head * HD;
  for (int t = 0; t < WT; ++t) {
    const long long r = s_row[t];
    const float* src = r >= 0 ? qkv + r * 3 * C : qkv_bias;
    sk[t * KV_LD + lane] = src[koff + lane];
    sv[t * KV_LD + lane] = src[voff + lane];
  }
  for (int i = lane; i < WT * WT; i += 32) sb[i] = bias_dense[(size_t)head * WT * WT + i];
  __syncwarp();
  const float scale = 0.17677669529663687f;  // 32^-0.5
  for (int pass = 0; pass < 2; ++pass) {
    const int i = pass * 32 + lane;
    if (i >= WT) break;
    const long long r = s_row[i];
    const float* qsrc = (r >= 0 ? qkv + r * 3 * C : qkv_bias) + qoff;
    float q[HD];
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
      const float4 t = *reinterpret_cast<const float4*>(qsrc + d);
      q[d] = t.x * scale; q[d + 1] = t.y * scale; q[d + 2] = t.z * scale; q[d + 3] = t.w * scale;
    }
    float s[WT];
    const int reg_i = s_region[i];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < WT; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(sk + j * KV_LD + d);
        a += q[d] * kk.x + q[d + 1] * kk.y + q[d + 2] * kk.z + q[d + 3] * kk.w;
      }
      a += sb[i * WT + j];
      if (s_region[j] != reg_i) a += -100.0f;
      s[j] = a;
      m = fmaxf(m, a);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < WT; ++j) {
      s[j] = expf(s[j] - m);
      sum += s[j];
    }
    const float inv = 1.0f / sum;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < WT; ++j) {
      const float pj = s[j] * inv;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(sv + j * KV_LD + d);
        o[d] += pj * vv.x; o[d + 1] += pj * vv.y; o[d + 2] += pj * vv.z; o[d + 3] += pj * vv.w;
      }
    }
    if (r >= 0) {
      float* dst = out + r * C + head * HD;
#pragma unroll
      for (int d = 0; d < HD; d += 4)
        *reinterpret_cast<float4*>(dst + d) =
            make_float4(round_tf32(o[d]), round_tf32(o[d + 1]), round_tf32(o[d + 2]), round_tf32(o[d + 3]));
    }
  }
}

}  // namespace occ

using namespace occ;

// qkv (rows, 3C) with rows = B*X*Y*(Z+1) token-ordered (voxel tokens then BEV tokens); out (rows, C).
// bias_dense = relative_position_bias_table[relative_position_index] arranged (heads, 49, 49).
extern "C" int occ_window_attention(const float* qkv, const float* qkv_bias, const float* bias_dense, float* out, int B,
                                    int X, int Y, int Z, int C, int heads, int shift, cudaStream_t stream) {
  OCC_REQUIRE(qkv && qkv_bias && bias_dense && out);
  OCC_REQUIRE(B > 0 && X > 0 && Y > 0 && Z > 0 && heads > 0 && C == heads * HD);
  WinGeom g;
  g.B = B; g.X = X; g.Y = Y; g.Z = Z; g.C = C; g.heads = heads; g.shift = shift ? 1 : 0;
  g.nWx = (X + WS - 1) / WS; g.nWy = (Y + WS - 1) / WS;
  g.Xp = g.nWx * WS; g.Yp = g.nWy * WS;
  g.vox_rows = (long long)B * X * Y * Z;
  const long long nwin = (long long)B * (Z + 1) * g.nWx * g.nWy;
  OCC_REQUIRE(nwin < (1ll << 31));
  const size_t smem = 4 * WARP_SMEM * sizeof(float);
  static bool configured = false;
  if (!configured) {
    OCC_CUDA(cudaFuncSetAttribute(window_attn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((unsigned)nwin, (heads + 3) / 4);
  window_attn_simt_kernel<<<grid, 128, smem, stream>>>(qkv, qkv_bias, bias_dense, out, g);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}
