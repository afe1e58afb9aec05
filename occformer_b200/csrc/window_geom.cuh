// Shared geometry of the two window-attention kernels (window_attn.cu: q/k/v from HBM; swin_attn_fused.cu: QKV
// projection inside the kernel): 7x7 windows over the B*(Z+1) X-Y images, token -> global row, shift-mask regions.
#pragma once
#include "occ_common.cuh"
#include "occ_ptx.cuh"

namespace occ {

constexpr int WS = 7;
constexpr int WT = WS * WS;  // 49 tokens
constexpr int HD = 32;       // head dim (multihead_base_channel, dualpath_block.py:32)
constexpr int WA_STAGES = 4;
constexpr int WA_TILE = 128 * HD * 4;                  // 16 KB: 128 rows x 128 B
constexpr int WA_BIAS_FLOATS = 2404;                   // 49*49 padded to a 16-byte multiple
constexpr int WA_OFF_ROWS = 3 * WA_TILE;               // 49152 (8-byte aligned)
constexpr int WA_OFF_REGION = WA_OFF_ROWS + 128 * 8;   // 50176
constexpr int WA_OFF_SAME = WA_OFF_REGION + 128 * 4;   // 50688: per window half, 9 x uint64 "key j is in region r" masks
constexpr int WA_OFF_UNI = WA_OFF_SAME + 2 * 9 * 8;    // 50832: per window half, 1 if all 49 tokens share a region
constexpr int WA_OFF_SRC = WA_OFF_UNI + 8;             // 50840: per row, the global source pointer of its q slice (8 B)
constexpr int WA_STAGE_BYTES = 51 * 1024;              // 52224 >= 50840 + 1024, multiple of 1024
constexpr int WA_BIAS_LD = 52;                         // padded bias row pitch (floats): 13 conflict-free LDS.128 per row
constexpr int WA_BIAS_BYTES = 10240;                   // resident bias of this CTA's head, (49, 52) floats, * log2(e)
constexpr int WA_THREADS = 512;
constexpr uint32_t WA_TMEM_COLS = 512;                 // S/P: 2 x 128, O: 2 x 64

struct WinGeom {
  int B, X, Y, Z, C, heads, shift;
  int head_major;  // qkv columns ordered [head][q|k|v][32] instead of the reference's [q|k|v][head][32]
  int Xp, Yp, nWx, nWy;
  long long vox_rows;  // B*X*Y*Z
  long long nwin;
  int passes;  // tensor-core passes per 32-k block (occ_common.cuh: mma_passes)
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

// Inverse of window_token_row for real tokens: token (img, x, y) -> row of the WINDOW-LAYOUT buffer, window * 64 + t with
// window = (wx * nWy + wy) * n_images + img (the
// 49 tokens of a window are 49 consecutive rows, windows are 64 rows apart, rows 49..63 stay zero; a pair of windows is
// one 128-row MMA tile that a single TMA box fetches).  Window pad tokens (positions >= X or >= Y of the padded image)
// have no source token: their rows are never written and stay zero.
__device__ __forceinline__ long long window_layout_row(const WinGeom& g, int img, int x, int y) {
  int xs = x, ys = y;
  if (g.shift) {
    xs = x - 3; if (xs < 0) xs += g.Xp;
    ys = y - 3; if (ys < 0) ys += g.Yp;
  }
  const int wx = xs / WS, wy = ys / WS;
  const int t = (xs - wx * WS) * WS + (ys - wy * WS);
  // window index = (wx, wy) major, image minor: the Z + 1 images of a (b, x, y) column -- which the producer handles
  // together -- land 32 KB apart instead of one whole image (27 MB at 200 x 200) apart, and a pair of windows is the
  // same (wx, wy) in two adjacent height slices, whose output rows are neighbours in token order (measured neutral
  // against the image-major order: neither the producer's scatter nor the attention kernel is bound by that locality)
  return (((long long)wx * g.nWy + wy) * ((long long)g.B * (g.Z + 1)) + img) * 64 + t;
}

inline WinGeom make_win_geom(int B, int X, int Y, int Z, int C, int heads, int shift) {
  WinGeom g;
  g.B = B; g.X = X; g.Y = Y; g.Z = Z; g.C = C; g.heads = heads; g.shift = shift ? 1 : 0;
  g.head_major = 1;
  g.nWx = (X + WS - 1) / WS; g.nWy = (Y + WS - 1) / WS;
  g.Xp = g.nWx * WS; g.Yp = g.nWy * WS;
  g.vox_rows = (long long)B * X * Y * Z;
  g.nwin = (long long)B * (Z + 1) * g.nWx * g.nWy;
  g.passes = mma_passes();
  return g;
}

}  // namespace occ
