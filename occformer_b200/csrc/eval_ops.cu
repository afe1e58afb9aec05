// Evaluation tail on the device: the integer counts that feed the path's single collective.
//
//   SSCMetrics.get_score_completion / get_score_semantic_and_completion   projects/mmdet3d_plugin/utils/ssc_metric.py:104-168
//   OccupancyFormer.simple_evaluation_semantic + fast_hist_crop           occformer/detectors/occupancyformer.py:219-224,246-254
//                                                                         utils/metric_util.py:8-23
//   summed over ranks by one all-reduce / gather                          occformer/apis/test.py:195-212
//
// Both are confusion matrices: conf[target][pred] over the voxels (target != ignore) resp. hist[gt-1][pred-1] over the
// LiDAR points (gt in 1..K-1, pred = 1 + argmax of the class scores without the 'empty' column).  One pass over the
// labels, per-CTA histogram in shared memory, 64-bit global atomics once per CTA and bin -- exact integers in any order.
#include "occ_common.cuh"

namespace occ {

constexpr int EV_MAXK = 32;

__global__ void __launch_bounds__(256)
confusion_kernel(const unsigned char* __restrict__ pred, const unsigned char* __restrict__ target, long long n, int K,
                 int ignore, unsigned long long* __restrict__ conf) {
  __shared__ unsigned int sh[EV_MAXK * EV_MAXK];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x * 16;
  for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16; base < n; base += stride) {
    if (base + 16 <= n && ((reinterpret_cast<uintptr_t>(pred + base) | reinterpret_cast<uintptr_t>(target + base)) & 15) == 0) {
      const uint4 p4 = __ldg(reinterpret_cast<const uint4*>(pred + base));
      const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(target + base));
      const unsigned int pw[4] = {p4.x, p4.y, p4.z, p4.w}, tw[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int t = (tw[w] >> (8 * k)) & 255, p = (pw[w] >> (8 * k)) & 255;
          if (t != ignore && t < K && p < K) atomicAdd(&sh[t * K + p], 1u);
        }
    } else {
      for (long long i = base; i < n && i < base + 16; ++i) {
        const int t = target[i], p = pred[i];
        if (t != ignore && t < K && p < K) atomicAdd(&sh[t * K + p], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K; i += blockDim.x)
    if (sh[i]) atomicAdd(&conf[i], (unsigned long long)sh[i]);
}

// conf (K,K) [target][pred] -> out[0..2] = completion tp, fp, fn (occupied = class > 0); out[3 + c], out[3 + K + c],
// out[3 + 2K + c] = semantic tp, fp, fn of class c  (ssc_metric.py:104-168: masked by target != 255, no nonempty mask)
__global__ void ssc_pack_kernel(const unsigned long long* __restrict__ conf, int K, long long* __restrict__ out) {
  const int c = threadIdx.x;
  if (c < K) {
    unsigned long long row = 0, col = 0;
    for (int j = 0; j < K; ++j) { row += conf[c * K + j]; col += conf[j * K + c]; }
    const unsigned long long tp = conf[c * K + c];
    out[3 + c] = (long long)tp;
    out[3 + K + c] = (long long)(col - tp);
    out[3 + 2 * K + c] = (long long)(row - tp);
  }
  if (c == 0) {
    unsigned long long ctp = 0, cfp = 0, cfn = 0;
    for (int t = 0; t < K; ++t)
      for (int p = 0; p < K; ++p) {
        const unsigned long long v = conf[t * K + p];
        if (t > 0 && p > 0) ctp += v;
        else if (t == 0 && p > 0) cfp += v;
        else if (t > 0 && p == 0) cfn += v;
      }
    out[0] = (long long)ctp; out[1] = (long long)cfp; out[2] = (long long)cfn;
  }
}

// scores (n, K) fp32 point class scores (column 0 = 'empty'), labels (n) int64 in 0..K-1 (0 = unlabelled, skipped):
// hist[(gt-1)*(K-1) + (pred-1)] += 1 with pred = 1 + argmax(scores[:, 1:]) (first maximum, as torch.argmax)
__global__ void __launch_bounds__(256)
lidarseg_hist_kernel(const float* __restrict__ scores, const long long* __restrict__ labels, int n, int K,
                     unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[EV_MAXK * EV_MAXK];
  const int M = K - 1;
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long long g = labels[i];
    if (g < 1 || g >= K) continue;
    const float* s = scores + (size_t)i * K;
    int best = 1;
    float bv = s[1];
    for (int c = 2; c < K; ++c)
      if (s[c] > bv) { bv = s[c]; best = c; }
    atomicAdd(&sh[((int)g - 1) * M + best - 1], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M * M; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

}  // namespace occ

using namespace occ;

// pred, target: n uint8 labels; conf_ws: K*K int64 scratch (overwritten); out: 3 + 3K int64 (overwritten)
extern "C" int occ_ssc_counts(const unsigned char* pred, const unsigned char* target, long long n, int K, int ignore,
                              long long* conf_ws, long long* out, cudaStream_t stream) {
  OCC_REQUIRE(pred && target && conf_ws && out && n >= 0 && K >= 2 && K <= EV_MAXK);
  OCC_CUDA(cudaMemsetAsync(conf_ws, 0, (size_t)K * K * sizeof(long long), stream));
  if (n > 0) {
    long long blocks = (n + 256 * 16 - 1) / (256 * 16);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    confusion_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pred, target, n, K, ignore,
                                                           reinterpret_cast<unsigned long long*>(conf_ws));
    OCC_LAUNCH_CHECK();
  }
  ssc_pack_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const unsigned long long*>(conf_ws), K, out);
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}

// hist ((K-1)*(K-1) int64) is ACCUMULATED (+=): zero it once per evaluation run
extern "C" int occ_lidarseg_hist(const float* scores, const long long* labels, int n, int K, long long* hist,
                                 cudaStream_t stream) {
  OCC_REQUIRE(hist && n >= 0 && K >= 2 && K <= EV_MAXK);
  if (n == 0) return OCC_OK;
  OCC_REQUIRE(scores && labels);
  int blocks = (n + 255) / 256;
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  lidarseg_hist_kernel<<<blocks, 256, 0, stream>>>(scores, labels, n, K, reinterpret_cast<unsigned long long*>(hist));
  OCC_LAUNCH_CHECK();
  return OCC_OK;
}
