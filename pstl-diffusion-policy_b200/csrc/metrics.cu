// metrics.cu — diversity metrics of the sampling test on the device (reference nusc_api.py:817-877,
// measure_diversity): per (scene, lane mode) the masked population std of the accepted samples' way-points,
// averaged over the 2*nt features, and the summed convex-hull area of the accepted samples' positions at every step
// (upstream: numpy masked arrays + one scipy/Qhull ConvexHull call per (scene, mode, step) on the host).
#include "common.cuh"

#define PSTL_DIV_MAXM 128

// area of the convex hull of n points (Andrew's monotone chain on a sorted copy; shoelace).  Fewer than three
// points or collinear points give 0, where Qhull raises and upstream's except-branch counts 0.
__device__ double hull_area(const double* xs, const double* ys, int n, int* idx, int* hull) {
  if (n < 3) return 0.0;
  for (int i = 0; i < n; ++i) idx[i] = i;
  for (int i = 1; i < n; ++i) {  // insertion sort by (x, y)
    const int k = idx[i];
    int j = i - 1;
    while (j >= 0 && (xs[idx[j]] > xs[k] || (xs[idx[j]] == xs[k] && ys[idx[j]] > ys[k]))) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = k;
  }
  auto cross = [&](int o, int a, int b) {
    return (xs[a] - xs[o]) * (ys[b] - ys[o]) - (ys[a] - ys[o]) * (xs[b] - xs[o]);
  };
  int h = 0;
  for (int i = 0; i < n; ++i) {  // lower hull
    while (h >= 2 && cross(hull[h - 2], hull[h - 1], idx[i]) <= 0.0) --h;
    hull[h++] = idx[i];
  }
  const int lower = h + 1;
  for (int i = n - 2; i >= 0; --i) {  // upper hull
    while (h >= lower && cross(hull[h - 2], hull[h - 1], idx[i]) <= 0.0) --h;
    hull[h++] = idx[i];
  }
  --h;  // last point repeats the first
  if (h < 3) return 0.0;
  double a2 = 0.0;
  for (int i = 0; i < h; ++i) {
    const int p = hull[i], q = hull[(i + 1) % h];
    a2 += xs[p] * ys[q] - xs[q] * ys[p];
  }
  return 0.5 * fabs(a2);
}

// one block per (scene b, mode l); trajs (bs, m, 3, 2*nt), scores / valids (bs, m, 3)
__global__ void __launch_bounds__(64) k_diversity(const float* __restrict__ trajs, const float* __restrict__ scores,
                                                  const float* __restrict__ valids, int m, int nt,
                                                  float* __restrict__ std_out, float* __restrict__ vol_out) {
  __shared__ int acc_idx[PSTL_DIV_MAXM];
  __shared__ int n_acc_s;
  __shared__ double red[64];
  const int b = blockIdx.x / 3, l = blockIdx.x % 3, F = 2 * nt, tid = threadIdx.x;
  if (tid == 0) {
    int n = 0;
    for (int j = 0; j < m; ++j)
      if (scores[((size_t)b * m + j) * 3 + l] > 0.f) acc_idx[n++] = j;
    n_acc_s = n;
  }
  __syncthreads();
  const int n_acc = n_acc_s;
  auto at = [&](int j, int f) { return trajs[(((size_t)b * m + j) * 3 + l) * F + f]; };
  // masked population std per feature, mean over the features
  double part = 0.0;
  for (int f = tid; f < F; f += blockDim.x) {
    if (n_acc > 0) {
      double mean = 0.0;
      for (int q = 0; q < n_acc; ++q) mean += (double)at(acc_idx[q], f);
      mean /= n_acc;
      double var = 0.0;
      for (int q = 0; q < n_acc; ++q) { const double d = (double)at(acc_idx[q], f) - mean; var += d * d; }
      part += sqrt(var / n_acc);
    }
  }
  red[tid] = part;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < blockDim.x; ++i) s += red[i];
    std_out[b * 3 + l] = (float)(s / F);
  }
  __syncthreads();
  // summed hull area over the steps
  double vol = 0.0;
  const bool lane_valid = valids[((size_t)b * m) * 3 + l] == 1.f;
  if (lane_valid && n_acc >= 3) {
    double xs[PSTL_DIV_MAXM], ys[PSTL_DIV_MAXM];
    int idx[PSTL_DIV_MAXM], hull[2 * PSTL_DIV_MAXM];
    for (int t = tid; t < nt; t += blockDim.x) {
      for (int q = 0; q < n_acc; ++q) { xs[q] = at(acc_idx[q], 2 * t); ys[q] = at(acc_idx[q], 2 * t + 1); }
      vol += hull_area(xs, ys, n_acc, idx, hull);
    }
  }
  red[tid] = vol;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < blockDim.x; ++i) s += red[i];
    vol_out[b * 3 + l] = (float)s;
  }
}

extern "C" int pstl_diversity(const float* trajs, const float* scores, const float* valids, int n_scenes, int m, int nt,
                              float* std_out, float* vol_out, pstl_stream_t stream) {
  PSTL_CHECK_ARG(trajs && scores && valids && std_out && vol_out, "null argument");
  PSTL_CHECK_ARG(m >= 1 && m <= PSTL_DIV_MAXM && nt >= 1, "1 <= samples <= 128");
  if (n_scenes <= 0) return PSTL_OK;
  k_diversity<<<n_scenes * 3, 64, 0, (cudaStream_t)stream>>>(trajs, scores, valids, m, nt, std_out, vol_out);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}
