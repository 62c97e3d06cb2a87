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

// --------------------------------------------------------------------------------------
// acc / scene_acc of the sampling test (nusc_train.py:23-27, 336-343): mask_mean((score > 0), valid) over all chains
// and mask_mean((max over the S samples of a (scene, mode)) > 0, valid of the scene's lanes).  Rows n = (b*S + r)*3 + m.
// --------------------------------------------------------------------------------------
__global__ void k_accuracy_partial(const float* __restrict__ scores, const float* __restrict__ valid, int S,
                                   float* __restrict__ part /* (bs, 4): acc num, valid sum, scene num, lane-valid sum */) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ float s_num[3], s_val[3];
  __shared__ int s_pos[3];
  if (tid < 3) { s_num[tid] = 0.f; s_val[tid] = 0.f; s_pos[tid] = 0; }
  __syncthreads();
  for (int m = 0; m < 3; ++m) {
    float num = 0.f, val = 0.f;
    bool pos = false;  // max over the samples > 0  <=>  some sample > 0
    for (int r = tid; r < S; r += blockDim.x) {
      const size_t n = ((size_t)b * S + r) * 3 + m;
      const float sc = scores[n], v = valid[n];
      num += (sc > 0.f ? 1.f : 0.f) * v;
      val += v;
      pos = pos || sc > 0.f;
    }
    for (int o = 16; o > 0; o >>= 1) {
      num += __shfl_xor_sync(0xffffffffu, num, o);
      val += __shfl_xor_sync(0xffffffffu, val, o);
    }
    const bool any_pos = __any_sync(0xffffffffu, pos);
    if ((tid & 31) == 0) {  // counts are small integers (validity is 0/1): float atomics are exact and order-independent
      atomicAdd(&s_num[m], num);
      atomicAdd(&s_val[m], val);
      if (any_pos) atomicOr(&s_pos[m], 1);
    }
  }
  __syncthreads();
  if (tid == 0) {
    float scene_num = 0.f, lane_val = 0.f;
    for (int m = 0; m < 3; ++m) {
      const bool pos = s_pos[m] != 0;
      const float v0 = valid[((size_t)b * S) * 3 + m];
      scene_num += (pos ? 1.f : 0.f) * v0;
      lane_val += v0;
    }
    part[b * 4 + 0] = s_num[0] + s_num[1] + s_num[2];
    part[b * 4 + 1] = s_val[0] + s_val[1] + s_val[2];
    part[b * 4 + 2] = scene_num;
    part[b * 4 + 3] = lane_val;
  }
}

__global__ void k_accuracy_final(const float* __restrict__ part, int bs, int S, float* __restrict__ out) {
  __shared__ float sh[4][32];
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = threadIdx.x; b < bs; b += blockDim.x)
    for (int q = 0; q < 4; ++q) a[q] += part[b * 4 + q];
  for (int q = 0; q < 4; ++q) {
    for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
    if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = a[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < 4; ++q)
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t[q] += sh[q][w];
    const float n_all = (float)bs * (float)S * 3.f, n_lane = (float)bs * 3.f;
    out[0] = (t[0] / n_all) / fmaxf(t[1] / n_all, 1e-2f);    // mean(loss*mask) / clip(mean(mask), 1e-2)
    out[1] = (t[2] / n_lane) / fmaxf(t[3] / n_lane, 1e-2f);
  }
}

extern "C" int pstl_accuracy(const float* scores, const float* valid, int n_scenes, int S, float* partial, float* out,
                             pstl_stream_t stream) {
  PSTL_CHECK_ARG(scores && valid && partial && out && S >= 1, "bad argument");
  if (n_scenes <= 0) return PSTL_OK;
  k_accuracy_partial<<<n_scenes, 64, 0, (cudaStream_t)stream>>>(scores, valid, S, partial);
  PSTL_LAUNCH_CHECK();
  k_accuracy_final<<<1, 256, 0, (cudaStream_t)stream>>>(partial, n_scenes, S, out);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

