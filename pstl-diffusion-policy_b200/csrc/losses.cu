// losses.cu — the RefineNet training losses on the device, value and gradient in one call (reference
// nusc_train.py:411 loss_stl; :439-466 the --diverse_loss branch: determinantal-point-process diversity
// tr(I - (Q S Q + I)^-1) over groups of n_randoms/n_shards samples of one (scene, lane mode) plus the masked
// regulariser; :468-478 the plain branch: per-channel regulariser + soft box penalty).
//
// Upstream builds (groups, G, G, 40) differences, a batched torch.inverse (LU, one cuSOLVER call) and lets autograd
// replay it all backwards (~40 launches).  Here one warp owns one group: lane i holds row i of the G x G matrices in
// shared memory, the inverse is an in-place Gauss-Jordan sweep (the matrix is I + PSD, so no pivoting), and the
// gradient uses d tr(M^-1)/dM = -(M^-2)^T directly.  Global denominators (the mask_mean clips) come from a
// deterministic two-level reduction (per-group partials, one finalising block, fp64), then one elementwise kernel
// adds the regulariser / STL-hinge terms that depend on them.
#include "common.cuh"

namespace {

constexpr int kPart = 8;  // floats of partial sums per group / block

struct LossArgs {
  const float *rect, *nn, *scores, *valid;
  float *d_rect, *d_scores;
  float* part;
  float* fin;  // 16 floats: finalised scale factors for the gradient kernel
  float* losses;
  int n_scenes, S, G, n_shards, n_groups, n_part, T2;
  long long N;
  float w_max, a_max, thres, stl_weight, div_scale, div_weight, reg_w, extra_w;
  int diverse, detach;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- --diverse_loss: one warp per DPP group ---------------------------------------------------------------------
// shared memory per warp: s (G x (T2+1)) normalised samples, o (G x (T2+1)) their gradient, and four G x (G+1)
// matrices: Sim, Dist, Inv (M then M^-1), A (-(weight/groups) M^-2, then the pair coefficients)
__host__ __device__ inline int loss_warp_floats(int G, int T2) { return 2 * G * (T2 + 1) + 4 * G * (G + 1); }

__global__ void __launch_bounds__(128) k_refine_groups(LossArgs a, int wpb) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = a.G, LD = G + 1, T2 = a.T2, LS = T2 + 1;
  const int g = blockIdx.x * wpb + warp;
  if (g >= a.n_groups) return;  // warps are independent: no block-wide barrier below
  float* s = sm + (size_t)warp * loss_warp_floats(G, T2);
  float* o = s + G * LS;
  float* Sim = o + G * LS;
  float* Dist = Sim + G * LD;
  float* Inv = Dist + G * LD;
  float* A = Inv + G * LD;
  // group (b, m, shard) of reshape(bs,S,3,.).permute(0,2,1,3).reshape(bs*3*NS, S/NS, .) (:443)
  const int shard = g % a.n_shards, bm = g / a.n_shards, m = bm % 3, b = bm / 3;
  auto row_of = [&](int i) { return ((size_t)b * a.S + (size_t)shard * G + i) * 3 + m; };
  const bool grad = a.d_rect != nullptr;

  // samples / [w_max, a_max] (:444-445) and the masked regulariser partial (:465)
  float reg = 0.f;
  for (int e = lane; e < G * T2; e += 32) {
    const int i = e / T2, c = e - i * T2;
    const size_t r = row_of(i);
    const float rc = a.rect[r * T2 + c], d = rc - a.nn[r * T2 + c];
    s[i * LS + c] = rc / ((c & 1) ? a.a_max : a.w_max);
    if (a.scores[r] >= 0.f) reg += d * d;
  }
  float sc = 0.f, vl = 0.f, q = 0.f;
  if (lane < G) {
    const size_t r = row_of(lane);
    sc = a.scores[r];
    vl = a.valid[r];
    q = (sc > 0.f) ? (a.detach ? 1.f : expf(sc)) : 0.f;  // :449-452
  }
  const float hinge = (lane < G) ? fmaxf(a.thres - sc, 0.f) * vl : 0.f;
  const float maskc = (lane < G && sc >= 0.f) ? 1.f : 0.f;
  __syncwarp();

  // dist = ||s_i - s_j||, sim = exp(-scale dist) (:448-449); M = Q sim Q + I (:453-457)
  for (int j = 0; j < G; ++j) {
    const float qj = __shfl_sync(0xffffffffu, q, j);
    if (lane < G) {
      float d2 = 0.f;
      for (int c = 0; c < T2; ++c) {
        const float d = s[lane * LS + c] - s[j * LS + c];
        d2 = fmaf(d, d, d2);
      }
      const float dist = sqrtf(d2);
      const float sim = expf(-a.div_scale * dist);
      Dist[lane * LD + j] = dist;
      Sim[lane * LD + j] = sim;
      Inv[lane * LD + j] = q * sim * qj + (lane == j ? 1.f : 0.f);
    }
  }
  // in-place Gauss-Jordan: after step k column k holds the k-th column of the running inverse
  for (int k = 0; k < G; ++k) {
    __syncwarp();
    const float ip = 1.f / Inv[k * LD + k];
    __syncwarp();
    if (lane < G) Inv[k * LD + lane] = (lane == k) ? ip : Inv[k * LD + lane] * ip;
    __syncwarp();
    if (lane < G && lane != k) {
      const float f = Inv[lane * LD + k];
      for (int j = 0; j < G; ++j) Inv[lane * LD + j] = (j == k) ? -f * ip : fmaf(-f, Inv[k * LD + j], Inv[lane * LD + j]);
    }
  }
  __syncwarp();
  // diversity = tr(I - M^-1) (:458-459); the partial is its negative
  const float tr = warp_sum(lane < G ? Inv[lane * LD + lane] - 1.f : 0.f);
  reg = warp_sum(reg);
  const float hsum = warp_sum(hinge), vsum = warp_sum(vl), msum = warp_sum(maskc);
  if (lane == 0) {
    float* p = a.part + (size_t)g * kPart;
    p[0] = hsum; p[1] = vsum; p[2] = reg; p[3] = msum; p[4] = tr; p[5] = 0.f; p[6] = 0.f; p[7] = 0.f;
  }
  if (!grad) return;

  // d loss_div / dM = -(weight/groups) (M^-2)^T ; M = Q S Q + I
  const float coef = a.div_weight / (float)a.n_groups;
  float dq = 0.f;
  for (int j = 0; j < G; ++j) {
    const float qj = __shfl_sync(0xffffffffu, q, j);
    if (lane < G) {
      float acc = 0.f;
      for (int k = 0; k < G; ++k) acc = fmaf(Inv[lane * LD + k], Inv[k * LD + j], acc);
      const float aij = -coef * acc;
      const float sim = Sim[lane * LD + j], dist = Dist[lane * LD + j];
      dq = fmaf(2.f * aij * sim, qj, dq);  // M is symmetric: row and column contributions are equal
      // d/d dist_ij of both (i,j) and (j,i) entries, divided by dist for the norm's gradient (0 at dist = 0, as torch)
      A[lane * LD + j] = (dist > 0.f) ? 2.f * (-a.div_scale * sim * aij * q * qj) / dist : 0.f;
    }
  }
  if (lane < G) {
    for (int c = 0; c < T2; ++c) {
      const float si = s[lane * LS + c];
      float acc = 0.f;
      for (int j = 0; j < G; ++j) acc = fmaf(A[lane * LD + j], si - s[j * LS + c], acc);
      o[lane * LS + c] = acc;
    }
    a.d_scores[row_of(lane)] = a.detach ? 0.f : dq * q;  // d q / d quality = q where quality > 0
  }
  __syncwarp();
  for (int e = lane; e < G * T2; e += 32) {
    const int i = e / T2, c = e - i * T2;
    a.d_rect[row_of(i) * T2 + c] = o[i * LS + c] / ((c & 1) ? a.a_max : a.w_max);
  }
}

// ---- plain branch (:468-478): block partials of the channel regularisers, the box penalty and the STL hinge ------
__global__ void __launch_bounds__(256) k_refine_plain(LossArgs a) {
  __shared__ float red[8][kPart];
  float p[kPart] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long total = a.N * a.T2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / a.T2;
    const int c = (int)(e - r * a.T2);
    const float lim = (c & 1) ? a.a_max : a.w_max;
    const float rc = a.rect[e], d = (rc - a.nn[e]) / lim, u = rc / lim;
    p[2 + (c & 1)] += d * d;
    p[5 + (c & 1)] += fmaxf(u * u - 1.f, 0.f);
    if (c == 0) {
      const float vl = a.valid[r];
      p[0] += fmaxf(a.thres - a.scores[r], 0.f) * vl;
      p[1] += vl;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kPart; ++k) {
    const float v = warp_sum(p[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kPart) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    a.part[(size_t)blockIdx.x * kPart + threadIdx.x] = v;
  }
}

// ---- finalise: sums in fp64, fixed order; losses[0..4] = loss, loss_stl, loss_reg, loss_diversity, extra_loss_reg ---
__global__ void __launch_bounds__(256) k_refine_finalize(LossArgs a) {
  __shared__ double red[256][kPart];
  double p[kPart] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < a.n_part; i += 256)
    for (int k = 0; k < kPart; ++k) p[k] += (double)a.part[(size_t)i * kPart + k];
  for (int k = 0; k < kPart; ++k) red[threadIdx.x][k] = p[k];
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st)
      for (int k = 0; k < kPart; ++k) red[threadIdx.x][k] += red[threadIdx.x + st][k];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const double N = (double)a.N, E = N * a.T2;
  const double* s = red[0];
  const double vden = fmax(s[1] / N, 1e-2);  // mask_mean's clip (:23-27)
  const double loss_stl = (s[0] / N) / vden * a.stl_weight;
  double loss_reg, loss_div = 0.0, extra = 0.0, loss;
  float* f = a.fin;
  f[0] = (float)(a.stl_weight / (N * vden));  // d loss / d score_i = -f0 valid_i [thres - score_i > 0]
  if (a.diverse) {
    const double mden = fmax(s[3] / N, 1e-2);
    loss_reg = (s[2] / E) / mden;  // reported unscaled (:465), weighted in the total (:466)
    loss_div = s[4] / (double)a.n_groups * a.div_weight;
    loss = loss_stl + loss_reg * a.reg_w + loss_div;
    f[1] = (float)(a.reg_w * 2.0 / (E * mden));
    f[2] = f[3] = f[4] = 0.f;
  } else {
    const double half = E * 0.5;  // elements per channel
    loss_reg = (s[2] / half + s[3] / half) * a.reg_w;
    extra = (s[5] / half + s[6] / half) * a.extra_w;
    loss = loss_stl + loss_reg + extra;
    f[1] = (float)(a.reg_w * 2.0 / (half * (double)a.w_max * a.w_max));
    f[2] = (float)(a.reg_w * 2.0 / (half * (double)a.a_max * a.a_max));
    f[3] = (float)(a.extra_w * 2.0 / (half * (double)a.w_max * a.w_max));
    f[4] = (float)(a.extra_w * 2.0 / (half * (double)a.a_max * a.a_max));
  }
  a.losses[0] = (float)loss;
  a.losses[1] = (float)loss_stl;
  a.losses[2] = (float)loss_reg;
  a.losses[3] = (float)loss_div;
  a.losses[4] = (float)extra;
  a.losses[5] = (float)(s[1] / N);  // mean(valid)
  a.losses[6] = a.diverse ? (float)(s[3] / N) : 0.f;  // mean(score >= 0)
  a.losses[7] = 0.f;
}

// ---- gradient terms that needed the global denominators -----------------------------------------------------------
__global__ void __launch_bounds__(256) k_refine_grads(LossArgs a) {
  const long long total = a.N * a.T2;
  const float k_stl = a.fin[0], k1 = a.fin[1], k2 = a.fin[2], k3 = a.fin[3], k4 = a.fin[4];
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / a.T2;
    const int c = (int)(e - r * a.T2);
    const float rc = a.rect[e], d = rc - a.nn[e];
    if (a.diverse) {
      if (a.scores[r] >= 0.f) a.d_rect[e] += k1 * d;
    } else {
      const float lim = (c & 1) ? a.a_max : a.w_max, u = rc / lim;
      a.d_rect[e] = ((c & 1) ? k2 : k1) * d + ((u * u - 1.f > 0.f) ? ((c & 1) ? k4 : k3) * rc : 0.f);
    }
    if (c == 0) {
      const float gs = (a.thres - a.scores[r] > 0.f) ? -k_stl * a.valid[r] : 0.f;
      a.d_scores[r] = (a.diverse ? a.d_scores[r] : 0.f) + gs;
    }
  }
}

int plain_blocks() { return 148 * 4; }

}  // namespace

extern "C" size_t pstl_refine_losses_workspace_bytes(const pstl_loss_cfg* c) {
  if (!c || c->n_scenes <= 0) return 64;
  size_t n_part = plain_blocks();
  if (c->diverse_loss && c->n_shards > 0) n_part = (size_t)c->n_scenes * 3 * c->n_shards;
  return (n_part * kPart + 16) * sizeof(float);
}

extern "C" int pstl_refine_losses(const pstl_loss_cfg* c, const float* rect_controls, const float* nn_controls,
                                  const float* scores, const float* valid, float* losses, float* d_rect, float* d_scores,
                                  void* workspace, pstl_stream_t stream) {
  PSTL_CHECK_ARG(c && rect_controls && nn_controls && scores && valid && losses && workspace, "null argument");
  PSTL_CHECK_ARG((d_rect == nullptr) == (d_scores == nullptr), "d_rect and d_scores are given together or not at all");
  PSTL_CHECK_ARG(c->n_scenes > 0 && c->S > 0 && c->nt > 0, "empty batch");
  PSTL_CHECK_ARG(c->w_max > 0.f && c->a_max > 0.f, "control limits must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  LossArgs a{};
  a.rect = rect_controls; a.nn = nn_controls; a.scores = scores; a.valid = valid;
  a.d_rect = d_rect; a.d_scores = d_scores; a.losses = losses;
  a.n_scenes = c->n_scenes; a.S = c->S; a.T2 = 2 * c->nt;
  a.N = (long long)c->n_scenes * c->S * 3;
  a.w_max = c->w_max; a.a_max = c->a_max; a.thres = c->stl_nn_thres; a.stl_weight = c->stl_weight;
  a.div_scale = c->diversity_scale; a.div_weight = c->diversity_weight;
  a.reg_w = c->rect_reg_loss; a.extra_w = c->extra_rect_reg;
  a.diverse = c->diverse_loss != 0; a.detach = c->diverse_detach != 0;
  a.part = static_cast<float*>(workspace);
  if (a.diverse) {
    PSTL_CHECK_ARG(c->n_shards > 0 && c->S % c->n_shards == 0, "n_randoms must be a multiple of n_shards");
    a.n_shards = c->n_shards;
    a.G = c->S / c->n_shards;
    if (a.G > 32) {
      pstl_set_error("diversity loss: groups of %d samples (n_randoms / n_shards); at most 32 are built", a.G);
      return PSTL_ERR_UNSUPPORTED;
    }
    a.n_groups = c->n_scenes * 3 * c->n_shards;
    a.n_part = a.n_groups;
    a.fin = a.part + (size_t)a.n_part * kPart;
    const size_t per_warp = (size_t)loss_warp_floats(a.G, a.T2) * sizeof(float);
    int wpb = (int)((48 * 1024) / per_warp);
    if (wpb < 1) {
      pstl_set_error("diversity loss: a group of %d x %d floats does not fit shared memory", a.G, a.T2);
      return PSTL_ERR_UNSUPPORTED;
    }
    if (wpb > 4) wpb = 4;
    k_refine_groups<<<(a.n_groups + wpb - 1) / wpb, wpb * 32, wpb * per_warp, st>>>(a, wpb);
    PSTL_LAUNCH_CHECK();
  } else {
    a.n_part = plain_blocks();
    a.fin = a.part + (size_t)a.n_part * kPart;
    k_refine_plain<<<a.n_part, 256, 0, st>>>(a);
    PSTL_LAUNCH_CHECK();
  }
  k_refine_finalize<<<1, 256, 0, st>>>(a);
  PSTL_LAUNCH_CHECK();
  if (d_rect) {
    const long long total = a.N * a.T2;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    k_refine_grads<<<blocks, 256, 0, st>>>(a);
    PSTL_LAUNCH_CHECK();
  }
  return PSTL_OK;
}
