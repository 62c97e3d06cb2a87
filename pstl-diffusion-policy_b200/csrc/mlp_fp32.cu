// mlp_fp32.cu — fp32 SIMT path of the denoiser / sampler / RefineNet (the 1e-5 parity mode) and
// the host-side orchestration of the reverse loop.
//
// Algebra: of policy_net's 303 input columns, 224 (scene feature) are constant
// per scene and 32 (time embedding) constant per step, so
//   W1 . [feat, x, temb, hl, stlp] + b1 = (W1f . feat + b1)[scene] + (W1t . temb)[step] + W1p . [x, hl, stlp]
// and the per-step GEMM depth drops from 303 to 47.  The same split applies to rect_net (271 -> 47).
#include <math.h>

#include "mlp_common.cuh"

enum { EPI_PLAIN = 0, EPI_DDPM = 1, EPI_REFINE = 2 };

struct LinArgs {
  const float* X; int ldx;
  const float* W; int ldw;
  const float* bias;
  const float* rowbias; int rows_per_group; int ldrb;
  const float* bias2;
  const float* res; int ldres;
  float* Y; int ldy;
  int M, K, Nout, act;
  // EPI_DDPM
  float c1, c2, sqrt_beta;
  const float* z;
  unsigned long long seed, offset;
  const unsigned long long* offset_dev;
  int step, noise_mode;  // 0: none, 1: injected, 2: philox
  float* xio; int ldxio;
  float* mu_out;
  float* iter_out;
  float w_max, a_max;
  int clip;
  // EPI_REFINE
  const float* u0;
  const float* scores;
};

#define BN 64
#define BK 16

// Tiled SGEMM with fused epilogues.  TMR rows x 4 columns per thread (block tile 16*TMR x 64); the next K-slab is
// fetched into registers while the current one is multiplied.  Every output element is one fmaf chain over k in
// ascending order, whatever the tile shape, so results do not depend on the batch size or the tile chosen.
template <int EPI, int TMR>
__global__ void __launch_bounds__(256) k_linear(LinArgs a) {
  constexpr int BM_ = 16 * TMR;
  constexpr int APT = BM_ * BK / 256;  // A elements per thread per slab (8 or 2)
  __shared__ __align__(16) float As[BK][BM_ + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM_;
  const int n0 = blockIdx.y * BN;
  float acc[TMR][4];
#pragma unroll
  for (int i = 0; i < TMR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ar = tid / (BK / APT), ak = (tid % (BK / APT)) * APT;  // A: row ar, k-offset ak..ak+APT-1
  const int br = tid >> 2, bk = (tid & 3) * 4;                      // B: row br, k-offset bk..bk+3
  const long long am = m0 + ar;
  const int bn = n0 + br;
  float ra[APT], rb4[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < APT; ++j) {
      const int k = k0 + ak + j;
      ra[j] = (am < a.M && k < a.K) ? a.X[am * a.ldx + k] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + bk + j;
      rb4[j] = (bn < a.Nout && k < a.K) ? a.W[(long long)bn * a.ldw + k] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < a.K; k0 += BK) {
#pragma unroll
    for (int j = 0; j < APT; ++j) As[ak + j][ar] = ra[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[bk + j][br] = rb4[j];
    __syncthreads();
    if (k0 + BK < a.K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TMR];
#pragma unroll
      for (int i = 0; i < TMR; i += 2) {
        const float2 t2 = *reinterpret_cast<const float2*>(&As[kk][ty * TMR + i]);
        av[i] = t2.x; av[i + 1] = t2.y;
      }
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < TMR; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TMR; ++i) {
    const long long m = m0 + ty * TMR + i;
    if (m >= a.M) continue;
    const float* rb = a.rowbias ? a.rowbias + (m / a.rows_per_group) * a.ldrb : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.Nout) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[n];
      if (rb) v += rb[n];
      if (a.bias2) v += a.bias2[n];
      if (EPI == EPI_PLAIN) {
        if (a.res) v += a.res[m * a.ldres + n];
        if (a.act) v = fmaxf(v, 0.f);
        a.Y[m * a.ldy + n] = v;
      } else if (EPI == EPI_DDPM) {
        // nusc_model.py:162 eps = net + x ; nusc_train.py:587,628
        const float x = a.xio[m * a.ldxio + n];
        const float eps = v + x;
        const float mu = a.c2 * (x - a.c1 * eps);
        if (a.mu_out) {
          a.mu_out[m * a.Nout + n] = mu;
        } else {
          float z = 0.f;
          if (a.noise_mode == 1) z = a.z[m * a.Nout + n];
          else if (a.noise_mode == 2) z = pstl_noise_at(a.seed, a.offset + (a.offset_dev ? *a.offset_dev : 0ull), a.step, m, n);
          const float xn = mu + a.sqrt_beta * z;
          a.xio[m * a.ldxio + n] = xn;
          if (a.iter_out) {  // normalize_diff, nusc_train.py:647-655
            const float sc = (n & 1) ? a.a_max : a.w_max;
            float u = xn * sc;
            if (a.clip) u = fminf(fmaxf(u, -sc), sc);
            a.iter_out[m * a.Nout + n] = u;
          }
        }
      } else {  // EPI_REFINE, nusc_model.py:212-233
        const float r = tanhf(v);
        const float init = a.u0[m * a.Nout + n];
        const float lim = (n & 1) ? a.a_max : a.w_max;
        const float mk = (r >= 0.f) ? 1.f : 0.f;
        const float lo = r * (init - (-lim));
        const float hi = r * (lim - init);
        const float merged = lo * (1.f - mk) + hi * mk;
        const float viol = (a.scores[m] < 0.f) ? 1.f : 0.f;
        float o = init + merged * viol;
        if (a.clip) o = fminf(fmaxf(o, -lim), lim);
        a.Y[m * a.ldy + n] = o;
      }
    }
  }
}

static void lin_defaults(LinArgs& a) { memset(&a, 0, sizeof(a)); a.rows_per_group = 1; }

template <int EPI>
static int launch_linear(const LinArgs& a, cudaStream_t st) {
  if (a.M <= 0) return PSTL_OK;
  // small problems: 32-row tiles so the grid covers the SMs
  if ((long long)pstl_ceil_div(a.M, 128) * pstl_ceil_div(a.Nout, BN) < 296) {
    dim3 grid(pstl_ceil_div(a.M, 32), pstl_ceil_div(a.Nout, BN));
    k_linear<EPI, 2><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid(pstl_ceil_div(a.M, 128), pstl_ceil_div(a.Nout, BN));
    k_linear<EPI, 8><<<grid, 256, 0, st>>>(a);
  }
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

extern "C" int pstl_linear(const float* x, const float* w, const float* b, int M, int K, int Nout, int act, float* y,
                           pstl_stream_t stream) {
  PSTL_CHECK_ARG(x && w && y && K > 0 && Nout > 0, "bad argument");
  LinArgs a;
  lin_defaults(a);
  a.X = x; a.ldx = K; a.W = w; a.ldw = K; a.bias = b; a.Y = y; a.ldy = Nout; a.M = M; a.K = K; a.Nout = Nout; a.act = act;
  return launch_linear<EPI_PLAIN>(a, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------------------
// Three-layer ReLU MLP in one launch (the scene encoders, reference nusc_model.py:82-91: in -> 256 -> 256 -> out):
// a block owns 32 rows and keeps their activations in shared memory ([k][row], so a thread reads its rows of one
// k with float4 loads); the weights stream through 16-deep K slabs with the next slab prefetched into registers.
// Each output is one fmaf chain over k ascending with the bias added last — bit-identical to three k_linear calls.
// --------------------------------------------------------------------------------------
#define MLP3_ROWS 32
#define MLP3_H 256
#define MLP3_LDA (MLP3_ROWS + 4)
#define MLP3_LDW (MLP3_H + 4)

// One layer: acc[RT][4] = sum_k A[k][row(s)] * W[col][k] over K; W (NR, K) row-major in global memory, NR = 256
// (thread = 8 rows x 4 columns) or NR <= 32 padded to 32 (thread = 1 row x 4 columns).  Slab loads are coalesced:
// 16 consecutive threads read the 16 consecutive k of one weight row.
template <int NR, int RT>
__device__ __forceinline__ void mlp3_layer(const float* __restrict__ As, int K, const float* __restrict__ W, int n_rows,
                                           float* __restrict__ Ws, float (&acc)[RT][4]) {
  constexpr int PER = NR * BK / 256;  // slab elements per thread (16 or 2)
  constexpr int TXN = NR / 4;         // threads along the columns
  const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float rw[PER];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = tid + 256 * i, n = e / BK, kk = e % BK;
      rw[i] = (n < n_rows && k0 + kk < K) ? W[(size_t)n * K + k0 + kk] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    __syncthreads();  // previous slab fully consumed
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = tid + 256 * i;
      Ws[(e % BK) * MLP3_LDW + e / BK] = rw[i];
    }
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
    const int kn = (K - k0 < BK) ? K - k0 : BK;
#pragma unroll 4
    for (int kk = 0; kk < kn; ++kk) {
      float av[RT];
      if constexpr (RT == 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(As + (k0 + kk) * MLP3_LDA + ty * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(As + (k0 + kk) * MLP3_LDA + ty * 8 + 4);
        av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
        av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
      } else {
        av[0] = As[(k0 + kk) * MLP3_LDA + ty];
      }
      const float4 b0 = *reinterpret_cast<const float4*>(Ws + kk * MLP3_LDW + tx * 4);
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
}

struct Mlp3Problem {
  const float *x, *w0, *b0, *w2, *b2, *w4, *b4;
  float* y;
  int M, in_dim, out_dim, block0;  // block0: first block of this problem in the grid
};
struct Mlp3Batch {
  Mlp3Problem p[PSTL_MLP3_MAX];
  int n;
};

// up to PSTL_MLP3_MAX independent MLPs in one grid (the three scene encoders run side by side)
__global__ void __launch_bounds__(256) k_mlp3(const __grid_constant__ Mlp3Batch batch) {
  extern __shared__ __align__(16) float sm3[];
  float* As = sm3;                         // [256][MLP3_LDA]: layer input, k-major
  float* Ws = As + MLP3_H * MLP3_LDA;      // [16][MLP3_LDW]
  const int tid = threadIdx.x;
  int pi = 0;
  for (int i = 1; i < batch.n; ++i)
    if ((int)blockIdx.x >= batch.p[i].block0) pi = i;
  const Mlp3Problem& P = batch.p[pi];
  const float* __restrict__ x = P.x;
  const float *w0 = P.w0, *b0 = P.b0, *w2 = P.w2, *b2 = P.b2, *w4 = P.w4, *b4 = P.b4;
  float* __restrict__ y = P.y;
  const int M = P.M, in_dim = P.in_dim, out_dim = P.out_dim;
  const long long m0 = (long long)((int)blockIdx.x - P.block0) * MLP3_ROWS;
  for (int i = tid; i < MLP3_ROWS * in_dim; i += 256) {
    const int r = i / in_dim, k = i - r * in_dim;
    As[k * MLP3_LDA + r] = (m0 + r < M) ? x[(m0 + r) * in_dim + k] : 0.f;
  }
  // layers 1 and 2 (ReLU), outputs back into As as the next layer's input
#pragma unroll 1
  for (int layer = 0; layer < 2; ++layer) {
    float acc[8][4];
    const int tx = tid & 63, ty = tid >> 6;
    __syncthreads();
    mlp3_layer<MLP3_H, 8>(As, layer == 0 ? in_dim : MLP3_H, layer == 0 ? w0 : w2, MLP3_H, Ws, acc);
    const float* bias = layer == 0 ? b0 : b2;
    __syncthreads();  // every thread is done reading As
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = tx * 4 + j;
      const float bb = bias[n];
#pragma unroll
      for (int i = 0; i < 8; ++i) As[n * MLP3_LDA + ty * 8 + i] = fmaxf(acc[i][j] + bb, 0.f);
    }
  }
  __syncthreads();
  // layer 3: out_dim <= 32 columns, thread = (row tid/8, columns (tid%8)*4..+3)
  float acc3[1][4];
  mlp3_layer<32, 1>(As, MLP3_H, w4, out_dim, Ws, acc3);
  const int r = tid >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = (tid & 7) * 4 + j;
    if (n < out_dim && m0 + r < M) y[(m0 + r) * out_dim + n] = acc3[0][j] + b4[n];
  }
}

// y_i (M_i, out_i) = W4 relu(W2 relu(W0 x_i + b0) + b2) + b4 for n independent problems; hidden width 256
extern "C" int pstl_mlp3_batch(const pstl_mlp3_problem* probs, int n, pstl_stream_t stream) {
  PSTL_CHECK_ARG(probs && n >= 1 && n <= PSTL_MLP3_MAX, "1..PSTL_MLP3_MAX problems");
  Mlp3Batch b;
  memset(&b, 0, sizeof(b));
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    const pstl_mlp3_problem& q = probs[i];
    PSTL_CHECK_ARG(q.x && q.w0 && q.b0 && q.w2 && q.b2 && q.w4 && q.b4 && q.y, "null argument");
    PSTL_CHECK_ARG(q.hidden == MLP3_H && q.in_dim >= 1 && q.in_dim <= MLP3_H && q.out_dim >= 1 && q.out_dim <= 32,
                   "pstl_mlp3 is built for hidden = 256, out <= 32");
    if (q.M <= 0) continue;
    Mlp3Problem& p = b.p[b.n++];
    p.x = q.x; p.w0 = q.w0; p.b0 = q.b0; p.w2 = q.w2; p.b2 = q.b2; p.w4 = q.w4; p.b4 = q.b4; p.y = q.y;
    p.M = q.M; p.in_dim = q.in_dim; p.out_dim = q.out_dim; p.block0 = blocks;
    blocks += pstl_ceil_div(q.M, MLP3_ROWS);
  }
  if (!blocks) return PSTL_OK;
  const size_t smem = sizeof(float) * ((size_t)MLP3_H * MLP3_LDA + (size_t)BK * MLP3_LDW);
  PSTL_CUDA(cudaFuncSetAttribute(k_mlp3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mlp3<<<blocks, 256, smem, (cudaStream_t)stream>>>(b);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

extern "C" int pstl_mlp3(const float* x, int M, int in_dim, const float* w0, const float* b0, const float* w2,
                         const float* b2, const float* w4, const float* b4, int hidden, int out_dim, float* y,
                         pstl_stream_t stream) {
  pstl_mlp3_problem q;
  q.x = x; q.w0 = w0; q.b0 = b0; q.w2 = w2; q.b2 = b2; q.w4 = w4; q.b4 = b4; q.y = y;
  q.M = M; q.in_dim = in_dim; q.hidden = hidden; q.out_dim = out_dim;
  return pstl_mlp3_batch(&q, 1, stream);
}

// --------------------------------------------------------------------------------------
// small helper kernels
// --------------------------------------------------------------------------------------
// xin[n] = [x (T2) | hl | stlp(6) | 0]
// draw != 0: x is not given and x_T ~ N(0,1) comes from the sampler's Philox stream at step word `draw_step`
// (= steps, which no reverse step uses).  One thread per group of four columns: one Philox call yields the four
// normals of a group (same counter layout as pstl_noise_at), and the row is written with 16-byte stores.
__global__ void k_pack_xin(const float* __restrict__ x, int ldx_src, const float* __restrict__ hl,
                           const float* __restrict__ stlp, float* __restrict__ xin, long long N, int T2, int draw,
                           unsigned long long seed, unsigned long long offset,
                           const unsigned long long* __restrict__ offset_dev, int draw_step) {
  constexpr int Q = PSTL_XIN_LD / 4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Q) return;
  const long long n = i / Q;
  const int c0 = (int)(i - n * Q) * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (draw && c0 + 3 < T2) {
    const unsigned long long off = offset + (offset_dev ? *offset_dev : 0ull);
    const uint4 ctr = make_uint4((unsigned)(n & 0xffffffff), (unsigned)(n >> 32), (unsigned)(c0 >> 2),
                                 (unsigned)draw_step + (unsigned)off);
    const uint4 r = pstl_philox(ctr, make_uint2((unsigned)(seed & 0xffffffff), (unsigned)(seed >> 32)));
    pstl_box_muller(r.x, r.y, v[0], v[1]);
    pstl_box_muller(r.z, r.w, v[2], v[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c0 + e;
      if (c < T2) {
        if (draw) v[e] = pstl_noise_at(seed, offset + (offset_dev ? *offset_dev : 0ull), draw_step, n, c);
        else v[e] = x ? x[n * ldx_src + c] : 0.f;
      } else if (c == T2) {
        v[e] = hl[n];
      } else if (c < T2 + 7) {
        v[e] = stlp[n * 6 + (c - T2 - 1)];
      }
    }
  }
  *reinterpret_cast<float4*>(xin + n * PSTL_XIN_LD + c0) = make_float4(v[0], v[1], v[2], v[3]);
}

// RefineNet shard pooling (nusc_model.py:186-200): rows n=(b*R+r)*3+m; max over the `per` samples of a shard
__global__ void k_group_fuse(const float* __restrict__ g, const float* __restrict__ u0, float* __restrict__ xin,
                             long long N, int T2, int R, int per) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over (group, col)
  const long long n_groups = N / per;                                    // (b, shard, m) triples
  if (i >= n_groups * T2) return;
  const long long gidx = i / T2;
  const int c = (int)(i - gidx * T2);
  const int m = (int)(gidx % 3);
  const long long bs_sh = gidx / 3;            // b * n_shards + shard
  const int n_shards = R / per;
  const long long b = bs_sh / n_shards;
  const int sh = (int)(bs_sh - b * n_shards);
  float mx = -INFINITY;
  for (int j = 0; j < per; ++j) {
    const long long n = (b * R + sh * per + j) * 3 + m;
    mx = fmaxf(mx, g[n * T2 + c]);
  }
  for (int j = 0; j < per; ++j) {
    const long long n = (b * R + sh * per + j) * 3 + m;
    xin[n * PSTL_XIN_LD + c] = u0[n * T2 + c] + mx;
  }
}

// RefineNet front end in one launch (nusc_model.py:186-204): merge_net (T2 -> 32 -> 32 -> T2, ReLU) on every chain of
// one scene, max over the `per` samples of each (shard, mode) group, fused = u0 + pooled, packed with [hl | stlp | 0]
// into the engine's input rows.  One block per scene, one thread per chain for the MLP (weights and the scene's
// controls in shared memory), then a coalesced cooperative write of the packed rows.  The sums run k-ascending
// through fmaf with the bias added last, exactly like k_linear.
#define PSTL_MERGE_H 32
__global__ void __launch_bounds__(256) k_merge_fuse(const float* __restrict__ u0, const float* __restrict__ w0,
                                                    const float* __restrict__ b0, const float* __restrict__ w2,
                                                    const float* __restrict__ b2, const float* __restrict__ w4,
                                                    const float* __restrict__ b4, const float* __restrict__ hl,
                                                    const float* __restrict__ stlp, float* __restrict__ xin, int T2, int R,
                                                    int per) {
  extern __shared__ __align__(16) float smf[];
  constexpr int MH = PSTL_MERGE_H;
  const int rows = 3 * R, ld = T2 + 1;
  const int T2p = (T2 + 3) & ~3;
  float* W0t = smf;                 // [T2][MH]   transposed: the inner index is the output unit (LDS.128 = 4 units)
  float* W2t = W0t + T2 * MH;       // [MH][MH]
  float* W4t = W2t + MH * MH;       // [MH][T2p]
  float* B0 = W4t + MH * T2p;       // MH
  float* B2 = B0 + MH;              // MH
  float* B4 = B2 + MH;              // T2p
  float* X = B4 + T2p;              // [rows][ld]: controls in, merge_net output over them (row r belongs to thread r)
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * rows;
  // all global reads are float4 and issued before their first use (T2 % 4 == 0, checked by the launcher)
  const int T2q = T2 / 4;
#pragma unroll 4
  for (int i = tid; i < MH * T2q; i += blockDim.x) {  // w0 (MH, T2) row-major -> W0t[k][j]
    const float4 v = reinterpret_cast<const float4*>(w0)[i];
    const int j = i / T2q, k = (i - j * T2q) * 4;
    W0t[k * MH + j] = v.x; W0t[(k + 1) * MH + j] = v.y; W0t[(k + 2) * MH + j] = v.z; W0t[(k + 3) * MH + j] = v.w;
  }
#pragma unroll 4
  for (int i = tid; i < T2 * (MH / 4); i += blockDim.x) {  // w4 (T2, MH) row-major -> W4t[in][c]
    const float4 v = reinterpret_cast<const float4*>(w4)[i];
    const int c = i / (MH / 4), j = (i - c * (MH / 4)) * 4;
    W4t[j * T2p + c] = v.x; W4t[(j + 1) * T2p + c] = v.y; W4t[(j + 2) * T2p + c] = v.z; W4t[(j + 3) * T2p + c] = v.w;
  }
  for (int i = tid; i < MH * (MH / 4); i += blockDim.x) {  // w2 (MH, MH) -> W2t[k][j]
    const float4 v = reinterpret_cast<const float4*>(w2)[i];
    const int j = i / (MH / 4), k = (i - j * (MH / 4)) * 4;
    W2t[k * MH + j] = v.x; W2t[(k + 1) * MH + j] = v.y; W2t[(k + 2) * MH + j] = v.z; W2t[(k + 3) * MH + j] = v.w;
  }
  for (int i = tid; i < MH; i += blockDim.x) { B0[i] = b0[i]; B2[i] = b2[i]; }
  for (int i = tid; i < T2; i += blockDim.x) B4[i] = b4[i];
  const float4* u4 = reinterpret_cast<const float4*>(u0 + row0 * T2);
#pragma unroll 8
  for (int i = tid; i < rows * T2q; i += blockDim.x) {
    const float4 v = u4[i];
    const int r = i / T2q, k = (i - r * T2q) * 4;
    float* x = X + r * ld + k;
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  }
  __syncthreads();
  for (int r = tid; r < rows; r += blockDim.x) {
    float h1[MH], h2[MH];
#pragma unroll
    for (int j = 0; j < MH; ++j) { h1[j] = 0.f; h2[j] = 0.f; }
    for (int k = 0; k < T2; ++k) {
      const float x = X[r * ld + k];
#pragma unroll
      for (int j = 0; j < MH; j += 4) {
        const float4 w = *reinterpret_cast<const float4*>(W0t + k * MH + j);
        h1[j] = fmaf(x, w.x, h1[j]); h1[j + 1] = fmaf(x, w.y, h1[j + 1]);
        h1[j + 2] = fmaf(x, w.z, h1[j + 2]); h1[j + 3] = fmaf(x, w.w, h1[j + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < MH; ++j) h1[j] = fmaxf(h1[j] + B0[j], 0.f);
#pragma unroll
    for (int k = 0; k < MH; ++k) {
#pragma unroll
      for (int j = 0; j < MH; j += 4) {
        const float4 w = *reinterpret_cast<const float4*>(W2t + k * MH + j);
        h2[j] = fmaf(h1[k], w.x, h2[j]); h2[j + 1] = fmaf(h1[k], w.y, h2[j + 1]);
        h2[j + 2] = fmaf(h1[k], w.z, h2[j + 2]); h2[j + 3] = fmaf(h1[k], w.w, h2[j + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < MH; ++j) h2[j] = fmaxf(h2[j] + B2[j], 0.f);
    for (int c = 0; c < T2; c += 4) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int k = 0; k < MH; ++k) {
        const float4 w = *reinterpret_cast<const float4*>(W4t + k * T2p + c);
        a0 = fmaf(h2[k], w.x, a0); a1 = fmaf(h2[k], w.y, a1); a2 = fmaf(h2[k], w.z, a2); a3 = fmaf(h2[k], w.w, a3);
      }
      X[r * ld + c] = a0 + B4[c];
      if (c + 1 < T2) X[r * ld + c + 1] = a1 + B4[c + 1];
      if (c + 2 < T2) X[r * ld + c + 2] = a2 + B4[c + 2];
      if (c + 3 < T2) X[r * ld + c + 3] = a3 + B4[c + 3];
    }
  }
  __syncthreads();
  constexpr int LQ = PSTL_XIN_LD / 4;
  float4* o4 = reinterpret_cast<float4*>(xin + row0 * PSTL_XIN_LD);
#pragma unroll 4
  for (int i = tid; i < rows * LQ; i += blockDim.x) {
    const int r = i / LQ, c = (i - r * LQ) * 4;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < T2) {
      const float4 uu = u4[r * T2q + c / 4];
      const int ri = r / 3, m = r - ri * 3, sh = ri / per;
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      for (int j = 0; j < per; ++j) {
        const float* g = X + ((sh * per + j) * 3 + m) * ld + c;
#pragma unroll
        for (int e = 0; e < 4; ++e) mx[e] = fmaxf(mx[e], g[e]);
      }
      o[0] = uu.x + mx[0]; o[1] = uu.y + mx[1]; o[2] = uu.z + mx[2]; o[3] = uu.w + mx[3];
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cc = c + e;
        if (cc == T2) o[e] = hl[row0 + r];
        else if (cc < T2 + 7) o[e] = stlp[(row0 + r) * 6 + (cc - T2 - 1)];
      }
    }
    o4[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void k_finish_step(const float* __restrict__ mu, float* __restrict__ xin, const float* __restrict__ z,
                              long long N, int T2, float sqrt_beta, int noise_mode, unsigned long long seed,
                              unsigned long long offset, const unsigned long long* __restrict__ offset_dev, int step,
                              float* __restrict__ iter_out, float w_max, float a_max, int clip) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * T2) return;
  const long long n = i / T2;
  const int c = (int)(i - n * T2);
  float zz = 0.f;
  if (noise_mode == 1) zz = z[i];
  else if (noise_mode == 2) zz = pstl_noise_at(seed, offset + (offset_dev ? *offset_dev : 0ull), step, n, c);
  const float xn = mu[i] + sqrt_beta * zz;
  xin[n * PSTL_XIN_LD + c] = xn;
  if (iter_out) {
    const float sc = (c & 1) ? a_max : w_max;
    float u = xn * sc;
    if (clip) u = fminf(fmaxf(u, -sc), sc);
    iter_out[i] = u;
  }
}

__global__ void k_extract_x(const float* __restrict__ xin, float* __restrict__ out, long long N, int T2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * T2) return;
  const long long n = i / T2;
  out[i] = xin[n * PSTL_XIN_LD + (int)(i - n * T2)];
}

// --------------------------------------------------------------------------------------
// handle
// --------------------------------------------------------------------------------------
int pstl_tc_create(pstl_denoiser* d);   // denoiser_tc.cu
void pstl_tc_destroy(pstl_denoiser* d);
int pstl_tc_sample(pstl_denoiser* d, const float* cscene, int rows_per_scene, const float* ct, float* xin, int N,
                   const float* sched, int steps, const float* noise, unsigned long long seed,
                   unsigned long long offset, float w_max, float a_max, int clip, int keep_last_k, float* iterates_out,
                   int first_step, int last_step, float* mu_out, cudaStream_t st);
int pstl_tc_refine(pstl_denoiser* d, const float* cscene, int rows_per_scene, const float* xin, int N, const float* u0,
                   const float* scores, float w_max, float a_max, int clip_rect, float* out, cudaStream_t st);
bool pstl_tc_has_refine(pstl_denoiser* d);
bool pstl_tc_fits(pstl_denoiser* d, int rows_per_scene);  // the handle's engine can tile rows_per_scene

// the hoisted first-layer column blocks the per-step GEMMs read (copies of the caller's weights)
static cudaError_t snapshot_first_layers(pstl_denoiser* d, cudaStream_t st) {
  const pstl_weights* w = &d->w;
  const int in1 = w->feat_dim + d->T2 + w->time_dim + 7;  // 303
  const size_t f = sizeof(float);
  // [x | hl | stlp] <- columns [feat : feat+T2], [feat+T2+time], [feat+T2+time+1 : +7]
  cudaError_t e = cudaMemcpy2DAsync(d->w1p, d->kin * f, w->p0_w + w->feat_dim, in1 * f, d->T2 * f, w->hidden,
                                    cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(d->w1p + d->T2, d->kin * f, w->p0_w + w->feat_dim + d->T2 + w->time_dim, in1 * f, 7 * f,
                          w->hidden, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && d->r1p) {
    const int inr = w->feat_dim + 7 + d->T2;  // 271: [feat | hl | stlp | fused]
    e = cudaMemcpy2DAsync(d->r1p, d->kin * f, w->r0_w + w->feat_dim + 7, inr * f, d->T2 * f, w->rect_hidden,
                          cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(d->r1p + d->T2, d->kin * f, w->r0_w + w->feat_dim, inr * f, 7 * f, w->rect_hidden,
                            cudaMemcpyDeviceToDevice, st);
  }
  return e;
}

int pstl_tc_refresh(pstl_denoiser* d, cudaStream_t st);  // denoiser_tc.cu

extern "C" int pstl_denoiser_refresh(pstl_denoiser_t d, pstl_stream_t stream) {
  PSTL_CHECK_ARG(d, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = snapshot_first_layers(d, st);
  if (e != cudaSuccess) {
    pstl_set_error("pstl_denoiser_refresh: %s", cudaGetErrorString(e));
    return PSTL_ERR_CUDA;
  }
  return d->tc ? pstl_tc_refresh(d, st) : PSTL_OK;
}

extern "C" int pstl_denoiser_create(const pstl_weights* w, int precision, pstl_denoiser_t* out) {
  PSTL_CHECK_ARG(w && out, "null argument");
  PSTL_CHECK_ARG(w->p0_w && w->p0_b && w->p2_w && w->p2_b && w->p4_w && w->p4_b, "policy_net weights required");
  PSTL_CHECK_ARG(w->T > 0 && w->hidden > 0 && w->feat_dim > 0 && w->time_dim > 0, "bad dims");
  PSTL_CHECK_ARG(precision == PSTL_PRECISION_FP32 || precision == PSTL_PRECISION_BF16 || precision == PSTL_PRECISION_F16X3 ||
                 precision == PSTL_PRECISION_F16, "bad precision");
  if (2 * w->T + 7 > PSTL_XIN_LD) {  // the packed input row [x | hl | stlp] has a fixed leading dimension
    pstl_set_error("pstl_denoiser_create: T = %d needs 2T+7 <= %d packed input columns", w->T, PSTL_XIN_LD);
    return PSTL_ERR_UNSUPPORTED;
  }
  pstl_denoiser* d = new pstl_denoiser();
  d->w = *w;
  d->precision = precision;
  d->T2 = 2 * w->T;
  d->kin = d->T2 + 7;
  d->w1p = d->r1p = nullptr;
  d->tc = nullptr;
  d->offset_dev = nullptr;
  const size_t f = sizeof(float);
  cudaError_t e = cudaMalloc(&d->w1p, (size_t)w->hidden * d->kin * f);
  if (e == cudaSuccess && w->r0_w) e = cudaMalloc(&d->r1p, (size_t)w->rect_hidden * d->kin * f);
  if (e == cudaSuccess) e = snapshot_first_layers(d, nullptr);
  if (e != cudaSuccess) {
    pstl_set_error("pstl_denoiser_create: %s", cudaGetErrorString(e));
    pstl_denoiser_destroy(d);
    return PSTL_ERR_CUDA;
  }
  if (precision != PSTL_PRECISION_FP32) {
    int rc = pstl_tc_create(d);
    if (rc) {
      pstl_denoiser_destroy(d);
      return rc;
    }
  }
  *out = d;
  return PSTL_OK;
}

extern "C" int pstl_denoiser_destroy(pstl_denoiser_t d) {
  if (!d) return PSTL_OK;
  if (d->tc) pstl_tc_destroy(d);
  cudaFree(d->w1p);
  cudaFree(d->r1p);
  delete d;
  return PSTL_OK;
}

extern "C" int pstl_denoiser_set_engine(pstl_denoiser_t d, int engine) {
  PSTL_CHECK_ARG(d, "null handle");
  PSTL_CHECK_ARG(engine >= 0 && engine <= 2, "engine must be 0 (automatic), 1 (one-SM) or 2 (CTA pair)");
  d->tc_engine = engine;
  return PSTL_OK;
}

extern "C" int pstl_denoiser_set_noise_counter(pstl_denoiser_t d, const uint64_t* device_counter) {
  PSTL_CHECK_ARG(d, "null handle");
  d->offset_dev = reinterpret_cast<const unsigned long long*>(device_counter);
  return PSTL_OK;
}

static size_t guidance_ws_floats(pstl_denoiser_t d, int N, const pstl_guidance_cfg* g) {
  if (!g) return 0;
  size_t tape = pstl_score_workspace_bytes(g->progs, N, d->w.T, 1) / sizeof(float);
  return pstl_align_floats((size_t)N * d->T2) + pstl_align_floats(tape);
}

static void carve(pstl_denoiser_t d, int N, int n_scenes, int steps, const pstl_guidance_cfg* g, void* ws,
                  DenoiserWs* o, size_t* total) {
  const int H = d->w.hidden > d->w.rect_hidden ? d->w.hidden : d->w.rect_hidden;
  float* p = (float*)ws;
  size_t off = 0;
  auto take = [&](size_t n) { float* r = p ? p + off : nullptr; off += pstl_align_floats(n); return r; };
  o->xin = take((size_t)N * PSTL_XIN_LD);
  o->h1 = take((size_t)N * H);
  o->h2 = take((size_t)N * H);
  o->g = take((size_t)N * d->T2);
  o->cscene = take((size_t)n_scenes * H);
  o->ct = take((size_t)(steps > 1 ? steps : 1) * H);
  o->adam = g ? take((size_t)3 * N * d->T2) : nullptr;
  o->gws = g ? take(guidance_ws_floats(d, N, g)) : nullptr;
  *total = off * sizeof(float);
}

extern "C" size_t pstl_denoiser_workspace_bytes(pstl_denoiser_t d, int N, int n_scenes, const pstl_guidance_cfg* g) {
  if (!d) return 0;
  DenoiserWs w;
  size_t total;
  carve(d, N, n_scenes, PSTL_MAX_STEPS, g, nullptr, &w, &total);
  return total;
}

// hoisted first-layer terms: cscene = feat . W[:, :feat]^T + b ; ct = temb . W[:, tcol:tcol+time]^T
static int hoist(const float* W, int ldw, const float* b, int H, const float* feat, int n_scenes, int feat_dim,
                 float* cscene, const float* temb, int steps, int tcol, int time_dim, float* ct, cudaStream_t st) {
  LinArgs a;
  lin_defaults(a);
  a.X = feat; a.ldx = feat_dim; a.W = W; a.ldw = ldw; a.bias = b; a.Y = cscene; a.ldy = H;
  a.M = n_scenes; a.K = feat_dim; a.Nout = H;
  int rc = launch_linear<EPI_PLAIN>(a, st);
  if (rc || !temb) return rc;
  lin_defaults(a);
  a.X = temb; a.ldx = time_dim; a.W = W + tcol; a.ldw = ldw; a.Y = ct; a.ldy = H; a.M = steps; a.K = time_dim; a.Nout = H;
  return launch_linear<EPI_PLAIN>(a, st);
}

static int mlp_hidden(pstl_denoiser_t d, const DenoiserWs& w, int N, int rows_per_scene, const float* W1p, int H,
                      const float* ct_row, const float* W2, const float* b2, cudaStream_t st) {
  LinArgs a;
  lin_defaults(a);
  a.X = w.xin; a.ldx = PSTL_XIN_LD; a.W = W1p; a.ldw = d->kin; a.rowbias = w.cscene; a.rows_per_group = rows_per_scene;
  a.ldrb = H; a.bias2 = ct_row; a.Y = w.h1; a.ldy = H; a.M = N; a.K = d->kin; a.Nout = H; a.act = 1;
  int rc = launch_linear<EPI_PLAIN>(a, st);
  if (rc) return rc;
  lin_defaults(a);
  a.X = w.h1; a.ldx = H; a.W = W2; a.ldw = H; a.bias = b2; a.Y = w.h2; a.ldy = H; a.M = N; a.K = H; a.Nout = H; a.act = 1;
  return launch_linear<EPI_PLAIN>(a, st);
}

extern "C" int pstl_denoiser_eps(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                                 const float* hl, const float* stlp, const float* x, int N, const float* temb_row,
                                 float* eps_out, void* workspace, pstl_stream_t stream) {
  PSTL_CHECK_ARG(d && scene_feat && hl && stlp && x && temb_row && eps_out && workspace, "null argument");
  PSTL_CHECK_ARG(rows_per_scene >= 1 && (long long)n_scenes * rows_per_scene >= N, "bad scene mapping");
  cudaStream_t st = (cudaStream_t)stream;
  DenoiserWs w;
  size_t total;
  carve(d, N, n_scenes, 1, nullptr, workspace, &w, &total);
  const int H = d->w.hidden, T2 = d->T2;
  const int in1 = d->w.feat_dim + T2 + d->w.time_dim + 7;
  int rc = hoist(d->w.p0_w, in1, d->w.p0_b, H, scene_feat, n_scenes, d->w.feat_dim, w.cscene, temb_row, 1,
                 d->w.feat_dim + T2, d->w.time_dim, w.ct, st);
  if (rc) return rc;
  const long long tot = (long long)N * (PSTL_XIN_LD / 4);
  k_pack_xin<<<pstl_ceil_div(tot, 256), 256, 0, st>>>(x, T2, hl, stlp, w.xin, N, T2, 0, 0ull, 0ull, nullptr, 0);
  PSTL_LAUNCH_CHECK();
  rc = mlp_hidden(d, w, N, rows_per_scene, d->w1p, H, w.ct, d->w.p2_w, d->w.p2_b, st);
  if (rc) return rc;
  LinArgs a;
  lin_defaults(a);
  a.X = w.h2; a.ldx = H; a.W = d->w.p4_w; a.ldw = H; a.bias = d->w.p4_b; a.res = x; a.ldres = T2;
  a.Y = eps_out; a.ldy = T2; a.M = N; a.K = H; a.Nout = T2;
  return launch_linear<EPI_PLAIN>(a, st);
}

extern "C" int pstl_denoiser_sample(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                                    const float* hl, const float* stlp, int N, const float* sched, const float* temb,
                                    int steps, const float* x_init, const float* noise, uint64_t seed, uint64_t offset,
                                    float w_max, float a_max, int clip, int keep_last_k,
                                    const pstl_guidance_cfg* guidance, float* iterates_out, float* x_final,
                                    void* workspace, pstl_stream_t stream) {
  PSTL_CHECK_ARG(d && scene_feat && hl && stlp && sched && temb && workspace, "null argument");
  PSTL_CHECK_ARG(x_init || !noise, "injected noise needs x_init");
  PSTL_CHECK_ARG(steps >= 2 && keep_last_k >= 0 && keep_last_k <= steps - 1, "bad steps / keep_last_k");
  PSTL_CHECK_ARG(steps <= PSTL_MAX_STEPS, "steps exceeds PSTL_MAX_STEPS (the workspace's per-step table)");
  PSTL_CHECK_ARG(rows_per_scene >= 1 && (long long)n_scenes * rows_per_scene >= N, "bad scene mapping");
  PSTL_CHECK_ARG(!keep_last_k || iterates_out, "iterates_out required");
  if (N <= 0) return PSTL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DenoiserWs w;
  size_t total;
  carve(d, N, n_scenes, steps, guidance, workspace, &w, &total);
  const int H = d->w.hidden, T2 = d->T2;
  const int in1 = d->w.feat_dim + T2 + d->w.time_dim + 7;
  int rc = hoist(d->w.p0_w, in1, d->w.p0_b, H, scene_feat, n_scenes, d->w.feat_dim, w.cscene, temb, steps,
                 d->w.feat_dim + T2, d->w.time_dim, w.ct, st);
  if (rc) return rc;
  const long long tot = (long long)N * (PSTL_XIN_LD / 4);
  k_pack_xin<<<pstl_ceil_div(tot, 256), 256, 0, st>>>(x_init, T2, hl, stlp, w.xin, N, T2, x_init ? 0 : 1, seed, offset,
                                                          d->offset_dev, steps);
  PSTL_LAUNCH_CHECK();
  const float* hs = sched;  // host pointer by contract
  const float *beta = hs, *alpha = hs + steps, *abar = hs + 2 * steps;
  const size_t NT2 = (size_t)N * T2;

  // which reverse steps are guided (nusc_train.py:589-598): the caller's per-step mask, or i <= before
  auto guided_at = [&](int i) {
    return guidance && (guidance->step_mask ? guidance->step_mask[i] != 0 : i <= guidance->before);
  };
  // the tcgen05 engines run the unguided steps (and the posterior mean of guided ones); a split-operand handle whose
  // 256-row tile does not fit rows_per_scene falls back to the fp32 SIMT chain
  const bool tc_guided = d->precision == PSTL_PRECISION_BF16 || d->precision == PSTL_PRECISION_F16 ||
                         (d->precision == PSTL_PRECISION_F16X3 && pstl_tc_fits(d, rows_per_scene));
  for (int i = steps - 1; i >= 1; --i) {
    const bool guided = guided_at(i);
    if (tc_guided && !guided) {
      // a maximal run of unguided steps [lo, i] is ONE launch of the tcgen05 engine; guided steps need mu
      // materialised and go through the single-step path below
      int lo = i;
      while (lo - 1 >= 1 && !guided_at(lo - 1)) --lo;
      rc = pstl_tc_sample(d, w.cscene, rows_per_scene, w.ct, w.xin, N, sched, steps, noise, seed, offset, w_max, a_max,
                          clip, keep_last_k, iterates_out, i, lo, nullptr, st);
      if (rc) break;
      i = lo;
      continue;
    }
    if (!(guided && tc_guided)) {
      rc = mlp_hidden(d, w, N, rows_per_scene, d->w1p, H, w.ct + (size_t)i * H, d->w.p2_w, d->w.p2_b, st);
      if (rc) break;
    }
    const int zi = steps - 1 - i;  // index into the injected z stream (i = steps-1 first)
    const int noise_mode = (i > 1) ? (noise ? 1 : 2) : 0;
    const int kidx = keep_last_k - i;  // iterate after step i is the (i)-th from the end
    float* it_out = (kidx >= 0 && keep_last_k > 0) ? iterates_out + (size_t)kidx * NT2 : nullptr;
    LinArgs a;
    lin_defaults(a);
    a.X = w.h2; a.ldx = H; a.W = d->w.p4_w; a.ldw = H; a.bias = d->w.p4_b; a.M = N; a.K = H; a.Nout = T2;
    a.c1 = (1.0f - alpha[i]) / sqrtf(1.0f - abar[i]);
    a.c2 = 1.0f / sqrtf(alpha[i]);
    a.sqrt_beta = sqrtf(beta[i]);
    a.z = (noise && i > 1) ? noise + (size_t)zi * NT2 : nullptr;
    a.seed = seed; a.offset = offset; a.offset_dev = d->offset_dev; a.step = i; a.noise_mode = noise_mode;
    a.xio = w.xin; a.ldxio = PSTL_XIN_LD;
    a.mu_out = guided ? w.g : nullptr;
    a.iter_out = guided ? nullptr : it_out;
    a.w_max = w_max; a.a_max = a_max; a.clip = clip;
    if (guided && tc_guided)  // one reverse step on the tcgen05 engine, posterior mean into w.g
      rc = pstl_tc_sample(d, w.cscene, rows_per_scene, w.ct, w.xin, N, sched, steps, nullptr, seed, offset, w_max, a_max, clip,
                          0, nullptr, i, i, w.g, st);
    else
      rc = launch_linear<EPI_DDPM>(a, st);
    if (rc) break;
    if (guided) {
      float* m = w.adam;
      float* v = w.adam + NT2;
      float* anchor = w.adam + 2 * NT2;
      cudaError_t e = cudaMemsetAsync(m, 0, sizeof(float) * 2 * NT2, st);
      if (e != cudaSuccess) { rc = PSTL_ERR_CUDA; pstl_set_error("memset: %s", cudaGetErrorString(e)); break; }
      for (int j = 0; j < guidance->niters && !rc; ++j)
        rc = pstl_guidance_step(guidance->progs, guidance->scenes, guidance->sp, hl, guidance->state0, stlp,
                                guidance->valid, N, guidance->thres, guidance->inv_norm, guidance->inv_norm_dev, guidance->lr, beta[i], j, w.g,
                                m, v, anchor, w.gws, stream);
      if (rc) break;
      k_finish_step<<<pstl_ceil_div((long long)NT2, 256), 256, 0, st>>>(w.g, w.xin, a.z, N, T2, a.sqrt_beta, noise_mode,
                                                                       seed, offset, d->offset_dev, i, it_out, w_max, a_max, clip);
      cudaError_t le = cudaGetLastError();
      if (le != cudaSuccess) { rc = PSTL_ERR_CUDA; pstl_set_error("k_finish_step: %s", cudaGetErrorString(le)); break; }
    }
  }
  if (rc) return rc;
  if (x_final) {
    k_extract_x<<<pstl_ceil_div((long long)NT2, 256), 256, 0, st>>>(w.xin, x_final, N, T2);
    PSTL_LAUNCH_CHECK();
  }
  return PSTL_OK;
}

// RefineNet input stage: hoisted scene term of rect_net.0, merge_net + shard max-pool + fuse -> packed rows w.xin
static int refine_inputs(pstl_denoiser_t d, const DenoiserWs& w, const float* scene_feat, int n_scenes, const float* hl,
                         const float* stlp, const float* u0, int N, int n_randoms, int n_shards, cudaStream_t st) {
  const int H = d->w.rect_hidden, T2 = d->T2, MH = d->w.merge_hidden;
  const int inr = d->w.feat_dim + 7 + T2;
  int rc = hoist(d->w.r0_w, inr, d->w.r0_b, H, scene_feat, n_scenes, d->w.feat_dim, w.cscene, nullptr, 0, 0, 0, nullptr, st);
  if (rc) return rc;
  const int per = n_randoms / n_shards;
  const int T2p = (T2 + 3) & ~3;
  const size_t merge_smem = sizeof(float) * ((size_t)MH * T2 + MH * MH + (size_t)MH * T2p + 2 * MH + T2p + (size_t)3 * n_randoms * (T2 + 1));
  LinArgs a;
  if (MH == PSTL_MERGE_H && T2 % 4 == 0 && merge_smem <= 160 * 1024) {
    // merge_net + shard max-pool + fuse + pack: one launch, one block per scene
    PSTL_CUDA(cudaFuncSetAttribute(k_merge_fuse, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    k_merge_fuse<<<N / (3 * n_randoms), 256, merge_smem, st>>>(u0, d->w.m0_w, d->w.m0_b, d->w.m2_w, d->w.m2_b, d->w.m4_w,
                                                            d->w.m4_b, hl, stlp, w.xin, T2, n_randoms, per);
    PSTL_LAUNCH_CHECK();
  } else {
    // merge_net on every chain (nusc_model.py:186)
    lin_defaults(a);
    a.X = u0; a.ldx = T2; a.W = d->w.m0_w; a.ldw = T2; a.bias = d->w.m0_b; a.Y = w.h1; a.ldy = MH; a.M = N; a.K = T2; a.Nout = MH; a.act = 1;
    if ((rc = launch_linear<EPI_PLAIN>(a, st))) return rc;
    lin_defaults(a);
    a.X = w.h1; a.ldx = MH; a.W = d->w.m2_w; a.ldw = MH; a.bias = d->w.m2_b; a.Y = w.h2; a.ldy = MH; a.M = N; a.K = MH; a.Nout = MH; a.act = 1;
    if ((rc = launch_linear<EPI_PLAIN>(a, st))) return rc;
    lin_defaults(a);
    a.X = w.h2; a.ldx = MH; a.W = d->w.m4_w; a.ldw = MH; a.bias = d->w.m4_b; a.Y = w.g; a.ldy = T2; a.M = N; a.K = MH; a.Nout = T2;
    if ((rc = launch_linear<EPI_PLAIN>(a, st))) return rc;
    const long long tot = (long long)N * (PSTL_XIN_LD / 4);
    k_pack_xin<<<pstl_ceil_div(tot, 256), 256, 0, st>>>(nullptr, 0, hl, stlp, w.xin, N, T2, 0, 0ull, 0ull, nullptr, 0);
    PSTL_LAUNCH_CHECK();
    const long long gtot = (long long)(N / per) * T2;
    k_group_fuse<<<pstl_ceil_div(gtot, 256), 256, 0, st>>>(w.g, u0, w.xin, N, T2, n_randoms, per);
    PSTL_LAUNCH_CHECK();
  }
  return PSTL_OK;
}

extern "C" int pstl_refine(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                           const float* hl, const float* stlp, const float* u0, const float* scores, int N,
                           int n_randoms, int n_shards, float w_max, float a_max, int clip_rect, float* out,
                           void* workspace, pstl_stream_t stream) {
  PSTL_CHECK_ARG(d && scene_feat && hl && stlp && u0 && scores && out && workspace, "null argument");
  PSTL_CHECK_ARG(d->w.r0_w && d->w.m0_w && d->r1p, "handle was created without rect_net / merge_net weights");
  PSTL_CHECK_ARG(n_shards > 0 && n_randoms % n_shards == 0 && N % (3 * n_randoms) == 0,
                 "rows must be (scene, n_randoms, 3 modes) with n_shards | n_randoms");
  if (N <= 0) return PSTL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DenoiserWs w;
  size_t total;
  carve(d, N, n_scenes, 1, nullptr, workspace, &w, &total);
  const int H = d->w.rect_hidden, T2 = d->T2;
  int rc = refine_inputs(d, w, scene_feat, n_scenes, hl, stlp, u0, N, n_randoms, n_shards, st);
  if (rc) return rc;
  LinArgs a;
  if ((((d->precision == PSTL_PRECISION_BF16 || d->precision == PSTL_PRECISION_F16) && (128 + rows_per_scene - 1) / rows_per_scene + 1 <= 8) ||
       (d->precision == PSTL_PRECISION_F16X3 && pstl_tc_fits(d, rows_per_scene))) && pstl_tc_has_refine(d))
    return pstl_tc_refine(d, w.cscene, rows_per_scene, w.xin, N, u0, scores, w_max, a_max, clip_rect, out, st);
  rc = mlp_hidden(d, w, N, rows_per_scene, d->r1p, H, nullptr, d->w.r2_w, d->w.r2_b, st);
  if (rc) return rc;
  lin_defaults(a);
  a.X = w.h2; a.ldx = H; a.W = d->w.r4_w; a.ldw = H; a.bias = d->w.r4_b; a.Y = out; a.ldy = T2; a.M = N; a.K = H; a.Nout = T2;
  a.u0 = u0; a.scores = scores; a.w_max = w_max; a.a_max = a_max; a.clip = clip_rect;
  return launch_linear<EPI_REFINE>(a, st);
}

// --------------------------------------------------------------------------------------
// RefineNet backward (the --rect_head training step; upstream: autograd over Net.rect_forward, nusc_model.py:182-235,
// with Adam over net.rect_net.parameters() only, nusc_train.py:1228-1233): given d loss / d rect_controls, the
// gradients of rect_net's three weight matrices and biases.  Activations are recomputed (stateless call); the weight
// gradients dW = dY^T . X are reductions over the N rows, computed as split-K tiles with a fixed-order second pass
// (deterministic); the 224 scene-feature columns of rect_net.0 reduce per scene first (dh1 summed over a scene's
// rows, then a (n_scenes)-deep product) — the same hoisting the forward uses.
// --------------------------------------------------------------------------------------

// d out / d y of the interval head (nusc_model.py:212-232): y -> tanh -> rescale to the headroom -> [score<0] -> clip
__global__ void __launch_bounds__(256) k_refine_dy(float* __restrict__ y, const float* __restrict__ u0,
                                                   const float* __restrict__ scores, const float* __restrict__ d_out,
                                                   long long total, int T2, float w_max, float a_max, int clip) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long long r = e / T2;
  const int c = (int)(e - r * T2);
  float g = 0.f;
  if (scores[r] < 0.f) {
    const float lim = (c & 1) ? a_max : w_max, init = u0[e];
    const float rr = tanhf(y[e]);
    const float dm = (rr >= 0.f) ? (lim - init) : (init - (-lim));
    const float o = init + rr * dm;
    const bool outside = clip && (o < -lim || o > lim);
    if (!outside) g = d_out[e] * dm * (1.f - rr * rr);
  }
  y[e] = g;
}

__global__ void __launch_bounds__(256) k_relu_mask(float* __restrict__ g, const float* __restrict__ h, long long total4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  float4 gv = reinterpret_cast<float4*>(g)[i];
  const float4 hv = reinterpret_cast<const float4*>(h)[i];
  gv.x = hv.x > 0.f ? gv.x : 0.f; gv.y = hv.y > 0.f ? gv.y : 0.f;
  gv.z = hv.z > 0.f ? gv.z : 0.f; gv.w = hv.w > 0.f ? gv.w : 0.f;
  reinterpret_cast<float4*>(g)[i] = gv;
}

__global__ void __launch_bounds__(256) k_transpose(const float* __restrict__ W, int rows, int cols, float* __restrict__ Wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i - r * cols;
  Wt[(size_t)c * rows + r] = W[i];
}

// partial C[z] (Mo x No) = A[r0:r1]^T . B[r0:r1]; A (R, lda >= Mo), B (R, ldb >= No); 64 x 64 tile, 4 x 4 per thread
#define WG_T 64
#define WG_K 16
__global__ void __launch_bounds__(256) k_wgrad(const float* __restrict__ A, int lda, int Mo, const float* __restrict__ B,
                                               int ldb, int No, long long R, int rows_per_split, float* __restrict__ part) {
  __shared__ __align__(16) float As[WG_K][WG_T + 4], Bs[WG_K][WG_T + 4];
  const int tm = blockIdx.x * WG_T, tn = blockIdx.y * WG_T, tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.z * rows_per_split;
  const long long r1 = (r0 + rows_per_split < R) ? r0 + rows_per_split : R;
  const int tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 4, lc = (tid & 15) * 4;  // this thread's slab element: row lr, columns lc..lc+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  auto fetch = [&](const float* __restrict__ P, int ld, int ncols, int c0, long long row) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < r1) {
      const float* q = P + row * ld + c0;
      if (c0 + 3 < ncols && (ld & 3) == 0) v = *reinterpret_cast<const float4*>(q);
      else {
        if (c0 < ncols) v.x = q[0];
        if (c0 + 1 < ncols) v.y = q[1];
        if (c0 + 2 < ncols) v.z = q[2];
        if (c0 + 3 < ncols) v.w = q[3];
      }
    }
    return v;
  };
  float4 na = fetch(A, lda, Mo, tm + lc, r0 + lr), nb = fetch(B, ldb, No, tn + lc, r0 + lr);
  for (long long r = r0; r < r1; r += WG_K) {
    *reinterpret_cast<float4*>(&As[lr][lc]) = na;
    *reinterpret_cast<float4*>(&Bs[lr][lc]) = nb;
    __syncthreads();
    if (r + WG_K < r1) {  // next slab in flight while this one is multiplied
      na = fetch(A, lda, Mo, tm + lc, r + WG_K + lr);
      nb = fetch(B, ldb, No, tn + lc, r + WG_K + lr);
    }
#pragma unroll
    for (int k = 0; k < WG_K; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = part + (size_t)blockIdx.z * Mo * No;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = tm + ty * 4 + i;
    if (m >= Mo) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = tn + tx * 4 + j;
      if (n < No) out[(size_t)m * No + n] = acc[i][j];
    }
  }
}

// column sums of A (R, lda) over a row range: part[z][m]
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ A, int lda, int Mo, long long R,
                                                int rows_per_split, float* __restrict__ part) {
  const int m = threadIdx.x;
  if (m >= Mo) return;
  const long long r0 = (long long)blockIdx.x * rows_per_split;
  const long long r1 = (r0 + rows_per_split < R) ? r0 + rows_per_split : R;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += A[r * lda + m];
  part[(size_t)blockIdx.x * Mo + m] = s;
}

// out[m * ldo + map(n)] = sum over the splits in order; xin_map: packed-row column order -> rect_net.0 column order
__global__ void __launch_bounds__(256) k_split_reduce(const float* __restrict__ part, int splits, int Mo, int No,
                                                      float* __restrict__ out, int ldo, int col0, int xin_T2, int xin_kin,
                                                      int col_tail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mo * No) return;
  const int m = i / No, n = i - m * No;
  int col = col0 + n;
  if (xin_T2 > 0) {  // packed row = [x or fused (T2) | hl | stlp6 | pad]: the two blocks sit apart in the weight matrix
    if (n >= xin_kin) return;
    col = n < xin_T2 ? col0 + n : col_tail + (n - xin_T2);
  }
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(size_t)z * Mo * No + i];
  out[(size_t)m * ldo + col] = s;
}

// sum of a scene's rows: out[scene][h]
__global__ void __launch_bounds__(256) k_scene_sum(const float* __restrict__ g, int H, int rows_per_scene, float* __restrict__ out) {
  const int h = threadIdx.x;
  if (h >= H) return;
  const float* p = g + (size_t)blockIdx.x * rows_per_scene * H + h;
  float s = 0.f;
  for (int r = 0; r < rows_per_scene; ++r) s += p[(size_t)r * H];
  out[(size_t)blockIdx.x * H + h] = s;
}

namespace {
struct RefineBwdWs {
  DenoiserWs base;
  float *dh, *w4t, *w2t, *w0t, *sdh, *part;
  size_t part_floats;
};

const size_t kWgradPartFloats = (size_t)4 << 20;  // 16 MB of split-K partials

void carve_bwd(pstl_denoiser_t d, int N, int n_scenes, void* ws, RefineBwdWs* o, size_t* total) {
  size_t base_bytes;
  carve(d, N, n_scenes, 1, nullptr, ws, &o->base, &base_bytes);
  const int H = d->w.hidden > d->w.rect_hidden ? d->w.hidden : d->w.rect_hidden;
  float* p = ws ? (float*)((char*)ws + base_bytes) : nullptr;
  size_t off = 0;
  auto take = [&](size_t n) { float* r = p ? p + off : nullptr; off += pstl_align_floats(n); return r; };
  o->dh = take((size_t)N * H);
  o->w4t = take((size_t)H * d->T2);
  o->w2t = take((size_t)H * H);
  o->w0t = take((size_t)H * (d->w.feat_dim + d->T2 + d->w.time_dim + 7));
  o->sdh = take((size_t)n_scenes * H);
  o->part = take(kWgradPartFloats);
  o->part_floats = kWgradPartFloats;
  *total = base_bytes + off * sizeof(float);
}

// C (Mo x No, written at out[m*ldo + col0 + n]) = A^T B over R rows
int wgrad(const float* A, int lda, int Mo, const float* B, int ldb, int No, long long R, float* part, size_t part_floats,
          float* out, int ldo, int col0, int xin_T2, int xin_kin, int col_tail, cudaStream_t st) {
  const int tiles = pstl_ceil_div(Mo, WG_T) * pstl_ceil_div(No, WG_T);
  long long splits = (148 * 4 + tiles - 1) / tiles;
  const long long max_by_rows = (R + 255) / 256, max_by_ws = (long long)(part_floats / ((size_t)Mo * No));
  if (splits > max_by_rows) splits = max_by_rows;
  if (splits > max_by_ws) splits = max_by_ws;
  if (splits < 1) splits = 1;
  int rows_per_split = (int)((R + splits - 1) / splits);
  rows_per_split = (rows_per_split + WG_K - 1) / WG_K * WG_K;
  splits = (R + rows_per_split - 1) / rows_per_split;
  dim3 grid(pstl_ceil_div(Mo, WG_T), pstl_ceil_div(No, WG_T), (unsigned)splits);
  k_wgrad<<<grid, 256, 0, st>>>(A, lda, Mo, B, ldb, No, R, rows_per_split, part);
  PSTL_LAUNCH_CHECK();
  k_split_reduce<<<pstl_ceil_div(Mo * No, 256), 256, 0, st>>>(part, (int)splits, Mo, No, out, ldo, col0, xin_T2, xin_kin, col_tail);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

int colsum(const float* A, int lda, int Mo, long long R, float* part, float* out, cudaStream_t st) {
  long long splits = (R + 255) / 256;
  if (splits > 592) splits = 592;
  const int rows_per_split = (int)((R + splits - 1) / splits);
  splits = (R + rows_per_split - 1) / rows_per_split;
  k_colsum<<<(unsigned)splits, 256, 0, st>>>(A, lda, Mo, R, rows_per_split, part);
  PSTL_LAUNCH_CHECK();
  k_split_reduce<<<pstl_ceil_div(Mo, 256), 256, 0, st>>>(part, (int)splits, 1, Mo, out, Mo, 0, 0, 0, 0);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}
}  // namespace

namespace {
// Backward of a hoisted three-layer ReLU MLP (RefineNet's rect_net, the denoiser's policy_net): dy (N, T2) ->
// the six parameter gradients (+ the gradient of the per-scene feature).  Expects the forward activations in the
// workspace: packed rows w.base.xin, h1, h2 (post-ReLU).  Layer-0 weight columns: [feature F | ...]; the packed-row
// blocks [x (T2) | hl, stlp6] land at columns col_x / col_tail, the optional per-row time embedding at col_t.
struct MlpBwd {
  const float *W0, *W2, *W4;
  int in0, H, col_x, col_tail;
  const float* temb_rows;
  int time_dim, col_t;
  float *g_w0, *g_b0, *g_w2, *g_b2, *g_w4, *g_b4, *d_feat;
};

int mlp_backward(pstl_denoiser_t d, const RefineBwdWs& w, const float* dy, const float* scene_feat, int n_scenes,
                 int rows_per_scene, int N, const MlpBwd& m, cudaStream_t st) {
  const DenoiserWs& b = w.base;
  const int H = m.H, T2 = d->T2, F = d->w.feat_dim;
  int rc;
  LinArgs a;
  k_transpose<<<pstl_ceil_div(T2 * H, 256), 256, 0, st>>>(m.W4, T2, H, w.w4t);
  PSTL_LAUNCH_CHECK();
  k_transpose<<<pstl_ceil_div(H * H, 256), 256, 0, st>>>(m.W2, H, H, w.w2t);
  PSTL_LAUNCH_CHECK();
  // layer 4: dW4 (T2 x H) = dy^T h2, db4; dh2 = (dy W4) * [h2 > 0]
  if ((rc = wgrad(dy, T2, T2, b.h2, H, H, N, w.part, w.part_floats, m.g_w4, H, 0, 0, 0, 0, st))) return rc;
  if ((rc = colsum(dy, T2, T2, N, w.part, m.g_b4, st))) return rc;
  lin_defaults(a);
  a.X = dy; a.ldx = T2; a.W = w.w4t; a.ldw = T2; a.Y = w.dh; a.ldy = H; a.M = N; a.K = T2; a.Nout = H;
  if ((rc = launch_linear<EPI_PLAIN>(a, st))) return rc;
  const long long tot4 = (long long)N * H / 4;
  k_relu_mask<<<(unsigned)pstl_ceil_div(tot4, 256), 256, 0, st>>>(w.dh, b.h2, tot4);
  PSTL_LAUNCH_CHECK();
  // layer 2: dW2 (H x H) = dh2^T h1, db2; dh1 = (dh2 W2) * [h1 > 0] (into h2's storage, which is dead now)
  if ((rc = wgrad(w.dh, H, H, b.h1, H, H, N, w.part, w.part_floats, m.g_w2, H, 0, 0, 0, 0, st))) return rc;
  if ((rc = colsum(w.dh, H, H, N, w.part, m.g_b2, st))) return rc;
  lin_defaults(a);
  a.X = w.dh; a.ldx = H; a.W = w.w2t; a.ldw = H; a.Y = b.h2; a.ldy = H; a.M = N; a.K = H; a.Nout = H;
  if ((rc = launch_linear<EPI_PLAIN>(a, st))) return rc;
  k_relu_mask<<<(unsigned)pstl_ceil_div(tot4, 256), 256, 0, st>>>(b.h2, b.h1, tot4);
  PSTL_LAUNCH_CHECK();
  // layer 0: the packed-row columns; the time-embedding columns; the F scene-feature columns per scene
  if ((rc = wgrad(b.h2, H, H, b.xin, PSTL_XIN_LD, PSTL_XIN_LD, N, w.part, w.part_floats, m.g_w0, m.in0, m.col_x, T2, d->kin,
                  m.col_tail, st)))
    return rc;
  if (m.temb_rows &&
      (rc = wgrad(b.h2, H, H, m.temb_rows, m.time_dim, m.time_dim, N, w.part, w.part_floats, m.g_w0, m.in0, m.col_t, 0, 0, 0, st)))
    return rc;
  k_scene_sum<<<n_scenes, 256, 0, st>>>(b.h2, H, rows_per_scene, w.sdh);
  PSTL_LAUNCH_CHECK();
  if ((rc = wgrad(w.sdh, H, H, scene_feat, F, F, n_scenes, w.part, w.part_floats, m.g_w0, m.in0, 0, 0, 0, 0, st))) return rc;
  if ((rc = colsum(w.sdh, H, H, n_scenes, w.part, m.g_b0, st))) return rc;
  if (m.d_feat) {  // d loss / d scene feature = (sum of a scene's dh1 rows) . W0[:, :F]
    k_transpose<<<pstl_ceil_div(H * m.in0, 256), 256, 0, st>>>(m.W0, H, m.in0, w.w0t);
    PSTL_LAUNCH_CHECK();
    lin_defaults(a);
    a.X = w.sdh; a.ldx = H; a.W = w.w0t; a.ldw = H; a.Y = m.d_feat; a.ldy = F; a.M = n_scenes; a.K = H; a.Nout = F;
    if ((rc = launch_linear<EPI_PLAIN>(a, st))) return rc;
  }
  return PSTL_OK;
}
}  // namespace

extern "C" size_t pstl_refine_backward_workspace_bytes(pstl_denoiser_t d, int N, int n_scenes) {
  if (!d) return 0;
  RefineBwdWs w;
  size_t total;
  carve_bwd(d, N, n_scenes, nullptr, &w, &total);
  return total;
}

extern "C" int pstl_refine_backward(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                                    const float* hl, const float* stlp, const float* u0, const float* scores, int N,
                                    int n_randoms, int n_shards, float w_max, float a_max, int clip_rect,
                                    const float* d_out, float* g_r0_w, float* g_r0_b, float* g_r2_w, float* g_r2_b,
                                    float* g_r4_w, float* g_r4_b, int reuse_activations, void* workspace,
                                    pstl_stream_t stream) {
  PSTL_CHECK_ARG(d && scene_feat && hl && stlp && u0 && scores && d_out && workspace, "null argument");
  PSTL_CHECK_ARG(g_r0_w && g_r0_b && g_r2_w && g_r2_b && g_r4_w && g_r4_b, "null gradient output");
  PSTL_CHECK_ARG(d->w.r0_w && d->w.m0_w && d->r1p, "handle was created without rect_net / merge_net weights");
  PSTL_CHECK_ARG(n_shards > 0 && n_randoms % n_shards == 0 && N % (3 * n_randoms) == 0,
                 "rows must be (scene, n_randoms, 3 modes) with n_shards | n_randoms");
  PSTL_CHECK_ARG(N > 0 && (long long)n_scenes * rows_per_scene == N, "rows must be n_scenes * rows_per_scene");
  cudaStream_t st = (cudaStream_t)stream;
  RefineBwdWs w;
  size_t total;
  carve_bwd(d, N, n_scenes, workspace, &w, &total);
  const DenoiserWs& b = w.base;
  const int H = d->w.rect_hidden, T2 = d->T2, F = d->w.feat_dim, inr = F + 7 + T2;
  PSTL_CHECK_ARG(H <= 256 && H % 4 == 0, "rect_net hidden width must be a multiple of 4, at most 256");
  // forward, fp32, keeping the activations: xin, h1, h2 (post-ReLU) and y (pre-tanh); with reuse_activations the
  // first three are still in the workspace from the fp32 pstl_refine call and only y is rebuilt
  int rc = PSTL_OK;
  if (!reuse_activations) {
    if ((rc = refine_inputs(d, b, scene_feat, n_scenes, hl, stlp, u0, N, n_randoms, n_shards, st))) return rc;
    if ((rc = mlp_hidden(d, b, N, rows_per_scene, d->r1p, H, nullptr, d->w.r2_w, d->w.r2_b, st))) return rc;
  }
  LinArgs a;
  lin_defaults(a);
  a.X = b.h2; a.ldx = H; a.W = d->w.r4_w; a.ldw = H; a.bias = d->w.r4_b; a.Y = b.g; a.ldy = T2; a.M = N; a.K = H; a.Nout = T2;
  if ((rc = launch_linear<EPI_PLAIN>(a, st))) return rc;
  const long long tot = (long long)N * T2;
  k_refine_dy<<<(unsigned)pstl_ceil_div(tot, 256), 256, 0, st>>>(b.g, u0, scores, d_out, tot, T2, w_max, a_max, clip_rect);
  PSTL_LAUNCH_CHECK();
  MlpBwd m{};
  m.W0 = d->w.r0_w; m.W2 = d->w.r2_w; m.W4 = d->w.r4_w; m.in0 = inr; m.H = H;
  m.col_x = F + 7; m.col_tail = F;
  m.g_w0 = g_r0_w; m.g_b0 = g_r0_b; m.g_w2 = g_r2_w; m.g_b2 = g_r2_b; m.g_w4 = g_r4_w; m.g_b4 = g_r4_b;
  return mlp_backward(d, w, b.g, scene_feat, n_scenes, rows_per_scene, N, m, st);
}

// --------------------------------------------------------------------------------------
// Denoiser training step (reference nusc_train.py:539-555 diffusion_prep, :1352-1356 net(...), :432-436 loss_diffusion):
// eps prediction with ONE TIMESTEP PER ROW (the sampler's entry takes one per call), and its backward into policy_net
// and the scene feature (the encoders' own backward stays with autograd).  temb_rows (N, time_dim) is the sinusoidal
// embedding of each row's timestep; its layer-0 term is a K = time_dim product added before the ReLU.
// --------------------------------------------------------------------------------------
static int eps_rows_activations(pstl_denoiser_t d, const DenoiserWs& w, const float* scene_feat, int n_scenes,
                                int rows_per_scene, const float* hl, const float* stlp, const float* x, int N,
                                const float* temb_rows, cudaStream_t st) {
  const int H = d->w.hidden, T2 = d->T2, F = d->w.feat_dim, TD = d->w.time_dim;
  const int in1 = F + T2 + TD + 7;
  int rc = hoist(d->w.p0_w, in1, d->w.p0_b, H, scene_feat, n_scenes, F, w.cscene, nullptr, 0, 0, 0, nullptr, st);
  if (rc) return rc;
  const long long tot = (long long)N * (PSTL_XIN_LD / 4);
  k_pack_xin<<<pstl_ceil_div(tot, 256), 256, 0, st>>>(x, T2, hl, stlp, w.xin, N, T2, 0, 0ull, 0ull, nullptr, 0);
  PSTL_LAUNCH_CHECK();
  LinArgs a;
  lin_defaults(a);  // time term of layer 0 -> h2's storage (free until layer 2 writes it)
  a.X = temb_rows; a.ldx = TD; a.W = d->w.p0_w + F + T2; a.ldw = in1; a.Y = w.h2; a.ldy = H; a.M = N; a.K = TD; a.Nout = H;
  if ((rc = launch_linear<EPI_PLAIN>(a, st))) return rc;
  lin_defaults(a);
  a.X = w.xin; a.ldx = PSTL_XIN_LD; a.W = d->w1p; a.ldw = d->kin; a.rowbias = w.cscene; a.rows_per_group = rows_per_scene;
  a.ldrb = H; a.res = w.h2; a.ldres = H; a.Y = w.h1; a.ldy = H; a.M = N; a.K = d->kin; a.Nout = H; a.act = 1;
  if ((rc = launch_linear<EPI_PLAIN>(a, st))) return rc;
  lin_defaults(a);
  a.X = w.h1; a.ldx = H; a.W = d->w.p2_w; a.ldw = H; a.bias = d->w.p2_b; a.Y = w.h2; a.ldy = H; a.M = N; a.K = H; a.Nout = H; a.act = 1;
  return launch_linear<EPI_PLAIN>(a, st);
}

extern "C" int pstl_denoiser_eps_rows(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                                      const float* hl, const float* stlp, const float* x, int N, const float* temb_rows,
                                      float* eps_out, void* workspace, pstl_stream_t stream) {
  PSTL_CHECK_ARG(d && scene_feat && hl && stlp && x && temb_rows && eps_out && workspace, "null argument");
  PSTL_CHECK_ARG(rows_per_scene >= 1 && (long long)n_scenes * rows_per_scene >= N, "bad scene mapping");
  if (N <= 0) return PSTL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DenoiserWs w;
  size_t total;
  carve(d, N, n_scenes, 1, nullptr, workspace, &w, &total);
  int rc = eps_rows_activations(d, w, scene_feat, n_scenes, rows_per_scene, hl, stlp, x, N, temb_rows, st);
  if (rc) return rc;
  const int H = d->w.hidden, T2 = d->T2;
  LinArgs a;
  lin_defaults(a);  // eps = policy_net(...) + x (nusc_model.py:162)
  a.X = w.h2; a.ldx = H; a.W = d->w.p4_w; a.ldw = H; a.bias = d->w.p4_b; a.res = x; a.ldres = T2; a.Y = eps_out; a.ldy = T2;
  a.M = N; a.K = H; a.Nout = T2;
  return launch_linear<EPI_PLAIN>(a, st);
}

extern "C" int pstl_denoiser_eps_backward(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                                          const float* hl, const float* stlp, const float* x, int N,
                                          const float* temb_rows, const float* d_eps, float* g_p0_w, float* g_p0_b,
                                          float* g_p2_w, float* g_p2_b, float* g_p4_w, float* g_p4_b, float* d_scene_feat,
                                          int reuse_activations, void* workspace, pstl_stream_t stream) {
  PSTL_CHECK_ARG(d && scene_feat && hl && stlp && x && temb_rows && d_eps && workspace, "null argument");
  PSTL_CHECK_ARG(g_p0_w && g_p0_b && g_p2_w && g_p2_b && g_p4_w && g_p4_b, "null gradient output");
  PSTL_CHECK_ARG(N > 0 && (long long)n_scenes * rows_per_scene == N, "rows must be n_scenes * rows_per_scene");
  cudaStream_t st = (cudaStream_t)stream;
  RefineBwdWs w;
  size_t total;
  carve_bwd(d, N, n_scenes, workspace, &w, &total);
  const int H = d->w.hidden, T2 = d->T2, F = d->w.feat_dim, TD = d->w.time_dim;
  PSTL_CHECK_ARG(H <= 256 && H % 4 == 0, "hidden width must be a multiple of 4, at most 256");
  int rc = PSTL_OK;
  if (!reuse_activations &&
      (rc = eps_rows_activations(d, w.base, scene_feat, n_scenes, rows_per_scene, hl, stlp, x, N, temb_rows, st)))
    return rc;
  MlpBwd m{};
  m.W0 = d->w.p0_w; m.W2 = d->w.p2_w; m.W4 = d->w.p4_w; m.in0 = F + T2 + TD + 7; m.H = H;
  m.col_x = F; m.col_tail = F + T2 + TD;
  m.temb_rows = temb_rows; m.time_dim = TD; m.col_t = F + T2;
  m.g_w0 = g_p0_w; m.g_b0 = g_p0_b; m.g_w2 = g_p2_w; m.g_b2 = g_p2_b; m.g_w4 = g_p4_w; m.g_b4 = g_p4_b;
  m.d_feat = d_scene_feat;
  return mlp_backward(d, w, d_eps, scene_feat, n_scenes, rows_per_scene, N, m, st);
}
