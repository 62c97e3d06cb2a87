// mlp_common.cuh — denoiser handle layout shared by the fp32 SIMT path and the tcgen05 path
#pragma once
#include "common.cuh"

#define PSTL_XIN_LD 48  // packed per-chain input row: [x (2T=40) | hl | stlp(6) | pad]

struct pstl_denoiser {
  pstl_weights w;      // caller's device pointers (kept: the caller owns them for the handle's life)
  int precision;
  int T2;              // 2*T
  int kin;             // T2 + 7 : per-step GEMM depth after hoisting
  float* w1p;          // (hidden, kin)   policy_net.0 columns [x | hl | stlp]
  float* r1p;          // (rect_hidden, kin) rect_net.0 columns [fused | hl | stlp]
  void* tc;            // tcgen05 engine state (bf16 images), owned by denoiser_tc.cu
  int tc_engine;       // 0 = automatic, 1 = one-SM engine, 2 = CTA-pair engine (pstl_denoiser_set_engine)
  const unsigned long long* offset_dev;  // optional device word added to the Philox offset (CUDA-graph replays)
};

// workspace carve-up (floats) for N chains
struct DenoiserWs {
  float* xin;     // (N, 48)
  float* h1;      // (N, H)
  float* h2;      // (N, H)
  float* g;       // (N, T2)   merge output / mu buffer
  float* cscene;  // (n_scenes, H)
  float* ct;      // (steps, H)
  float* adam;    // (3, N, T2) m, v, anchor
  float* gws;     // guidance gradient + tape
};

static inline size_t pstl_align_floats(size_t n) { return (n + 63) & ~(size_t)63; }

// Philox4x32-10 (Salmon et al. 2011), counter-based normals for the throughput mode
__device__ __forceinline__ uint4 pstl_philox(uint4 ctr, uint2 key) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ void pstl_box_muller(unsigned a, unsigned b, float& z0, float& z1) {
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0,1]: never denormal, never 0
  const float u2 = (float)b * 2.3283064365386963e-10f;
  // r = sqrt(-2 ln u1) through the SFU (lg2 / sqrt approximations, ~2 ulp): noise, not arithmetic that has a reference
  float l2, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// z for (row, col) at reverse step `step`: one Philox call yields 4 normals for cols 4q..4q+3
__device__ __forceinline__ float pstl_noise_at(uint64_t seed, uint64_t offset, int step, long long row, int col) {
  uint4 ctr = make_uint4((unsigned)(row & 0xffffffff), (unsigned)(row >> 32), (unsigned)(col >> 2),
                         (unsigned)step + (unsigned)offset);
  uint2 key = make_uint2((unsigned)(seed & 0xffffffff), (unsigned)(seed >> 32));
  const uint4 r = pstl_philox(ctr, key);
  float z0, z1, z2, z3;
  pstl_box_muller(r.x, r.y, z0, z1);
  pstl_box_muller(r.z, r.w, z2, z3);
  const int k = col & 3;
  return k == 0 ? z0 : (k == 1 ? z1 : (k == 2 ? z2 : z3));
}
