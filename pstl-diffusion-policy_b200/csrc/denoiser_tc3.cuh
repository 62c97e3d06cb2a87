// denoiser_tc3.cuh — the fp32-GRADE tcgen05 engine of the DDPM reverse loop (PSTL_PRECISION_F16X3): every operand is
// carried as TWO fp16 pieces (x = x_hi + x_lo: 2 x 11 = 22 mantissa bits) and every product as three tensor-core MMAs
// (hi.hi + lo.hi + hi.lo; the lo.lo term is below 2^-22 of the product) with fp32 accumulation in TMEM.  Emulated on
// the CPU over the whole 99-step chain the final controls deviate from an fp64 run by 6e-7 of
// the control range — what fp32 arithmetic itself deviates by (5.4e-7; bf16 pieces: 2.4e-6; plain bf16 operands: 1e-3) — so
// this engine replaces the fp32 CUDA-core sampler (156 ms per 196,608 chains) at tensor-core speed.
// Range: a piece is an fp16 number, so |activation|, |weight| must stay below 65,504 (conversions saturate instead of
// producing inf); pieces below 6e-5 are fp16 subnormals (absolute error <= 3e-8, irrelevant next to O(1) sums).
// Included by denoiser_tc.cu after denoiser_tc2.cuh (cluster / pair helpers).
//
// Layout: the CTA pair of denoiser_tc2.cuh (tcgen05.mma.cta_group::2, each CTA holds half of N of every weight matrix,
// now as a hi image and a lo image: 2 x 92 KB of shared memory) with the TS operand form of k_denoiser_tc (activations
// never leave TMEM).  One 256-row tile per pair; per CTA
//   TMEM  D [0,256) fp32 | H_hi [256,384) | H_lo [384,512)   (fp16 pairs; H2 overwrites H1 in place once layer 2 has
//         retired); the layer-1 operand X = [x 40 | hl stlp 0 | one-hot class pairs] aliases H_hi/H_lo[0,32) (dead once
//         layer 1 has retired), layer 3's 48 columns alias D[0,48).
//   Biases: layer 1 through the one-hot K-step against the per-step bias columns (W1'_hi columns 48..63, (hi, lo) pairs
//         rewritten by the bias warp); layer 2 through ONE shared-memory (SS) MMA of a constant ones tile against a
//         (hi, mid, lo) image of b2 that initialises the accumulator; b3 in the epilogue.
// With H, D and X all in TMEM there is no room for N-half pipelining: the layers run back to back (L1 -> E1 -> L2 ->
// E2 -> L3 -> E3), layer 3 starting on the first half of H2; the noise of the step is drawn under layer 2.
#pragma once
#include <cuda_fp16.h>

namespace {

constexpr int k3OffWhi = 0;
constexpr int k3OffWlo = k2WeightBytes;
constexpr int k3OffB2 = 2 * k2WeightBytes;            // [128 n x 16 k] fp16, K-major no-swizzle: k 0..2 = (hi, mid, lo) of b2
constexpr int k3OffOnes = k3OffB2 + 128 * 16 * 2;      // [128 m x 16 k]: k 0..2 = 1
constexpr int k3OffB3 = k3OffOnes + 128 * 16 * 2;
constexpr int k3OffBar = k3OffB3 + 64 * 4;
constexpr int k3OffIt = k3OffBar + 256;                // [128 x 40] fp32 kept-iterate staging
constexpr int k3SmemBytes = k3OffIt + 128 * 40 * 4;
static_assert(k3SmemBytes + 1024 <= 227 * 1024, "shared memory budget");
constexpr uint32_t k3ColD = 0, k3ColHhi = 256, k3ColHlo = 384;
enum { k3BW = 0, k3BX, k3BH1, k3BH2a, k3BH2b, k3BD1, k3BD2, k3BD3, k3BB };

__device__ __forceinline__ void mma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// kind::f16 instruction descriptor with fp16 A and B (format code 0), D = f32, both K-major
__device__ __forceinline__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// (hi, lo) fp16 images of a pair of fp32 values: hi = rn(v) (saturating), lo = rn(v - hi); packed {second:16 | first:16}
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float r0 = v0 - f.x, r1 = v1 - f.y;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ __half half_sat(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }
// (hi, lo) of one value packed {lo:16 | hi:16}: hi at the lower address
__device__ __forceinline__ uint32_t split_f16(float v) {
  const __half h = half_sat(v);
  const __half l = half_sat(v - __half2float(h));
  return (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
}

// pair images: rank r at r * 2 * k2WeightBytes: [hi image | lo image], each in the layout of k_build_image2
__global__ void k_build_image3(const float* __restrict__ w1p, int kin, const float* __restrict__ w2,
                               const float* __restrict__ w3, int n3, uint8_t* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  auto put = [&](int rank, int off, float v) {
    const __half h = half_sat(v);
    const __half l = half_sat(v - __half2float(h));
    uint8_t* base = img + (size_t)rank * 2 * k2WeightBytes + off;
    *reinterpret_cast<__half*>(base) = h;
    *reinterpret_cast<__half*>(base + k2WeightBytes) = l;
  };
  if (i < 256 * 64) {
    const int n = i / 64, k = i % 64;
    put(n / 128, k2OffW1 + sw128_off(n % 128, k), k < kin ? w1p[n * kin + k] : 0.f);
  }
  if (i < 256 * 256) {
    const int n = i / 256, k = i % 256;
    put(n / 128, k2OffW2 + (k / 64) * (128 * 128) + sw128_off(n % 128, k % 64), w2[n * 256 + k]);
  }
  if (i < kN3 * 256) {
    const int n = i / 256, k = i % 256;
    put(n / 24, k2OffW3 + (k / 64) * (24 * 128) + sw128_off(n % 24, k % 64), n < n3 ? w3[n * 256 + k] : 0.f);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_denoiser_tc3(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - sraw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const uint32_t bar0 = sbase + k3OffBar;
  auto bar = [&](int which) { return bar0 + 8u * which; };
  const uint32_t lead0 = mapa_rank(bar0, 0);
  auto lbar = [&](int which) { return lead0 + 8u * which; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + k3OffBar + 8 * 24);
  float* b3s = reinterpret_cast<float*>(smem + k3OffB3);

  if (warp == kEpiWarps) {
    if (lane == 0) {
      mbar_init(bar(k3BW), 1);
      mbar_init(bar(k3BX), 2 * kEpiWarps);
      mbar_init(bar(k3BH1), 2 * kEpiWarps);
      mbar_init(bar(k3BH2a), 2 * kEpiWarps);
      mbar_init(bar(k3BH2b), 2 * kEpiWarps);
      mbar_init(bar(k3BD1), 1);
      mbar_init(bar(k3BD2), 1);
      mbar_init(bar(k3BD3), 1);
      mbar_init(bar(k3BB), 2);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // b2 image of this CTA's half of N ((hi, mid, lo) in k = 0..2) and the ones tile it is multiplied with
  for (int n = threadIdx.x; n < 128; n += kThreads) {
    const float v = a.b2[rank * 128 + n];
    const __half h = half_sat(v);
    const float r1 = v - __half2float(h);
    const __half m = half_sat(r1);
    const __half l = half_sat(r1 - __half2float(m));
    const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      *reinterpret_cast<__half*>(smem + k3OffB2 + flat16_off(n, k)) = k == 0 ? h : (k == 1 ? m : (k == 2 ? l : zero));
      *reinterpret_cast<__half*>(smem + k3OffOnes + flat16_off(n, k)) = k < 3 ? one : zero;
    }
  }
  for (int i = threadIdx.x; i < 64; i += kThreads) b3s[i] = i < 40 ? a.b3[i] : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem = 0u;

  if (warp == kEpiWarps && lane == 0) {
    mbar_expect_tx(bar(k3BW), 2 * k2WeightBytes);
    constexpr int kChunk = 23552;  // 8 equal bulk copies: hi image, then lo image
    const uint8_t* src = a.image + (size_t)rank * 2 * k2WeightBytes;
    for (int off = 0; off < 2 * k2WeightBytes; off += kChunk) bulk_g2s(sbase + off, src + off, kChunk, bar(k3BW));
  }

  const int n_tiles = (a.N + k2TileM - 1) / k2TileM;
  const int S = a.first_step - a.last_step + 1;

  if (warp == kEpiWarps + 1) {
    // ================= bias warp (both CTAs): this CTA's half of the layer-1 bias K-step =================
    mbar_wait(bar(k3BW), 0);
    uint32_t it = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const long long r0 = (long long)tile * k2TileM;
      const long long r1 = (r0 + k2TileM - 1 < a.N) ? r0 + k2TileM - 1 : (long long)a.N - 1;
      const int scene0 = (int)(r0 / a.rows_per_scene);
      const int n_cls = (int)(r1 / a.rows_per_scene) - scene0 + 1;
      for (int s = 0; s < S; ++s, ++it) {
        if (it > 0) mbar_wait(bar(k3BD1), (it - 1) & 1);  // the previous layer-1 MMAs have retired
        const float* ctr = a.ct + (size_t)(a.first_step - s) * kH + rank * 128;
        for (int c = 0; c < n_cls && c < kMaxClasses; ++c) {
          const float* src = a.cscene + (size_t)(scene0 + c) * kH + rank * 128;
#pragma unroll
          for (int n = lane; n < 128; n += 32)
            *reinterpret_cast<uint32_t*>(smem + k3OffWhi + k2OffW1 + sw128_off(n, 48 + 2 * c)) = split_f16(__ldg(src + n) + __ldg(ctr + n));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lbar(k3BB));
      }
    }
  } else if (warp == kEpiWarps) {
    if (rank == 0) {
      // ================= MMA issuer (leader CTA) =================
      mbar_wait(bar(k3BW), 0);
      constexpr uint32_t idN = make_idesc_f16(k2TileM, kH), id3 = make_idesc_f16(k2TileM, kN3);
      const uint64_t dW1h = make_desc(sbase + k3OffWhi + k2OffW1), dW1l = make_desc(sbase + k3OffWlo + k2OffW1);
      const uint64_t dB2 = make_desc_flat(sbase + k3OffB2, 128, 256), dOnes = make_desc_flat(sbase + k3OffOnes, 128, 256);
      uint32_t it = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        for (int s = 0; s < S; ++s, ++it) {
          const uint32_t ph = it & 1u;
          mbar_spin_cluster(bar(k3BB), ph);
          mbar_spin_cluster(bar(k3BX), ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // x | hl stlp: hi.hi + lo.hi + hi.lo
              mma2_ts(tmem + k3ColD, tmem + k3ColHhi + k * 8, dW1h + (uint64_t)(2 * k), idN, k > 0);
              mma2_ts(tmem + k3ColD, tmem + k3ColHlo + k * 8, dW1h + (uint64_t)(2 * k), idN, 1);
              mma2_ts(tmem + k3ColD, tmem + k3ColHhi + k * 8, dW1l + (uint64_t)(2 * k), idN, 1);
            }
            mma2_ts(tmem + k3ColD, tmem + k3ColHhi + 24, dW1h + 6, idN, 1);  // one-hot class x (hi, lo) bias columns
            tc_commit2(bar(k3BD1));
          }
          __syncwarp();
          mbar_spin_cluster(bar(k3BH1), ph);
          tc_fence_after();
          if (elect_one()) {
            mma2_ss(tmem + k3ColD, dOnes, dB2, idN, 0);  // D = b2 (three fp16 pieces): the only shared-memory A operand
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const uint64_t o = (uint64_t)((k & 3) * 2);
              const uint64_t dBh = make_desc(sbase + k3OffWhi + k2OffW2 + (k >> 2) * (128 * 128)) + o;
              const uint64_t dBl = make_desc(sbase + k3OffWlo + k2OffW2 + (k >> 2) * (128 * 128)) + o;
              mma2_ts(tmem + k3ColD, tmem + k3ColHhi + k * 8, dBh, idN, 1);
              mma2_ts(tmem + k3ColD, tmem + k3ColHlo + k * 8, dBh, idN, 1);
              mma2_ts(tmem + k3ColD, tmem + k3ColHhi + k * 8, dBl, idN, 1);
            }
            tc_commit2(bar(k3BD2));
          }
          __syncwarp();
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            mbar_spin_cluster(bar(part ? k3BH2b : k3BH2a), ph);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int k = part * 8; k < part * 8 + 8; ++k) {
                const uint64_t o = (uint64_t)((k & 3) * 2);
                const uint64_t dBh = make_desc(sbase + k3OffWhi + k2OffW3 + (k >> 2) * (24 * 128)) + o;
                const uint64_t dBl = make_desc(sbase + k3OffWlo + k2OffW3 + (k >> 2) * (24 * 128)) + o;
                mma2_ts(tmem + k3ColD, tmem + k3ColHhi + k * 8, dBh, id3, k > 0);
                mma2_ts(tmem + k3ColD, tmem + k3ColHlo + k * 8, dBh, id3, 1);
                mma2_ts(tmem + k3ColD, tmem + k3ColHhi + k * 8, dBl, id3, 1);
              }
              if (part) tc_commit2(bar(k3BD3));
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ================= epilogue warps (both CTAs) =================
    const int q = warp & 3, ch = warp >> 2;
    const int m = q * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    const int c0 = ch * 20;
    const unsigned off_base = (unsigned)a.offset + (a.offset_dev ? (unsigned)__ldg(a.offset_dev) : 0u);
    const uint2 key = make_uint2((unsigned)(a.seed & 0xffffffff), (unsigned)(a.seed >> 32));
    auto hand_over = [&](uint32_t leader_bar) {
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_bar);
    };
    uint32_t it = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const int row = tile * k2TileM + (int)rank * 128 + m;
      const bool live = row < a.N;
      const int rrow = live ? row : a.N - 1;
      const int scene0 = (tile * k2TileM) / a.rows_per_scene;
      const int cls = rrow / a.rows_per_scene - scene0;
      float x[20], pre[20];
      const float* xr = a.xin + (size_t)rrow * PSTL_XIN_LD;
#pragma unroll
      for (int j = 0; j < 20; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c0 + j);
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
      }
      // constant columns 40..47 (hl, stlp(6), 0) as (hi, lo) images; 48..63: ones at this row's scene class (hi only)
      uint32_t pch[4], pcl[4];
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + 40 + j);
        split_pair(v.x, v.y, pch[j / 2], pcl[j / 2]);
        split_pair(v.z, v.w, pch[j / 2 + 1], pcl[j / 2 + 1]);
      }
      auto store_x = [&]() {
        uint32_t ph_[10], pl_[10];
#pragma unroll
        for (int j = 0; j < 20; j += 2) split_pair(x[j], x[j + 1], ph_[j / 2], pl_[j / 2]);
        TMEM_ST_X8(tmem + lane_addr + k3ColHhi + c0 / 2, ph_);
        TMEM_ST_X2(tmem + lane_addr + k3ColHhi + c0 / 2 + 8, (ph_ + 8));
        TMEM_ST_X8(tmem + lane_addr + k3ColHlo + c0 / 2, pl_);
        TMEM_ST_X2(tmem + lane_addr + k3ColHlo + c0 / 2 + 8, (pl_ + 8));
        if (ch == 1) {
          uint32_t oh[8], zz[8];
#pragma unroll
          for (int c = 0; c < kMaxClasses; ++c) { oh[c] = (c == cls) ? 0x3C003C00u : 0u; zz[c] = 0u; }
          TMEM_ST_X4(tmem + lane_addr + k3ColHhi + 20, pch);
          TMEM_ST_X8(tmem + lane_addr + k3ColHhi + 24, oh);
          TMEM_ST_X4(tmem + lane_addr + k3ColHlo + 20, pcl);
          TMEM_ST_X8(tmem + lane_addr + k3ColHlo + 24, zz);
        }
      };
      store_x();
      hand_over(lbar(k3BX));

      for (int s = 0; s < S; ++s, ++it) {
        const uint32_t ph = it & 1u;
        const int i = a.first_step - s;
        // ---- layers 1 and 2: D -> relu -> (hi, lo) fp16 images -> H_hi / H_lo ----
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          mbar_spin(bar(layer == 0 ? k3BD1 : k3BD2), ph);
          tc_fence_after();
          // this thread: columns [128 p + 64 ch, +64) in part p (fp16 pairs: TMEM columns [64 p + 32 ch, +32) of H_hi / H_lo)
          uint32_t ra[16], rb[16];
          const uint32_t dsrc = tmem + lane_addr + k3ColD + ch * 64;
          const uint32_t hdst = tmem + lane_addr + ch * 32;
          auto emit = [&](const uint32_t(&r)[16], int cidx) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2)
              split_pair(fmaxf(__uint_as_float(r[j]), 0.f), fmaxf(__uint_as_float(r[j + 1]), 0.f), h[j / 2], l[j / 2]);
            const uint32_t col = hdst + (cidx >> 2) * 64 + (cidx & 3) * 8;
            TMEM_ST_X8(col + k3ColHhi, h);
            TMEM_ST_X8(col + k3ColHlo, l);
          };
          auto src_col = [&](int cidx) { return dsrc + (cidx >> 2) * 128 + (cidx & 3) * 16; };
          TMEM_LD_X16(src_col(0), ra);
#pragma unroll
          for (int cidx = 0; cidx < 8; cidx += 2) {
            tmem_wait_ld();
            TMEM_LD_X16(src_col(cidx + 1), rb);
            emit(ra, cidx);
            tmem_wait_ld();
            if (cidx + 2 < 8) TMEM_LD_X16(src_col(cidx + 2), ra);
            emit(rb, cidx + 1);
            // layer 2: the first 128 columns of H2 are complete: layer 3's first eight K-steps may start (they write
            // D[0,48), which is consumed; the load in flight reads columns >= 128)
            if (layer == 1 && cidx == 2) hand_over(lbar(k3BH2a));
          }
          hand_over(lbar(layer == 0 ? k3BH1 : k3BH2b));
          if (layer == 0) {
            // ---- this step's noise and the pre-folded part of the update, in the shadow of layer 2 ----
            const bool draw = i > 1 && !a.refine && !a.mu_out;
#pragma unroll
            for (int j = 0; j < 20; ++j) pre[j] = 0.f;
            if (draw) {
              if (a.noise) {
                const float* zr = a.noise + ((size_t)(a.steps - 1 - i) * a.N + (size_t)rrow) * 40 + c0;
#pragma unroll
                for (int j = 0; j < 20; j += 4) {
                  const float4 zz = *reinterpret_cast<const float4*>(zr + j);
                  pre[j] = zz.x; pre[j + 1] = zz.y; pre[j + 2] = zz.z; pre[j + 3] = zz.w;
                }
              } else {
                unsigned step_ctr = (unsigned)i + off_base;
                asm volatile("" : "+r"(step_ctr)::"memory");
                uint4 rn[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) rn[k] = pstl_philox(make_uint4((unsigned)rrow, 0u, (unsigned)(c0 / 4 + k), step_ctr), key);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                  pstl_box_muller(rn[k].x, rn[k].y, pre[4 * k], pre[4 * k + 1]);
                  pstl_box_muller(rn[k].z, rn[k].w, pre[4 * k + 2], pre[4 * k + 3]);
                }
              }
            }
          }
        }
        // ---- layer 3: eps, posterior mean, noise, next x — the same fp32 expressions as the fp32 path's epilogue
        //      (mlp_fp32.cu EPI_DDPM): mu = c2 (x - c1 eps), x' = mu + sb z ----
        const float c1 = a.c1[i], c2 = a.c2[i], sb = a.sb[i];
        mbar_spin(bar(k3BD3), ph);
        tc_fence_after();
        uint32_t r[20];
        TMEM_LD_X16(tmem + lane_addr + k3ColD + c0, r);
        TMEM_LD_X4(tmem + lane_addr + k3ColD + c0 + 16, (r + 16));
        tmem_wait_ld();
        if (a.refine) {
          const float viol = (a.scores[rrow] < 0.f) ? 1.f : 0.f;
          const float* u0r = a.u0 + (size_t)rrow * 40 + c0;
          float* orow = a.out + (size_t)rrow * 40 + c0;
#pragma unroll
          for (int j = 0; j < 20; ++j) {
            const float rr = tanhf(__uint_as_float(r[j]) + b3s[c0 + j]);
            const float init = u0r[j];
            const float lim = (j & 1) ? a.a_max : a.w_max;
            const float mk = (rr >= 0.f) ? 1.f : 0.f;
            const float merged = (rr * (init - (-lim))) * (1.f - mk) + (rr * (lim - init)) * mk;
            float o = init + merged * viol;
            if (a.clip) o = fminf(fmaxf(o, -lim), lim);
            if (live) orow[j] = o;
          }
          continue;
        }
        if (a.mu_out) {
          float* mo = a.mu_out + (size_t)rrow * 40 + c0;
#pragma unroll
          for (int j = 0; j < 20; ++j) {
            const float eps = __uint_as_float(r[j]) + b3s[c0 + j] + x[j];
            if (live) mo[j] = c2 * (x[j] - c1 * eps);
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 20; ++j) {
          const float eps = __uint_as_float(r[j]) + b3s[c0 + j] + x[j];
          x[j] = c2 * (x[j] - c1 * eps) + sb * pre[j];
        }
        if (s + 1 < S) {
          store_x();
          hand_over(lbar(k3BX));
        }
        const int kidx = a.keep - i;
        if (a.iterates && kidx >= 0) {
          float* stg = reinterpret_cast<float*>(smem + k3OffIt);
          if (warp == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          float* o = stg + m * 40 + c0;
#pragma unroll
          for (int j = 0; j < 20; j += 4) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lim = (e & 1) ? a.a_max : a.w_max;
              v[e] = x[j + e] * lim;
              if (a.clip) v[e] = fminf(fmaxf(v[e], -lim), lim);
            }
            *reinterpret_cast<float4*>(o + j) = make_float4(v[0], v[1], v[2], v[3]);
          }
          fence_proxy_async();
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (warp == 0 && lane == 0) {
            const int r0 = row - m;
            const int left = a.N - r0;
            const int rows = left < 0 ? 0 : (left < 128 ? left : 128);
            if (rows > 0) {
              float* dst = a.iterates + ((size_t)kidx * a.N + (size_t)r0) * 40;
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stg)),
                           "r"(rows * 160)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
      }
      if (live && !a.refine && !a.mu_out) {
        float* xw = a.xin + (size_t)row * PSTL_XIN_LD;
#pragma unroll
        for (int j = 0; j < 20; j += 4) *reinterpret_cast<float4*>(xw + c0 + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
      }
    }
    if (warp == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace
