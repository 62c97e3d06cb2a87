// denoiser_tc.cu — persistent tcgen05 / TMEM engine of the DDPM reverse loop (bf16 operands, fp32
// accumulate and fp32 chain state).  Replaces, for reverse steps without guidance, the per-step
// cat + 3 addmm + ReLU + residual of Net.forward (reference nusc_model.py:118-162) and the posterior
// update + randn_like of diffusion_rollout (reference nusc_train.py:580-629) with ONE launch.
//
// One CTA per SM; each CTA owns 128 chains at a time and runs ALL reverse steps on them:
//   * weights (hoisted W1' 256x48, W2 256x256, W3 40x256; bf16, K-major, 128B-swizzled image built
//     at handle creation) are pulled into shared memory once per CTA with cp.async.bulk (TMA 1D);
//   * layer 1:  D[128x256] (TMEM, fp32)  = X[128x48] (smem, bf16) . W1'^T          tcgen05.mma SS
//     epilogue: + c_scene[scene] + c_t[step], ReLU, pack bf16 -> H (TMEM)           tcgen05.ld / .st
//   * layer 2:  D[128x256]               = H[128x256] (TMEM) . W2^T                 tcgen05.mma TS
//     epilogue: + b2, ReLU, pack bf16 -> H (TMEM)
//   * layer 3:  D3[128x48]               = H . W3^T                                 tcgen05.mma TS
//     epilogue: eps = D3 + b3 + x ; mu ; x <- mu + sqrt(beta) z (Philox or injected) ; the fp32 state
//               never leaves registers; its bf16 image goes back to the X tile for the next step and,
//               in the last K steps, the normalised controls go to HBM.
// Every layer is issued as two N=128 halves with their own completion barriers, and the epilogue warps are
// split the same way (lane quarter = warp%4, column half = warp/4), so the epilogue of half 0 runs under the MMA
// of half 1, layer 2 starts on the first half of H1 while the second is still being packed, and layer 3 is
// half issued before the last hidden columns exist.
// TMEM columns: D [0,256)  H1 [256,384)  H2 [384,512);  D3 aliases H1[0,48) (dead once layer 2 retired) and the
// layer-1 operand X aliases H2[0,32) (dead once layer 3 retired).  Warp 8: barrier init, TMEM alloc, weight load
// and the MMA-issuing lane.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "mlp_common.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kH = 256;        // hidden width (checked at create)
constexpr int kK1 = 64;        // layer-1 depth: [x 40 | hl | stlp 6 | 0 | 8 one-hot class pairs] = 4 UMMA K-steps
constexpr int kN3 = 48;        // layer-3 width padded to a multiple of 16
constexpr int kMaxClasses = 8; // distinct scenes a 128-row tile may span
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;  // + MMA warp + bias warp

// shared-memory image offsets (bytes)
constexpr int kOffW1 = 0;                         // [256 x 64] bf16 SW128; cols 48..63 = per-step bias (hi,lo) per class
constexpr int kOffW2 = kOffW1 + 256 * 128;        // 4 K-blocks of [256 x 64]
constexpr int kOffW3 = kOffW2 + 4 * 256 * 128;    // 4 K-blocks of [48 x 64]
constexpr int kWeightBytes = kOffW3 + 4 * kN3 * 128;
constexpr int kOffB2 = kWeightBytes;              // [256 x 16] bf16, K-major no-swizzle: (hi,lo) of b2 in every class pair
constexpr int kOffB3 = kOffB2 + 256 * 16 * 2;
constexpr int kOffBar = kOffB3 + 64 * 4;
constexpr int kOffIt = kOffBar + 128;             // [128 x 40] fp32: one kept iterate of the tile, bulk-stored to HBM
constexpr int kSmemBytes = kOffIt + kTileM * 40 * 4;
static_assert(kWeightBytes == 188416, "weight image size");
static_assert(kSmemBytes + 1024 <= 227 * 1024, "shared memory budget");

constexpr uint32_t kColD = 0, kColH1 = 256, kColH2 = 384, kColD3 = 256, kColX = 384;  // X: 32 columns = 64 bf16 of layer-1 input
// Timeline instrumentation (clock64 stamps of one tile-step) exists only in builds made with -DPSTL_TC_DEBUG
// (tests/tc_timeline.py builds its own copy); the product library carries none of it.
#ifdef PSTL_TC_DEBUG
#define MSTAMP (a.dbg && blockIdx.x == 0 && it == 2 && lane == 0)
#define STAMP (a.dbg && blockIdx.x == 0 && it == 2 && warp == 0 && lane == 0)
#else
#define MSTAMP false
#define STAMP false
#endif

struct TcState {
  uint8_t* image;    // device: swizzled bf16 weight image of policy_net (kWeightBytes)
  uint8_t* image_r;  // same for rect_net (RefineNet head), or null
  uint8_t* image2;   // pair-engine layout of policy_net (two per-rank halves, denoiser_tc2.cuh)
  uint8_t* image2_r; // same for rect_net, or null
  uint8_t* image_h;  // fp16 operand images of the one-SM engine (PSTL_PRECISION_F16 handles only)
  uint8_t* image_h_r;
  uint8_t* image3;   // split-operand (hi | lo) pair images of policy_net (denoiser_tc3.cuh), PSTL_PRECISION_F16X3 handles only
  uint8_t* image3_r;
  float* zeros;      // 256 zeros: the "time" bias row of the RefineNet pass
  int sm_count;
};

// ---------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one elected lane of a converged warp (ptxas keeps the guarded tcgen05 issue on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  unsigned spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 24)) __trap();  // a protocol bug must fail the launch, not hang the GPU
  } while (!done);
}
// pure polling wait (mbarrier.test_wait never suspends the thread): used where a late wake-up of the
// suspending try_wait was measured to cost ~1000 cycles per step (tc timeline)
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
  uint32_t done;
  unsigned spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) __trap();
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// K-major, no swizzle: 8x16B core matrices; LBO = bytes between K-adjacent cores, SBO = between 8-row groups
__device__ __forceinline__ uint64_t make_desc_flat(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

#define TMEM_LD_X32(taddr, r)                                                                                        \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                      \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
      "%29,%30,%31}, [%32];"                                                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
      : "r"(taddr))

#define TMEM_LD_X16(taddr, r)                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr))

#define TMEM_LD_X4(taddr, r) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr))

#define TMEM_ST_X16(taddr, r)                                                                                    \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
               "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])        \
               : "memory")

#define TMEM_ST_X8(taddr, r)                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), \
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])                     \
               : "memory")

#define TMEM_ST_X4(taddr, r) \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory")
#define TMEM_ST_X2(taddr, r) \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory")

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// relu + round-to-nearest bf16 of (lo, hi) packed as {hi:16 | lo:16}
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// byte offset of element (row, k) of the [256 x 16] no-swizzle bias tile (LBO 128, SBO 256)
__host__ __device__ __forceinline__ int flat16_off(int row, int k) {
  return (row >> 3) * 256 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
}
// (hi, lo) bf16 split of an fp32 value packed as {lo:16 | hi:16}: hi at the lower address
__device__ __forceinline__ uint32_t split_bf16(float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  return (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
}

// The one-SM engine carries its operands either as bf16 (PSTL_PRECISION_BF16: 8 mantissa bits, fp32's exponent range) or as
// fp16 (PSTL_PRECISION_F16: 11 mantissa bits, saturating at 65,504); kind::f16 runs both at the same rate.
template <bool F16>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) {
  uint32_t d;
  if (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_relu16(float lo, float hi) {
  uint32_t d;
  if (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool F16>
__device__ __forceinline__ uint32_t split16(float v) {  // (hi, lo) pieces packed {lo:16 | hi:16}
  if (F16) {
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    const __half l = __float2half_rn(v - __half2float(h));
    return (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
  }
  return split_bf16(v);
}
template <bool F16>
__device__ __forceinline__ constexpr uint32_t make_idesc_t(int M, int N) {  // A/B format: 1 = bf16, 0 = fp16
  return (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row, k) inside a [rows x 64] bf16 K-major SW128 block
__host__ __device__ __forceinline__ int sw128_off(int row, int k) {
  return (row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) & 7) ^ (row & 7)) << 4) + (k & 7) * 2;
}

// ---------------------------------------------------------------------------------------
// weight image (run once per handle)
// ---------------------------------------------------------------------------------------
template <bool F16>
__global__ void k_build_image(const float* __restrict__ w1p, int kin, const float* __restrict__ w2,
                              const float* __restrict__ w3, int n3, uint8_t* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  auto put = [&](int off, float v) {
    if (F16) *reinterpret_cast<__half*>(img + off) = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    else *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(v);
  };
  if (i < 256 * 64) {  // W1'
    const int n = i / 64, k = i % 64;
    put(kOffW1 + sw128_off(n, k), k < kin ? w1p[n * kin + k] : 0.f);
  }
  if (i < 256 * 256) {  // W2: K-block kb = k/64
    const int n = i / 256, k = i % 256;
    put(kOffW2 + (k / 64) * (256 * 128) + sw128_off(n, k % 64), w2[n * 256 + k]);
  }
  if (i < kN3 * 256) {  // W3 rows >= n3 are zero
    const int n = i / 256, k = i % 256;
    put(kOffW3 + (k / 64) * (kN3 * 128) + sw128_off(n, k % 64), n < n3 ? w3[n * 256 + k] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------------------
struct TcArgs {
  const uint8_t* image;
  const float* cscene;  // (n_scenes, 256)  W1f.feat + b1
  const float* ct;      // (steps, 256)     W1t.temb(t)
  const float* b2;
  const float* b3;
  float* xin;           // (N, 48): [x(40) | hl | stlp(6) | 0], x updated in place at the end
  const float* noise;   // injected z (steps-2, N, 40) or null
  float* iterates;      // (keep, N, 40) or null
  float* mu_out;        // (N, 40) or null: single guided step — write the posterior mean, leave x alone
  float c1[128], c2[128], sb[128];  // per reverse step i: (1-a)/sqrt(1-abar), 1/sqrt(a), sqrt(beta)
  int N, rows_per_scene, steps, first_step, last_step, keep, clip;
  float w_max, a_max;
  unsigned long long seed, offset;
  const unsigned long long* offset_dev;  // optional device word added to offset (graph replays draw fresh noise)
  // RefineNet pass (Net.rect_forward, reference nusc_model.py:209-233): one "step", no x update
  int refine;
  const float* u0;      // (N, 40) controls being refined
  const float* scores;  // (N)
  float* out;           // (N, 40)
  long long* dbg;       // optional clock64 timeline of one tile-step (PSTL_TC_DEBUG)
};

template <bool F16>
__global__ void __launch_bounds__(kThreads, 1) k_denoiser_tc(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the shared-window address of the dynamic segment is uniform: keep the 1 KB alignment arithmetic on
  // that integer so every UMMA descriptor stays in uniform registers (no per-issue R2UR / ELECT loop)
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (sbase - sraw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  // d*/h* barriers come in pairs: [+0] column half 0, [+8 bytes] column half 1
  const uint32_t bar_w = smem_u32(&bars[0]), bar_x = smem_u32(&bars[1]), bar_d1 = smem_u32(&bars[2]),
                 bar_h1 = smem_u32(&bars[4]), bar_d2 = smem_u32(&bars[6]), bar_h2 = smem_u32(&bars[8]),
                 bar_d3 = smem_u32(&bars[10]), bar_b = smem_u32(&bars[11]);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(&bars[14]);
  float* b3s = reinterpret_cast<float*>(smem + kOffB3);

  if (warp == kEpiWarps) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_x, kEpiWarps);
      for (uint32_t h = 0; h < 2; ++h) {
        mbar_init(bar_d1 + 8 * h, 1);
        mbar_init(bar_h1 + 8 * h, kEpiWarps / 2);
        mbar_init(bar_d2 + 8 * h, 1);
        mbar_init(bar_h2 + 8 * h, kEpiWarps / 2);
      }
      mbar_init(bar_d3, 1);
      mbar_init(bar_b, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int n = threadIdx.x; n < kH; n += kThreads) {  // layer-2 bias tile: every class pair carries (hi, lo) of b2[n]
    const uint32_t v = split16<F16>(a.b2[n]);
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) *reinterpret_cast<uint32_t*>(smem + kOffB2 + flat16_off(n, 2 * c)) = v;
  }
  for (int i = threadIdx.x; i < 64; i += kThreads) b3s[i] = i < 40 ? a.b3[i] : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // all 512 columns are ours, so the allocation starts at lane 0 / column 0: a literal base keeps the
  // TMEM operands of every tcgen05 instruction immediate
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem = 0u;

  if (warp == kEpiWarps && lane == 0) {
    // weights: one TMA bulk stream into the resident image
    mbar_expect_tx(bar_w, kWeightBytes);
    constexpr int kChunk = 32768;
    for (int off = 0; off < kWeightBytes; off += kChunk) {
      const int n = (kWeightBytes - off) < kChunk ? (kWeightBytes - off) : kChunk;
      bulk_g2s(sbase + kOffW1 + off, a.image + off, n, bar_w);
    }
  }

  const int n_tiles = (a.N + kTileM - 1) / kTileM;
  const int n_steps = a.first_step - a.last_step + 1;
  uint32_t it = 0;  // tile-steps done by this CTA: every barrier completes once per tile-step

  if (warp == kEpiWarps + 1) {
    // ================= bias warp: W1' columns 48..63 for the next tile-step =================
    // It may overwrite them as soon as the layer-1 MMAs of the previous tile-step have retired (d1[1]).
    auto write_bias_cols = [&](int tile, int i) {
      const long long r0 = (long long)tile * kTileM;
      const long long r1 = (r0 + kTileM - 1 < a.N) ? r0 + kTileM - 1 : (long long)a.N - 1;
      const int scene0 = (int)(r0 / a.rows_per_scene);
      const int n_cls = (int)(r1 / a.rows_per_scene) - scene0 + 1;
      const float* ctr = a.ct + (size_t)i * kH;
      for (int c = 0; c < n_cls && c < kMaxClasses; ++c) {
        const float* src = a.cscene + (size_t)(scene0 + c) * kH;
#pragma unroll
        for (int n = lane; n < kH; n += 32)
          *reinterpret_cast<uint32_t*>(smem + kOffW1 + sw128_off(n, 48 + 2 * c)) = split16<F16>(__ldg(src + n) + __ldg(ctr + n));
      }
      fence_proxy_async();
      __syncwarp();
    };
    mbar_wait(bar_w, 0);  // the weight image (whose bias columns are zero) must have landed first
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int s = 0; s < n_steps; ++s, ++it) {
        if (it > 0) mbar_wait(bar_d1 + 8, (it - 1) & 1);
        write_bias_cols(tile, a.first_step - s);
        if (lane == 0) mbar_arrive(bar_b);
      }
    }
  } else if (warp == kEpiWarps) {
    // ================= MMA issuer: the whole warp walks the protocol, one elected lane issues =================
    mbar_wait(bar_w, 0);
    const uint32_t idh = make_idesc_t<F16>(kTileM, kH / 2), id256 = make_idesc_t<F16>(kTileM, kH), id3 = make_idesc_t<F16>(kTileM, kN3);
    const uint64_t dW1 = make_desc(sbase + kOffW1);
    // Biases ride on the tensor pipe.  The X tile carries, in columns 48..63, a pair of ones at the row's
    // scene class; W1' columns 48..63 carry (hi, lo) bf16 halves of c_scene[scene0+class] + c_t[step], rewritten
    // by the bias warp once the previous layer-1 MMA has retired (bar_b).  Layer 2 gets b2 through one extra
    // K-step against the same one-hot columns.
    const uint64_t dB2 = make_desc_flat(sbase + kOffB2, 128, 256);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int s = 0; s < n_steps; ++s, ++it) {
        const uint32_t ph = it & 1;
        mbar_spin(bar_b, ph);
        mbar_spin(bar_x, ph);
        if (MSTAMP) a.dbg[0] = clock64();
        tc_fence_after();
        if (elect_one()) {
          // layer 1, column halves 0 and 1 (W1' rows 128.. start 16 KB into the block: +1024 in descriptor units)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int k = 0; k < kK1 / 16; ++k)  // advancing 32 B inside the 128 B swizzle atom
              mma_ts(tmem + kColD + h * 128, tmem + kColX + k * 8, dW1 + (uint64_t)(h * 1024 + k * 2), idh, k > 0);
            tc_commit(bar_d1 + 8 * h);
          }
        }
        __syncwarp();
        if (MSTAMP) a.dbg[1] = clock64();
        // layer 2, half 0, first eight K-steps: needs H1[:, 0:128) only
        mbar_spin(bar_h1, ph);
        if (MSTAMP) a.dbg[2] = clock64();
        tc_fence_after();
        if (elect_one()) {
          mma_ts(tmem + kColD, tmem + kColX + 24, dB2, idh, 0);  // D = onehot(class) . b2
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t dB = make_desc(sbase + kOffW2 + (k >> 2) * (256 * 128)) + (uint64_t)((k & 3) * 2);
            mma_ts(tmem + kColD, tmem + kColH1 + k * 8, dB, idh, 1);
          }
        }
        __syncwarp();
        mbar_spin(bar_h1 + 8, ph);
        tc_fence_after();
        if (elect_one()) {
          // the bias of half 1 goes first: it is the last reader of X, which E2 (after d2[0]) overwrites
          mma_ts(tmem + kColD + 128, tmem + kColX + 24, dB2 + (uint64_t)256, idh, 0);
          // K-steps 8..15 (the half of H1 that just arrived) go out at full width, N = 256: both column halves in one
          // issue (149 cycles against 2 x 84); column half 0 is then complete
#pragma unroll
          for (int k = 8; k < 16; ++k) {
            const uint64_t dB = make_desc(sbase + kOffW2 + (k >> 2) * (256 * 128)) + (uint64_t)((k & 3) * 2);
            mma_ts(tmem + kColD, tmem + kColH1 + k * 8, dB, id256, 1);
          }
          tc_commit(bar_d2);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t dB = make_desc(sbase + kOffW2 + (k >> 2) * (256 * 128)) + (uint64_t)(1024 + (k & 3) * 2);
            mma_ts(tmem + kColD + 128, tmem + kColH1 + k * 8, dB, idh, 1);
          }
          tc_commit(bar_d2 + 8);
        }
        __syncwarp();
        if (MSTAMP) a.dbg[3] = clock64();
        mbar_spin(bar_h2, ph);
        if (MSTAMP) a.dbg[4] = clock64();
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t dB = make_desc(sbase + kOffW3 + (k >> 2) * (kN3 * 128)) + (uint64_t)((k & 3) * 2);
            mma_ts(tmem + kColD3, tmem + kColH2 + k * 8, dB, id3, k > 0);
          }
        }
        __syncwarp();
        mbar_spin(bar_h2 + 8, ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 8; k < 16; ++k) {
            const uint64_t dB = make_desc(sbase + kOffW3 + (k >> 2) * (kN3 * 128)) + (uint64_t)((k & 3) * 2);
            mma_ts(tmem + kColD3, tmem + kColH2 + k * 8, dB, id3, 1);
          }
          tc_commit(bar_d3);
        }
        if (MSTAMP) a.dbg[5] = clock64();
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3, half = warp >> 2;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    const int c0 = half * 20;  // this thread's 20 state columns
    const unsigned off_base = (unsigned)a.offset + (a.offset_dev ? (unsigned)__ldg(a.offset_dev) : 0u);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row = (long long)tile * kTileM + row_in_tile;
      const bool live = row < a.N;
      const long long rrow = live ? row : (long long)a.N - 1;
      const int scene0 = (int)(((long long)tile * kTileM) / a.rows_per_scene);
      const int cls = (int)(rrow / a.rows_per_scene) - scene0;
      // chain state: fp32 in registers for the whole reverse loop
      float x[20];
      const float* xr = a.xin + rrow * PSTL_XIN_LD;
#pragma unroll
      for (int j = 0; j < 20; j += 4) {  // 16-byte loads: the rows are 192 B apart, scalar loads cost a sector each
        const float4 q = *reinterpret_cast<const float4*>(xr + c0 + j);
        x[j] = q.x; x[j + 1] = q.y; x[j + 2] = q.z; x[j + 3] = q.w;
      }
      // the layer-1 operand lives in TMEM too (columns kColX..+32): bf16 pairs of [x | hl stlp 0 | class one-hots]
      // constant columns 40..47: hl, stlp(6), 0 ; 48..63: ones at this row's scene class.  X shares its columns
      // with H2, so the whole operand is rewritten every step (the half-1 warps carry the constants).
      uint32_t pc[12];
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(xr + 40 + j);
        pc[j / 2] = pack16<F16>(q.x, q.y);
        pc[j / 2 + 1] = pack16<F16>(q.z, q.w);
      }
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c) pc[4 + c] = (c == cls) ? (F16 ? 0x3C003C00u : 0x3F803F80u) : 0u;
      auto store_x = [&]() {
        uint32_t px[10];
#pragma unroll
        for (int j = 0; j < 20; j += 2) px[j / 2] = pack16<F16>(x[j], x[j + 1]);
        TMEM_ST_X8(tmem + lane_addr + kColX + c0 / 2, px);
        TMEM_ST_X2(tmem + lane_addr + kColX + c0 / 2 + 8, (px + 8));
        if (half == 1) {
          TMEM_ST_X4(tmem + lane_addr + kColX + 20, pc);
          TMEM_ST_X8(tmem + lane_addr + kColX + 24, (pc + 4));
        }
      };
      store_x();
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_x);

      for (int s = 0; s < n_steps; ++s, ++it) {
        const uint32_t ph = it & 1;
        const int i = a.first_step - s;  // reverse step index (t == i)
        float zn[20];
        // This step's noise z (20 normals per thread = 5 Philox calls) is drawn in ONE block right after this warp
        // has handed H1 over, i.e. in the shadow of the layer-2 MMA of its column half.  The five counter chains are
        // independent and interleave (a single Philox4x32-10 call is a ~200-cycle dependent chain); the step counter
        // passes through an opaque asm so the compiler cannot hoist the (barrier-independent) arithmetic to the top
        // of the loop, where it would sit on the critical path.
        const int zi = a.steps - 1 - i;
        const float* zr = (a.noise && i > 1) ? a.noise + ((size_t)zi * a.N + rrow) * 40 + c0 : nullptr;
        const bool draw = i > 1 && !a.refine && !a.mu_out;
        auto draw_noise = [&]() {
          unsigned step_ctr = (unsigned)i + off_base;
          asm volatile("" : "+r"(step_ctr)::"memory");
#pragma unroll
          for (int j = 0; j < 20; ++j) zn[j] = 0.f;
          if (draw) {
            if (zr) {
#pragma unroll
              for (int j = 0; j < 20; j += 4) {
                const float4 zz = *reinterpret_cast<const float4*>(zr + j);
                zn[j] = zz.x; zn[j + 1] = zz.y; zn[j + 2] = zz.z; zn[j + 3] = zz.w;
              }
            } else {
              const uint2 key = make_uint2((unsigned)(a.seed & 0xffffffff), (unsigned)(a.seed >> 32));
              uint4 rn[5];
#pragma unroll
              for (int q = 0; q < 5; ++q)
                rn[q] = pstl_philox(make_uint4((unsigned)(rrow & 0xffffffff), (unsigned)(rrow >> 32), (unsigned)(c0 / 4 + q), step_ctr), key);
#pragma unroll
              for (int q = 0; q < 5; ++q) {
                pstl_box_muller(rn[q].x, rn[q].y, zn[4 * q], zn[4 * q + 1]);
                pstl_box_muller(rn[q].z, rn[q].w, zn[4 * q + 2], zn[4 * q + 3]);
              }
            }
          }
        };
        // ---- layers 1 and 2: D (bias already accumulated by the MMA) -> relu -> bf16 -> H ----
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          if (STAMP && layer == 0) a.dbg[21] = clock64();
          mbar_spin((layer == 0 ? bar_d1 : bar_d2) + 8 * half, ph);
          if (STAMP) a.dbg[layer == 0 ? 8 : 11] = clock64();
          tc_fence_after();
          const uint32_t dsrc = tmem + lane_addr + kColD + half * 128;
          const uint32_t hdst = tmem + lane_addr + (layer == 0 ? kColH1 : kColH2) + half * 64;
          // 8 chunks of 16 accumulator columns, the next chunk's TMEM load in flight while one is packed
          uint32_t ra[16], rb[16], p[8];
          TMEM_LD_X16(dsrc, ra);
#pragma unroll
          for (int ch = 0; ch < 8; ch += 2) {
            tmem_wait_ld();
            TMEM_LD_X16(dsrc + (ch + 1) * 16, rb);
#pragma unroll
            for (int j = 0; j < 16; j += 2) p[j / 2] = pack_relu16<F16>(__uint_as_float(ra[j]), __uint_as_float(ra[j + 1]));
            TMEM_ST_X8(hdst + ch * 8, p);
            tmem_wait_ld();
            if (ch + 2 < 8) TMEM_LD_X16(dsrc + (ch + 2) * 16, ra);
#pragma unroll
            for (int j = 0; j < 16; j += 2) p[j / 2] = pack_relu16<F16>(__uint_as_float(rb[j]), __uint_as_float(rb[j + 1]));
            TMEM_ST_X8(hdst + (ch + 1) * 8, p);
          }
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive((layer == 0 ? bar_h1 : bar_h2) + 8 * half);
          if (STAMP) a.dbg[layer == 0 ? 9 : 12] = clock64();
          if (layer == 0) {
            draw_noise();
            if (STAMP) a.dbg[10] = clock64();
          }
        }
        // ---- layer 3: eps, posterior mean, noise, next x ----
        // x' = c2 (x - c1 (D3 + b3 + x)) + sb z = [c2 (1 - c1) x - c2 c1 b3 + sb z] - c2 c1 D3: the bracket is ready
        // before D3 is, so one multiply-add per column is left after the wait
        const float c1 = a.c1[i], c2 = a.c2[i], sb = a.sb[i];
        const float kx = c2 * (1.f - c1), kd = c2 * c1;
        float pre[20];
        if (!a.refine && !a.mu_out) {
#pragma unroll
          for (int j = 0; j < 20; ++j) pre[j] = kx * x[j] - kd * b3s[c0 + j] + sb * zn[j];
        }
        mbar_spin(bar_d3, ph);
        if (STAMP) a.dbg[13] = clock64();
        tc_fence_after();
        uint32_t r[20];
        TMEM_LD_X16(tmem + lane_addr + kColD3 + c0, r);
        TMEM_LD_X4(tmem + lane_addr + kColD3 + c0 + 16, (r + 16));
        tmem_wait_ld();
        if (a.refine) {
          // tanh-interval rescale towards the remaining headroom, only for violating rows
          const float viol = (a.scores[rrow] < 0.f) ? 1.f : 0.f;
          const float* u0r = a.u0 + rrow * 40 + c0;
          float* orow = a.out + rrow * 40 + c0;
#pragma unroll
          for (int j = 0; j < 20; ++j) {
            const float rr = tanhf(__uint_as_float(r[j]) + b3s[c0 + j]);
            const float init = u0r[j];
            const float lim = (j & 1) ? a.a_max : a.w_max;  // c0 is even: parity of j is parity of the column
            const float mk = (rr >= 0.f) ? 1.f : 0.f;
            const float merged = (rr * (init - (-lim))) * (1.f - mk) + (rr * (lim - init)) * mk;
            float o = init + merged * viol;
            if (a.clip) o = fminf(fmaxf(o, -lim), lim);
            if (live) orow[j] = o;
          }
          continue;
        }
        if (a.mu_out) {  // guided step: the mean goes to the guidance kernels, which add the noise afterwards
          float* mo = a.mu_out + rrow * 40 + c0;
#pragma unroll
          for (int j = 0; j < 20; ++j) {
            const float eps = __uint_as_float(r[j]) + b3s[c0 + j] + x[j];
            if (live) mo[j] = c2 * (x[j] - c1 * eps);
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 20; ++j) x[j] = fmaf(-kd, __uint_as_float(r[j]), pre[j]);
        // the next layer-1 operand first: the kept-iterate traffic below then runs in the shadow of the next step
        if (s + 1 < n_steps) {
          store_x();
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_x);
          if (STAMP) a.dbg[14] = clock64();
        }
        const int kidx = a.keep - i;
        if (a.iterates && kidx >= 0) {
          // kept iterate: the tile's 128 x 40 block is contiguous in HBM, so it is assembled in shared memory and
          // leaves as ONE bulk copy (per-thread 8-byte stores cost a 32-byte sector each: 1,800 cycles per step)
          float* stg = reinterpret_cast<float*>(smem + kOffIt);
          if (warp == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          float* o = stg + row_in_tile * 40 + c0;
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
            const long long r0 = (long long)tile * kTileM;
            const int rows = (int)(((long long)a.N - r0 < kTileM) ? (long long)a.N - r0 : kTileM);
            float* dst = a.iterates + ((size_t)kidx * a.N + r0) * 40;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stg)),
                         "r"(rows * 160)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (live && !a.refine && !a.mu_out) {
        float* xw = a.xin + row * PSTL_XIN_LD;
#pragma unroll
        for (int j = 0; j < 20; j += 4) *reinterpret_cast<float4*>(xw + c0 + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
      }
    }
    if (warp == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // kept iterates have landed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

#include "denoiser_tc2.cuh"
#include "denoiser_tc3.cuh"

// (re)build the bf16 weight images from the handle's current fp32 weights: no allocation, no synchronisation
int pstl_tc_refresh(pstl_denoiser* d, cudaStream_t st) {
  TcState* s = (TcState*)d->tc;
  PSTL_CHECK_ARG(s, "engine not created");
  k_build_image<false><<<(256 * 256 + 255) / 256, 256, 0, st>>>(d->w1p, d->kin, d->w.p2_w, d->w.p4_w, d->T2, s->image);
  PSTL_LAUNCH_CHECK();
  if (s->image_h) {
    k_build_image<true><<<(256 * 256 + 255) / 256, 256, 0, st>>>(d->w1p, d->kin, d->w.p2_w, d->w.p4_w, d->T2, s->image_h);
    PSTL_LAUNCH_CHECK();
  }
  k_build_image2<<<(256 * 256 + 255) / 256, 256, 0, st>>>(d->w1p, d->kin, d->w.p2_w, d->w.p4_w, d->T2, s->image2);
  PSTL_LAUNCH_CHECK();
  if (s->image3) {
    k_build_image3<<<(256 * 256 + 255) / 256, 256, 0, st>>>(d->w1p, d->kin, d->w.p2_w, d->w.p4_w, d->T2, s->image3);
    PSTL_LAUNCH_CHECK();
  }
  if (s->image_r) {
    k_build_image<false><<<(256 * 256 + 255) / 256, 256, 0, st>>>(d->r1p, d->kin, d->w.r2_w, d->w.r4_w, d->T2, s->image_r);
    PSTL_LAUNCH_CHECK();
    if (s->image_h_r) {
      k_build_image<true><<<(256 * 256 + 255) / 256, 256, 0, st>>>(d->r1p, d->kin, d->w.r2_w, d->w.r4_w, d->T2, s->image_h_r);
      PSTL_LAUNCH_CHECK();
    }
    k_build_image2<<<(256 * 256 + 255) / 256, 256, 0, st>>>(d->r1p, d->kin, d->w.r2_w, d->w.r4_w, d->T2, s->image2_r);
    PSTL_LAUNCH_CHECK();
    if (s->image3_r) {
      k_build_image3<<<(256 * 256 + 255) / 256, 256, 0, st>>>(d->r1p, d->kin, d->w.r2_w, d->w.r4_w, d->T2, s->image3_r);
      PSTL_LAUNCH_CHECK();
    }
  }
  return PSTL_OK;
}

int pstl_tc_create(pstl_denoiser* d) {
  if (d->w.hidden != kH || d->T2 != 40 || d->kin != 47) {
    pstl_set_error("the tcgen05 engines (PSTL_PRECISION_BF16 / _F16X3) are built for hidden=256, nt=20 (got hidden=%d, nt=%d)", d->w.hidden, d->w.T);
    return PSTL_ERR_UNSUPPORTED;
  }
  int dev = 0, cc_major = 0, sms = 0;
  PSTL_CUDA(cudaGetDevice(&dev));
  PSTL_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  PSTL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major != 10) {
    pstl_set_error("PSTL_PRECISION_BF16 needs an sm_100 device (tcgen05)");
    return PSTL_ERR_UNSUPPORTED;
  }
  TcState* s = new TcState();
  s->sm_count = sms;
  s->image = s->image_r = s->image2 = s->image2_r = s->image3 = s->image3_r = s->image_h = s->image_h_r = nullptr;
  s->zeros = nullptr;
  PSTL_CUDA(cudaMalloc(&s->image, kWeightBytes));
  PSTL_CUDA(cudaMemset(s->image, 0, kWeightBytes));
  PSTL_CUDA(cudaMalloc(&s->image2, 2 * k2WeightBytes));
  PSTL_CUDA(cudaMemset(s->image2, 0, 2 * k2WeightBytes));
  if (d->precision == PSTL_PRECISION_F16X3) {
    PSTL_CUDA(cudaMalloc(&s->image3, 4 * k2WeightBytes));
    PSTL_CUDA(cudaMemset(s->image3, 0, 4 * k2WeightBytes));
  }
  if (d->precision == PSTL_PRECISION_F16) {
    PSTL_CUDA(cudaMalloc(&s->image_h, kWeightBytes));
    PSTL_CUDA(cudaMemset(s->image_h, 0, kWeightBytes));
  }
  PSTL_CUDA(cudaMalloc(&s->zeros, kH * sizeof(float)));
  PSTL_CUDA(cudaMemset(s->zeros, 0, kH * sizeof(float)));
  if (d->r1p && d->w.rect_hidden == kH) {
    PSTL_CUDA(cudaMalloc(&s->image_r, kWeightBytes));
    PSTL_CUDA(cudaMemset(s->image_r, 0, kWeightBytes));
    PSTL_CUDA(cudaMalloc(&s->image2_r, 2 * k2WeightBytes));
    PSTL_CUDA(cudaMemset(s->image2_r, 0, 2 * k2WeightBytes));
    if (d->precision == PSTL_PRECISION_F16X3) {
      PSTL_CUDA(cudaMalloc(&s->image3_r, 4 * k2WeightBytes));
      PSTL_CUDA(cudaMemset(s->image3_r, 0, 4 * k2WeightBytes));
    }
    if (d->precision == PSTL_PRECISION_F16) {
      PSTL_CUDA(cudaMalloc(&s->image_h_r, kWeightBytes));
      PSTL_CUDA(cudaMemset(s->image_h_r, 0, kWeightBytes));
    }
  }
  d->tc = s;
  int rc = pstl_tc_refresh(d, nullptr);
  if (rc) return rc;
  PSTL_CUDA(cudaDeviceSynchronize());
  PSTL_CUDA(cudaFuncSetAttribute(k_denoiser_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024));
  PSTL_CUDA(cudaFuncSetAttribute(k_denoiser_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024));
  PSTL_CUDA(cudaFuncSetAttribute(k_denoiser_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes + 1024));
  PSTL_CUDA(cudaFuncSetAttribute(k_denoiser_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, k3SmemBytes + 1024));
  d->tc = s;
  return PSTL_OK;
}

void pstl_tc_destroy(pstl_denoiser* d) {
  TcState* s = (TcState*)d->tc;
  if (!s) return;
  cudaFree(s->image);
  cudaFree(s->image_r);
  cudaFree(s->image2);
  cudaFree(s->image2_r);
  cudaFree(s->image3);
  cudaFree(s->image_h);
  cudaFree(s->image_h_r);
  cudaFree(s->image3_r);
  cudaFree(s->zeros);
  delete s;
  d->tc = nullptr;
}

// Engine choice.  The pair engine (denoiser_tc2.cuh) wins once there is more than one wave of 128-row tiles; below
// that the one-SM engine spreads a small batch over more SMs.  PSTL_TC_ENGINE=1|2 (read once) pins one for tests.
static int tc_engine_for(const TcState* s, int engine, int N, int rows_per_scene) {
  static const int forced = [] {
    const char* e = getenv("PSTL_TC_ENGINE");
    return e ? atoi(e) : 0;
  }();
  const bool pair_ok = (k2TileM + rows_per_scene - 1) / rows_per_scene + 1 <= kMaxClasses && s->sm_count >= 2;
  if (engine == 0) engine = forced;
  if (engine == 1 || !pair_ok) return 1;
  if (engine == 2) return 2;
  return 1;
}

static bool tc_pair_fits(const TcState* s, int rows_per_scene) {
  return (k2TileM + rows_per_scene - 1) / rows_per_scene + 1 <= kMaxClasses && s->sm_count >= 2;
}

static int tc_launch(const TcState* s, int engine, TcArgs& a, int rows_per_scene, const uint8_t* image1, const uint8_t* image2,
                     const uint8_t* image3, const uint8_t* image_h, cudaStream_t st) {
  if (image3) {  // PSTL_PRECISION_F16X3: the split-operand pair engine
    PSTL_CHECK_ARG(tc_pair_fits(s, rows_per_scene), "rows_per_scene too small for the split-operand tcgen05 tile");
    a.image = image3;
    const int n_tiles = (a.N + k2TileM - 1) / k2TileM;
    const int pairs = n_tiles < s->sm_count / 2 ? n_tiles : s->sm_count / 2;
    k_denoiser_tc3<<<2 * pairs, kThreads, k3SmemBytes + 1024, st>>>(a);
    PSTL_LAUNCH_CHECK();
    return PSTL_OK;
  }
  if (image_h) {  // PSTL_PRECISION_F16: the one-SM engine on fp16 operands
    a.image = image_h;
    const int n_tiles = (a.N + kTileM - 1) / kTileM;
    const int grid = n_tiles < s->sm_count ? n_tiles : s->sm_count;
    k_denoiser_tc<true><<<grid, kThreads, kSmemBytes + 1024, st>>>(a);
    PSTL_LAUNCH_CHECK();
    return PSTL_OK;
  }
  if (tc_engine_for(s, engine, a.N, rows_per_scene) == 2) {
    a.image = image2;
    const int n_tiles = (a.N + k2TileM - 1) / k2TileM;
    const int pairs = n_tiles < s->sm_count / 2 ? n_tiles : s->sm_count / 2;
    k_denoiser_tc2<<<2 * pairs, kThreads, k2SmemBytes + 1024, st>>>(a);
  } else {
    a.image = image1;
    const int n_tiles = (a.N + kTileM - 1) / kTileM;
    const int grid = n_tiles < s->sm_count ? n_tiles : s->sm_count;
    k_denoiser_tc<false><<<grid, kThreads, kSmemBytes + 1024, st>>>(a);
  }
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

// runs reverse steps first_step .. last_step (inclusive, descending) on xin in place
int pstl_tc_sample(pstl_denoiser* d, const float* cscene, int rows_per_scene, const float* ct, float* xin, int N,
                   const float* sched, int steps, const float* noise, unsigned long long seed,
                   unsigned long long offset, float w_max, float a_max, int clip, int keep_last_k, float* iterates_out,
                   int first_step, int last_step, float* mu_out, cudaStream_t st) {
  TcState* s = (TcState*)d->tc;
  PSTL_CHECK_ARG(s, "engine not created");
  PSTL_CHECK_ARG(steps <= 128, "at most 128 diffusion steps");
  PSTL_CHECK_ARG((kTileM + rows_per_scene - 1) / rows_per_scene + 1 <= kMaxClasses, "rows_per_scene too small for the tcgen05 tile");
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.cscene = cscene; a.ct = ct; a.b2 = d->w.p2_b; a.b3 = d->w.p4_b; a.xin = xin;
  a.noise = noise; a.iterates = keep_last_k > 0 ? iterates_out : nullptr;
  a.mu_out = mu_out;
  const float *beta = sched, *alpha = sched + steps, *abar = sched + 2 * steps;
  for (int i = 1; i < steps; ++i) {
    a.c1[i] = (1.0f - alpha[i]) / sqrtf(1.0f - abar[i]);
    a.c2[i] = 1.0f / sqrtf(alpha[i]);
    a.sb[i] = sqrtf(beta[i]);
  }
  a.N = N; a.rows_per_scene = rows_per_scene; a.steps = steps; a.first_step = first_step; a.last_step = last_step;
  a.keep = keep_last_k; a.clip = clip; a.w_max = w_max; a.a_max = a_max; a.seed = seed; a.offset = offset; a.offset_dev = d->offset_dev;
#ifdef PSTL_TC_DEBUG
  long long* dbg = nullptr;
  if (getenv("PSTL_TC_DEBUG")) { cudaMalloc(&dbg, 48 * sizeof(long long)); cudaMemset(dbg, 0, 48 * sizeof(long long)); }
  a.dbg = dbg;
#endif
  int rc = tc_launch(s, d->tc_engine, a, rows_per_scene, s->image, s->image2, s->image3, s->image_h, st);
#ifdef PSTL_TC_DEBUG
  if (dbg) {
    cudaStreamSynchronize(st);
    long long h[48];
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    if (tc_engine_for(s, d->tc_engine, N, rows_per_scene) == 2) {
      const long long t0 = h[0];
      auto R = [&](int i) { return h[i] ? h[i] - t0 : -1LL; };
      fprintf(stderr, "[pstl pair-engine timeline, pair 0, 3rd step, cycles rel. to the MMA warp reaching L1(A)]\n"
                      "  MMA : L1A ready %lld issued %lld | L3B h2a %lld h2b %lld issued %lld | L2A h1 %lld issued %lld | L1B at %lld ready %lld issued %lld |"
                      " L3A h2a %lld h2b %lld issued %lld | L2B h1 %lld issued %lld\n"
                      "  EPI : E1A d1 %lld done %lld | E3B d3 %lld done %lld | prepA %lld | E2A d2 %lld part0 %lld done %lld | E1B d1 %lld done %lld |"
                      " E3A d3 %lld done %lld | prepB %lld | E2B d2 %lld part0 %lld done %lld\n",
              R(1), R(2), R(3), R(4), R(15), R(5), R(6), R(7), R(8), R(9), R(10), R(11), R(12), R(13), R(14),
              R(20), R(21), R(36), R(37), R(28), R(23), R(25), R(24), R(30), R(31), R(26), R(27), R(38), R(33), R(35), R(34));
    } else
    fprintf(stderr, "[pstl tc timeline, CTA 0, 3rd tile-step, cycles rel. to bar_x ready]\n  MMA : x %lld | mma1 issued %lld | h1[0] %lld | mma2 issued %lld | h2[0] %lld | mma3 issued %lld\n"
                    "  EPI0: d1 %lld | epi1 done %lld | noise done %lld | d2 %lld | epi2 done %lld | d3 %lld | epi3 done %lld\n",
            0LL, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[8] - h[0], h[9] - h[0], h[10] - h[0],
            h[11] - h[0], h[12] - h[0], h[13] - h[0], h[14] - h[0]);
    cudaFree(dbg);
  }
#endif
  return rc;
}

// RefineNet head on the same engine: xin rows = [fused(40) | hl | stlp], one pass, tanh-interval epilogue
int pstl_tc_refine(pstl_denoiser* d, const float* cscene, int rows_per_scene, const float* xin, int N, const float* u0,
                   const float* scores, float w_max, float a_max, int clip_rect, float* out, cudaStream_t st) {
  TcState* s = (TcState*)d->tc;
  PSTL_CHECK_ARG(s && s->image_r, "RefineNet image not built");
  PSTL_CHECK_ARG((kTileM + rows_per_scene - 1) / rows_per_scene + 1 <= kMaxClasses, "rows_per_scene too small for the tcgen05 tile");
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.cscene = cscene; a.ct = s->zeros; a.b2 = d->w.r2_b; a.b3 = d->w.r4_b;
  a.xin = const_cast<float*>(xin);
  a.N = N; a.rows_per_scene = rows_per_scene; a.steps = 2; a.first_step = 0; a.last_step = 0;  // one pass, ct row 0
  a.clip = clip_rect; a.w_max = w_max; a.a_max = a_max;
  a.refine = 1; a.u0 = u0; a.scores = scores; a.out = out;
  return tc_launch(s, d->tc_engine, a, rows_per_scene, s->image_r, s->image2_r, s->image3_r, s->image_h_r, st);
}

bool pstl_tc_fits(pstl_denoiser* d, int rows_per_scene) { return d->tc && tc_pair_fits((TcState*)d->tc, rows_per_scene); }

bool pstl_tc_has_refine(pstl_denoiser* d) { return d->tc && ((TcState*)d->tc)->image_r; }
