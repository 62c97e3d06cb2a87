// denoiser_tc2.cuh — the CTA-PAIR tcgen05 engine of the DDPM reverse loop: two 256-row tiles in flight per SM pair.
// Included by denoiser_tc.cu (same translation unit: PTX helpers, TcArgs and the weight handle live there).
//
// Why a second engine.  k_denoiser_tc keeps one 128-row tile per SM and all of it on chip, but that tile uses all
// 512 TMEM columns (fp32 D 256 + bf16 H1 128 + H2 128), so the tensor pipe idles through every epilogue hand-over
// (57 % busy, profiles/r2_ncu_k_denoiser_tc.txt).  A second tile needs 256 more accumulator columns, so the hidden
// activations have to leave TMEM, and 64 KB of H per tile does not fit in shared memory next to 184 KB of weights.
// tcgen05.mma.cta_group::2 solves both: the two SMs of a cluster share one M=256 MMA, each CTA supplies HALF of every
// weight matrix (N/2 rows of the K-major B operand: 92 KB) and its own 128 rows of A, and keeps its own 128 x 256
// fp32 accumulator.  Per CTA:
//   TMEM   D[slot 0] = columns [0,256), D[slot 1] = [256,512); layer 3's 48 columns alias D[slot][0,48)
//   SMEM   W1' half 16 KB | W2 half 64 KB | W3 half 12 KB | bias tile (slot 1) 4 KB | H[slot 0] 64 KB | H[slot 1] 64 KB
//          H is the A operand of every layer (K-major, 128 B swizzle, four 64-column K-blocks): X (layer-1 input,
//          K-block 0) -> H1 -> H2 are written in place by the epilogue warps (each layer's reads have retired before
//          the next epilogue starts), the kept iterates are staged in K-blocks 2..3 for their bulk store.
// One MMA-issuing thread (leader CTA) walks a static interleave of the two slots
//   L1(A,s) · L3(B,s-1) · L2(A,s) · L1(B,s) · L3(A,s) · L2(B,s)
// and the eight epilogue warps of BOTH CTAs follow it (E1(A) · E3(B) · noise(A) · E2(A) · E1(B) · E3(A) · noise(B) ·
// E2(B)), so every epilogue runs under the other slot's MMAs.  Hand-overs to the leader are cluster-scope mbarrier
// arrivals (count = 16 warps), MMA completion comes back to both CTAs by a multicast tcgen05.commit.
// Biases: layer 1 as in the one-SM engine (one-hot scene-class columns of X against a per-step bias K-step that the
// bias warp of each CTA rewrites for its half of N); b2 / b3 are added in the epilogue (fp32).
// Row -> (Philox counter, injected-noise index) is the same function as in k_denoiser_tc: both engines draw the same z.
#pragma once

namespace {

constexpr int k2TileM = 256;                       // rows per tile over the pair (128 per CTA)
constexpr int k2OffW1 = 0;                         // [128 n x 64 k] bf16 SW128; k 48..63 = slot-0 bias K-step
constexpr int k2OffW2 = k2OffW1 + 128 * 128;       // 4 K-blocks of [128 x 64]
constexpr int k2OffW3 = k2OffW2 + 4 * 128 * 128;   // 4 K-blocks of [24 x 64]
constexpr int k2WeightBytes = k2OffW3 + 4 * 24 * 128;
constexpr int k2OffBiasB = k2WeightBytes;          // [128 x 16] bf16 K-major no-swizzle: slot-1 bias K-step
constexpr int k2OffH = k2OffBiasB + 128 * 16 * 2;  // 2 slots x 4 K-blocks x [128 m x 64 k]
constexpr int k2HBytes = 4 * 128 * 128;
constexpr int k2OffStage = 2 * 128 * 128;          // kept-iterate staging inside H[slot]: K-blocks 2..3
constexpr int k2OffB2 = k2OffH + 2 * k2HBytes;     // 256 fp32
constexpr int k2OffB3 = k2OffB2 + 256 * 4;         // 64 fp32
constexpr int k2OffBar = k2OffB3 + 64 * 4;
constexpr int k2SmemBytes = k2OffBar + 256;
static_assert(k2WeightBytes == 94208 && k2OffH % 1024 == 0, "pair image layout");
static_assert(k2SmemBytes + 1024 <= 227 * 1024, "shared memory budget");

// barrier slots (8 bytes each): 0 = weights; per slot j at 1 + 10 j:
enum { kBX = 0, kBH1 = 1, kBH2a = 2, kBH2b = 3, kBD1 = 4, kBD2 = 5, kBD3 = 6, kBB = 7 };

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier of (possibly) the other CTA.  Default (CTA-scope) release, as CUTLASS's ClusterBarrier::arrive:
// what the leader's MMAs read afterwards is THIS CTA's own shared memory, written by this thread and already made
// visible to the async proxy by fence.proxy.async, so nothing has to become visible beyond the SM.  The
// .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR in front of every arrival (measured: +500..1000 cycles per
// hand-over, the pair engine ran at 4.2 ms instead of the one-SM engine's 3.05 ms).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_spin_cluster(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mma2_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// completion of all MMAs issued so far -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// clock64 timeline of one step of pair 0 (developer builds with -DPSTL_TC_DEBUG only; tests/tc_timeline.py)
#ifdef PSTL_TC_DEBUG
#define P2STAMP(on, idx) do { if ((on) && a.dbg) a.dbg[idx] = clock64(); } while (0)
#else
#define P2STAMP(on, idx) do { } while (0)
#endif

template <int V>
struct IntC {
  static constexpr int value = V;
};

// pair image: rank r holds rows n in [128 r, 128 r + 128) of W1' and W2 and rows [24 r, 24 r + 24) of the padded W3
__global__ void k_build_image2(const float* __restrict__ w1p, int kin, const float* __restrict__ w2,
                               const float* __restrict__ w3, int n3, uint8_t* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  auto put = [&](int rank, int off, float v) {
    *reinterpret_cast<__nv_bfloat16*>(img + (size_t)rank * k2WeightBytes + off) = __float2bfloat16_rn(v);
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_denoiser_tc2(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;  // same offset in both CTAs (same kernel, same static layout)
  uint8_t* smem = smem_raw + (sbase - sraw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const uint32_t bar0 = sbase + k2OffBar;
  const uint32_t bar_w = bar0;
  auto bar = [&](int slot, int which) { return bar0 + 8u * (1 + 10 * slot + which); };
  const uint32_t lead0 = mapa_rank(bar0, 0);  // the leader's barrier block in the cluster window
  auto lbar = [&](int slot, int which) { return lead0 + 8u * (1 + 10 * slot + which); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + k2OffBar + 8 * 24);
  float* b2s = reinterpret_cast<float*>(smem + k2OffB2);
  float* b3s = reinterpret_cast<float*>(smem + k2OffB3);

  if (warp == kEpiWarps) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      for (int j = 0; j < 2; ++j) {
        mbar_init(bar(j, kBX), 2 * kEpiWarps);
        mbar_init(bar(j, kBH1), 2 * kEpiWarps);
        mbar_init(bar(j, kBH2a), 2 * kEpiWarps);
        mbar_init(bar(j, kBH2b), 2 * kEpiWarps);
        mbar_init(bar(j, kBD1), 1);
        mbar_init(bar(j, kBD2), 1);
        mbar_init(bar(j, kBD3), 1);
        mbar_init(bar(j, kBB), 2);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kH; i += kThreads) b2s[i] = a.b2[i];
  for (int i = threadIdx.x; i < 64; i += kThreads) b3s[i] = i < 40 ? a.b3[i] : 0.f;
  for (int i = threadIdx.x; i < 128 * 16 * 2 / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem + k2OffBiasB)[i] = 0u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs are initialised before anyone arrives remotely
  tc_fence_after();
  if (*tmem_slot != 0u) __trap();  // all 512 columns: the allocation starts at column 0
  constexpr uint32_t tmem = 0u;

  if (warp == kEpiWarps && lane == 0) {
    mbar_expect_tx(bar_w, k2WeightBytes);
    constexpr int kChunk = 23552;  // 4 equal bulk copies
    const uint8_t* src = a.image + (size_t)rank * k2WeightBytes;
    for (int off = 0; off < k2WeightBytes; off += kChunk) bulk_g2s(sbase + off, src + off, kChunk, bar_w);
  }

  const int n_tiles = (a.N + k2TileM - 1) / k2TileM;
  const int S = a.first_step - a.last_step + 1;  // reverse steps per tile
  const int n_rounds = (n_tiles + 2 * n_pairs - 1) / (2 * n_pairs);

  if (warp == kEpiWarps + 1) {
    // ================= bias warp (both CTAs): this CTA's half of the layer-1 bias K-step of each slot =================
    mbar_wait(bar_w, 0);  // the weight image (zero bias columns) has landed
    for (int r = 0; r < n_rounds; ++r) {
      const int tA = r * 2 * n_pairs + pair, tB = tA + n_pairs;
      if (tA >= n_tiles) break;
      const bool vB = tB < n_tiles;
      for (int s = 0; s < S; ++s) {
        const int k = r * S + s;  // tile-step index of either slot
        const int i = a.first_step - s;
        const float* ctr = a.ct + (size_t)i * kH + rank * 128;
#pragma unroll 1
        for (int slot = 0; slot < 2; ++slot) {
          if (slot == 1 && !vB) break;
          if (k > 0) mbar_wait(bar(slot, kBD1), (k - 1) & 1);  // the previous layer-1 MMAs of this slot have retired
          const long long r0 = (long long)(slot ? tB : tA) * k2TileM;
          const long long r1 = (r0 + k2TileM - 1 < a.N) ? r0 + k2TileM - 1 : (long long)a.N - 1;
          const int scene0 = (int)(r0 / a.rows_per_scene);
          const int n_cls = (int)(r1 / a.rows_per_scene) - scene0 + 1;
          for (int c = 0; c < n_cls && c < kMaxClasses; ++c) {
            const float* src = a.cscene + (size_t)(scene0 + c) * kH + rank * 128;
#pragma unroll
            for (int n = lane; n < 128; n += 32) {
              const uint32_t v = split_bf16(__ldg(src + n) + __ldg(ctr + n));
              uint8_t* dst = slot ? smem + k2OffBiasB + flat16_off(n, 2 * c) : smem + k2OffW1 + sw128_off(n, 48 + 2 * c);
              *reinterpret_cast<uint32_t*>(dst) = v;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lbar(slot, kBB));
        }
      }
    }
  } else if (warp == kEpiWarps) {
    if (rank == 0) {
      // ================= MMA issuer (leader CTA): static interleave of the two slots =================
      mbar_wait(bar_w, 0);
      constexpr uint32_t idN = make_idesc(k2TileM, kH), id3 = make_idesc(k2TileM, kN3);
      const uint64_t dW1 = make_desc(sbase + k2OffW1);
      const uint64_t dBB = make_desc_flat(sbase + k2OffBiasB, 128, 256);
      auto L1 = [&](auto sc, uint32_t ph, bool dg) {
        constexpr int J = decltype(sc)::value;
        P2STAMP(dg, J ? 7 : 0);
        mbar_spin_cluster(bar(J, kBB), ph);
        mbar_spin_cluster(bar(J, kBX), ph);
        P2STAMP(dg, J ? 8 : 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dX = make_desc(sbase + k2OffH + J * k2HBytes);
#pragma unroll
          for (int k = 0; k < 3; ++k) mma2_ss(tmem + J * 256, dX + (uint64_t)(2 * k), dW1 + (uint64_t)(2 * k), idN, k > 0);
          mma2_ss(tmem + J * 256, dX + 6, J ? dBB : dW1 + 6, idN, 1);  // one-hot class columns x this step's bias rows
          tc_commit2(bar(J, kBD1));
        }
        __syncwarp();
        P2STAMP(dg, J ? 9 : 2);
      };
      auto L2 = [&](auto sc, uint32_t ph, bool dg) {
        constexpr int J = decltype(sc)::value;
        mbar_spin_cluster(bar(J, kBH1), ph);
        P2STAMP(dg, J ? 13 : 5);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const uint64_t dA = make_desc(sbase + k2OffH + J * k2HBytes + (k >> 2) * (128 * 128)) + (uint64_t)((k & 3) * 2);
            const uint64_t dB = make_desc(sbase + k2OffW2 + (k >> 2) * (128 * 128)) + (uint64_t)((k & 3) * 2);
            mma2_ss(tmem + J * 256, dA, dB, idN, k > 0);
          }
          tc_commit2(bar(J, kBD2));
        }
        __syncwarp();
        P2STAMP(dg, J ? 14 : 6);
      };
      auto L3 = [&](auto sc, uint32_t ph, bool dg) {
        constexpr int J = decltype(sc)::value;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          mbar_spin_cluster(bar(J, part ? kBH2b : kBH2a), ph);
          P2STAMP(dg, (J ? 3 : 10) + part);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = part * 8; k < part * 8 + 8; ++k) {
              const uint64_t dA = make_desc(sbase + k2OffH + J * k2HBytes + (k >> 2) * (128 * 128)) + (uint64_t)((k & 3) * 2);
              const uint64_t dB = make_desc(sbase + k2OffW3 + (k >> 2) * (24 * 128)) + (uint64_t)((k & 3) * 2);
              mma2_ss(tmem + J * 256, dA, dB, id3, k > 0);
            }
            if (part) tc_commit2(bar(J, kBD3));
          }
          __syncwarp();
        }
        P2STAMP(dg, J ? 15 : 12);
      };
      for (int r = 0; r < n_rounds; ++r) {
        const int tA = r * 2 * n_pairs + pair, tB = tA + n_pairs;
        if (tA >= n_tiles) break;
        const bool vB = tB < n_tiles;
        for (int s = 0; s < S; ++s) {
          const uint32_t ph = (uint32_t)(r * S + s) & 1u;
          const bool dg = blockIdx.x == 0 && r == 0 && s == 2 && lane == 0;
          L1(IntC<0>{}, ph, dg);
          if (vB && s > 0) L3(IntC<1>{}, ph ^ 1u, dg);
          L2(IntC<0>{}, ph, dg);
          if (vB) L1(IntC<1>{}, ph, dg);
          L3(IntC<0>{}, ph, dg);
          if (vB) L2(IntC<1>{}, ph, dg);
        }
        if (vB) L3(IntC<1>{}, (uint32_t)(r * S + S - 1) & 1u, false);
      }
    }
  } else {
    // ================= epilogue warps (both CTAs) =================
    const int q = warp & 3, ch = warp >> 2;  // TMEM lane quarter, column sub-half
    const int m = q * 32 + lane;             // row inside this CTA's half of the tile
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    const int c0 = ch * 20;                  // this thread's 20 state columns
    const uint32_t hrow = sbase + k2OffH + (m >> 3) * 1024 + (m & 7) * 128;  // this row's 128 B line in K-block 0 of slot 0
    const uint32_t sw = (uint32_t)(m & 7);
    const unsigned off_base = (unsigned)a.offset + (a.offset_dev ? (unsigned)__ldg(a.offset_dev) : 0u);
    const uint2 key = make_uint2((unsigned)(a.seed & 0xffffffff), (unsigned)(a.seed >> 32));

    float xA[20], xB[20], pre[20];
    uint32_t pcA[4], pcB[4];
    long long rowA = 0, rowB = 0;
    int clsA = 0, clsB = 0;

    auto tile_rows = [&](int tile, long long& row, long long& rrow, bool& live) {
      row = (long long)tile * k2TileM + rank * 128 + m;
      live = row < a.N;
      rrow = live ? row : (long long)a.N - 1;
    };
    // bf16 image of the layer-1 operand row: [x 40 | hl stlp(6) 0 | one-hot class pairs] -> K-block 0 of H[slot]
    auto store_x = [&](auto sc) {
      constexpr int J = decltype(sc)::value;
      float(&x)[20] = *(J ? &xB : &xA);
      const uint32_t(&pc)[4] = *(J ? &pcB : &pcA);
      const int cls = J ? clsB : clsA;
      uint32_t px[10];
#pragma unroll
      for (int j = 0; j < 20; j += 2) px[j / 2] = pack_bf16(x[j], x[j + 1]);
      const uint32_t base = hrow + J * k2HBytes;
      if (ch == 0) {
        st_shared_v4(base + ((0u ^ sw) << 4), px[0], px[1], px[2], px[3]);
        st_shared_v4(base + ((1u ^ sw) << 4), px[4], px[5], px[6], px[7]);
        st_shared_v2(base + ((2u ^ sw) << 4), px[8], px[9]);
      } else {
        st_shared_v2(base + ((2u ^ sw) << 4) + 8, px[0], px[1]);
        st_shared_v4(base + ((3u ^ sw) << 4), px[2], px[3], px[4], px[5]);
        st_shared_v4(base + ((4u ^ sw) << 4), px[6], px[7], px[8], px[9]);
        st_shared_v4(base + ((5u ^ sw) << 4), pc[0], pc[1], pc[2], pc[3]);
        uint32_t oh[8];
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) oh[c] = (c == cls) ? 0x3F803F80u : 0u;
        st_shared_v4(base + ((6u ^ sw) << 4), oh[0], oh[1], oh[2], oh[3]);
        st_shared_v4(base + ((7u ^ sw) << 4), oh[4], oh[5], oh[6], oh[7]);
      }
    };
    auto hand_over = [&](uint32_t leader_bar) {
      fence_proxy_async();  // generic-proxy writes of H / X -> visible to the MMA's async-proxy reads
      tc_fence_before();    // and our TMEM reads are ordered before the MMAs the arrival releases
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_bar);
    };
    // ---- layers 1 and 2: D -> (+b2) -> relu -> bf16 -> H[slot] (in place) ----
    auto e12 = [&](auto sc, int layer, uint32_t ph, bool dg) {
      constexpr int J = decltype(sc)::value;
      mbar_spin(bar(J, layer == 0 ? kBD1 : kBD2), ph);
      P2STAMP(dg, 20 + J * 10 + layer * 3);
      tc_fence_after();
      // this thread: columns [128 p + 64 ch, +64) of part p = K-block 2p + ch, all eight 16-byte chunks of its row
      uint32_t ra[16], rb[16];
      const uint32_t dsrc = tmem + lane_addr + J * 256 + ch * 64;
      const uint32_t hdst = hrow + J * k2HBytes + ch * (128 * 128);
      auto emit = [&](const uint32_t(&r)[16], int cidx) {  // cidx: 16-column chunk 0..7 (0..3 = part 0)
        const int p = cidx >> 2, c16 = cidx & 3;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (layer == 1) {
          const float4* bb = reinterpret_cast<const float4*>(b2s + p * 128 + ch * 64 + c16 * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b = bb[j];
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) w[j / 2] = pack_relu_bf16(v[j], v[j + 1]);
        const uint32_t kb = hdst + p * (2 * 128 * 128);
        st_shared_v4(kb + (((uint32_t)(2 * c16) ^ sw) << 4), w[0], w[1], w[2], w[3]);
        st_shared_v4(kb + (((uint32_t)(2 * c16 + 1) ^ sw) << 4), w[4], w[5], w[6], w[7]);
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
        // layer 2 hands its first 128 columns over early: layer 3's first eight K-steps only need those, and they
        // write D[0,48), which nobody reads any more (the load in flight is of columns >= 128)
        if (layer == 1 && cidx == 2) {
          hand_over(lbar(J, kBH2a));
          P2STAMP(dg, 20 + J * 10 + 5);
        }
      }
      hand_over(lbar(J, layer == 0 ? kBH1 : kBH2b));
      P2STAMP(dg, 20 + J * 10 + layer * 3 + 1);
    };
    // ---- this step's noise and the part of the posterior update that does not need layer 3 ----
    auto prepare = [&](auto sc, int s) {
      constexpr int J = decltype(sc)::value;
      float(&x)[20] = *(J ? &xB : &xA);
      const long long row = J ? rowB : rowA;
      const long long rrow = row < a.N ? row : (long long)a.N - 1;
      const int i = a.first_step - s;
      const bool draw = i > 1 && !a.refine && !a.mu_out;
#pragma unroll
      for (int j = 0; j < 20; ++j) pre[j] = 0.f;
      if (draw) {
        if (a.noise) {
          const float* zr = a.noise + ((size_t)(a.steps - 1 - i) * a.N + rrow) * 40 + c0;
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
          for (int k = 0; k < 5; ++k)
            rn[k] = pstl_philox(make_uint4((unsigned)(rrow & 0xffffffff), (unsigned)(rrow >> 32), (unsigned)(c0 / 4 + k), step_ctr), key);
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            pstl_box_muller(rn[k].x, rn[k].y, pre[4 * k], pre[4 * k + 1]);
            pstl_box_muller(rn[k].z, rn[k].w, pre[4 * k + 2], pre[4 * k + 3]);
          }
        }
      }
      // x' = c2 (x - c1 (D3 + b3 + x)) + sb z = [c2 (1 - c1) x - c2 c1 b3 + sb z] - c2 c1 D3
      const float c1 = a.c1[i], c2 = a.c2[i], sb = a.sb[i];
      const float kx = c2 * (1.f - c1), kd = c2 * c1;
#pragma unroll
      for (int j = 0; j < 20; ++j) pre[j] = kx * x[j] - kd * b3s[c0 + j] + sb * pre[j];
    };
    // ---- layer 3: eps, posterior mean, noise, next x ----
    auto e3 = [&](auto sc, int s, uint32_t ph, bool dg) {
      constexpr int J = decltype(sc)::value;
      float(&x)[20] = *(J ? &xB : &xA);
      const long long row = J ? rowB : rowA;
      const bool live = row < a.N;
      const long long rrow = live ? row : (long long)a.N - 1;
      const int i = a.first_step - s;
      mbar_spin(bar(J, kBD3), ph);
      P2STAMP(dg, 20 + J * 10 + 6);
      tc_fence_after();
      uint32_t r[20];
      TMEM_LD_X16(tmem + lane_addr + J * 256 + c0, r);
      TMEM_LD_X4(tmem + lane_addr + J * 256 + c0 + 16, (r + 16));
      tmem_wait_ld();
      if (a.refine) {
        const float viol = (a.scores[rrow] < 0.f) ? 1.f : 0.f;
        const float* u0r = a.u0 + rrow * 40 + c0;
        float* orow = a.out + rrow * 40 + c0;
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
        return;
      }
      if (a.mu_out) {
        const float c1 = a.c1[i], c2 = a.c2[i];
        float* mo = a.mu_out + rrow * 40 + c0;
#pragma unroll
        for (int j = 0; j < 20; ++j) {
          const float eps = __uint_as_float(r[j]) + b3s[c0 + j] + x[j];
          if (live) mo[j] = c2 * (x[j] - c1 * eps);
        }
        return;
      }
      const float kd = a.c2[i] * a.c1[i];
#pragma unroll
      for (int j = 0; j < 20; ++j) x[j] = fmaf(-kd, __uint_as_float(r[j]), pre[j]);
      if (s + 1 < S) {
        store_x(sc);
        hand_over(lbar(J, kBX));
      }
      P2STAMP(dg, 20 + J * 10 + 7);
      const int kidx = a.keep - i;
      if (a.iterates && kidx >= 0) {
        // kept iterate: this CTA's 128 x 40 block is contiguous in HBM: staged in H[slot] K-blocks 2..3 (free until the
        // next layer-1 epilogue of this slot), one bulk store, and its shared-memory read is awaited before moving on
        const uint32_t stg = sbase + k2OffH + J * k2HBytes + k2OffStage;
        float* o = reinterpret_cast<float*>(smem + k2OffH + J * k2HBytes + k2OffStage) + m * 40 + c0;
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
          const long long r0 = row - m;  // first row of this CTA's half
          const long long left = (long long)a.N - r0;
          const int rows = (int)(left < 0 ? 0 : (left < 128 ? left : 128));
          if (rows > 0) {
            float* dst = a.iterates + ((size_t)kidx * a.N + r0) * 40;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(stg), "r"(rows * 160)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      }
    };
    auto load_tile = [&](auto sc, int tile) {
      constexpr int J = decltype(sc)::value;
      float(&x)[20] = *(J ? &xB : &xA);
      uint32_t(&pc)[4] = *(J ? &pcB : &pcA);
      long long row, rrow;
      bool live;
      tile_rows(tile, row, rrow, live);
      (J ? rowB : rowA) = row;
      const int scene0 = (int)(((long long)tile * k2TileM) / a.rows_per_scene);
      (J ? clsB : clsA) = (int)(rrow / a.rows_per_scene) - scene0;
      const float* xr = a.xin + rrow * PSTL_XIN_LD;
#pragma unroll
      for (int j = 0; j < 20; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c0 + j);
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + 40 + j);
        pc[j / 2] = pack_bf16(v.x, v.y);
        pc[j / 2 + 1] = pack_bf16(v.z, v.w);
      }
      store_x(sc);
      hand_over(lbar(J, kBX));
    };
    auto save_tile = [&](auto sc) {
      constexpr int J = decltype(sc)::value;
      float(&x)[20] = *(J ? &xB : &xA);
      const long long row = J ? rowB : rowA;
      if (row < a.N && !a.refine && !a.mu_out) {
        float* xw = a.xin + row * PSTL_XIN_LD;
#pragma unroll
        for (int j = 0; j < 20; j += 4) *reinterpret_cast<float4*>(xw + c0 + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
      }
    };

    for (int r = 0; r < n_rounds; ++r) {
      const int tA = r * 2 * n_pairs + pair, tB = tA + n_pairs;
      if (tA >= n_tiles) break;
      const bool vB = tB < n_tiles;
      load_tile(IntC<0>{}, tA);
      if (vB) load_tile(IntC<1>{}, tB);
      for (int s = 0; s < S; ++s) {
        const uint32_t ph = (uint32_t)(r * S + s) & 1u;
        const bool dg = blockIdx.x == 0 && r == 0 && s == 2 && threadIdx.x == 0;
        e12(IntC<0>{}, 0, ph, dg);
        if (vB && s > 0) e3(IntC<1>{}, s - 1, ph ^ 1u, dg);
        prepare(IntC<0>{}, s);
        P2STAMP(dg, 28);
        e12(IntC<0>{}, 1, ph, dg);
        if (vB) e12(IntC<1>{}, 0, ph, dg);
        e3(IntC<0>{}, s, ph, dg);
        if (vB) {
          prepare(IntC<1>{}, s);
          P2STAMP(dg, 38);
          e12(IntC<1>{}, 1, ph, dg);
        }
      }
      if (vB) e3(IntC<1>{}, S - 1, (uint32_t)(r * S + S - 1) & 1u, false);
      save_tile(IntC<0>{});
      if (vB) save_tile(IntC<1>{});
    }
    if (warp == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer's MMAs / arrivals may still touch this CTA
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace
