// score_dense.cuh — time-parallel forward scorer for the reference's DENSE per-row layout
// (compute_stl_dense on per-row tensors, reference nusc_train.py:318-345: BASELINE configs 1 and 5).
//
// In that layout every trajectory owns its copy of the scene tensors (neighbours (K,T,7) = 28 K T bytes, three lanes),
// so the call is HBM-bound: 5,376 B per trajectory at T = 20, K = 8.  A thread-per-row kernel reads the
// 28-byte neighbour records with a stride of K T 28 B between lanes — one 32-byte sector per useful 28 B and no
// coalescing — and holds too little data in flight (650-770 GB/s measured).  Here:
//   * one THREAD per (row, time step): the pre-rolled pose of step t, the lane search and the K clearances of step t
//     are independent across t (the rollout is given), so a block of R rows x T steps works on R T poses at once;
//   * the neighbour block of the R rows — R contiguous runs of KC T 28 bytes — arrives by cp.async.bulk (TMA 1-D) into
//     shared memory, completion on an mbarrier; for long horizons / many neighbours the K axis is walked in chunks of
//     KC neighbours with two buffers, the copy of chunk c+1 in flight while chunk c is consumed.  Lane (r, t) reads
//     record (k, t) at word offset 7 t: a stride-7 access, conflict-free over the 32 banks;
//   * the temporal operators then run per (row, term): the X(t) column of every term sits in shared memory, one thread
//     per (row, term) folds it with the SAME online log-sum-exp recurrence, in the same order, as the streaming scorer
//     (score_stream.cuh), and one thread per row combines the terms — so the scores are bit-identical to that kernel's.
// The 16-pair circle clearances are bracketed by cheap centre-distance bounds first; only the (step, neighbour) pairs that
// can still decide the score are evaluated, compacted into a block-wide queue so that no lane idles on a culled pair.
// Only the row's own-mode lane polyline is read (the reference evaluates all three formulas and masks; equal whenever
// the unused formulas are finite).
#pragma once
#include "score_stream.cuh"

#if defined(__CUDACC__)
#define PSTL_DENSE_MAX_THREADS 256

__device__ __forceinline__ uint32_t dtp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dtp_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void dtp_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dtp_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void dtp_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

struct DenseTpCfg {
  int R;        // rows per block
  int KC;       // neighbours per chunk
  int nbuf;     // 1 (single chunk) or 2
  int chunk_floats;  // R-row stride inside a buffer: KC * T * 7
  int bulk_small;    // the per-row inputs (poses, lanes, pSTL parameters, modes) of a full block can travel as bulk copies
};

// order-preserving map float -> unsigned, so a shared-memory atomicMin works on signed floats
__device__ __forceinline__ unsigned dtp_enc(float f) {
  const unsigned b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float dtp_dec(unsigned u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

// shared-memory carve-up (32-bit words, every section 16-byte aligned)
__host__ __device__ __forceinline__ size_t dtp_align4(size_t n) { return (n + 3) & ~(size_t)3; }
struct DenseTpSmem {
  size_t nei, lanes, xcol, rowc, pose, exact, m1, queue, pin, misc, total;
};
__host__ __device__ __forceinline__ DenseTpSmem dtp_smem_layout(const DenseTpCfg& d, int T, int nseg) {
  DenseTpSmem L;
  size_t o = 0;
  L.nei = o;   o += dtp_align4((size_t)d.nbuf * d.R * d.chunk_floats);      // neighbour records, as in HBM
  L.lanes = o; o += dtp_align4((size_t)3 * d.R * nseg * 3);                  // the three lane polylines per row
  const size_t xq = (size_t)PSTL_MAX_TERMS * d.R * T > (size_t)d.R * d.KC * T ? (size_t)PSTL_MAX_TERMS * d.R * T
                                                                               : (size_t)d.R * d.KC * T;
  L.xcol = o;  o += dtp_align4(xq);                                          // X(t) per [term][row][t]; before that: the work queue
  L.rowc = o;  o += dtp_align4((size_t)d.R * (2 * PSTL_MAX_TERMS + 8));      // per row: g2[8], qc[8], p[6], margin, mode
  L.pose = o;  o += (size_t)4 * d.R * T;                                     // (x, y, cos, sin) per (row, t)
  L.exact = o; o += dtp_align4((size_t)d.R * T);                             // exact clearance minima, ordered-uint
  L.m1 = o;    o += dtp_align4(d.R);                                         // upper bound of the row's smallest clearance
  L.queue = L.xcol;                                                          // (row*T + t) << 16 | k  (dead before X(t) is written)
  L.pin = o;   o += dtp_align4((size_t)d.R * 6) + dtp_align4(d.R);           // landing zone of the pSTL parameters and modes
  L.misc = o;  o += 8;                                                       // 3 mbarriers, queue counter
  L.total = o;
  return L;
}
#define DTP_ROWC (2 * PSTL_MAX_TERMS + 8)

// TT / NSEG: compile-time horizon and lane-point count of the reference's defaults (index arithmetic and the segment
// search unroll), 0 = read them from the arguments
template <int TT, int NSEG>
__global__ void __launch_bounds__(PSTL_DENSE_MAX_THREADS)
k_score_dense_tp(const __grid_constant__ ScoreArgs a, const __grid_constant__ StreamPlans sp, const __grid_constant__ DenseTpCfg d) {
  extern __shared__ float4 sm4[];
  PstlEvalCfg c = a.cfg;
  if (TT) c.T = TT;
  if (NSEG) c.nseg = NSEG;
  const int T = TT ? TT : c.T, K = c.K, R = d.R;
  const DenseTpSmem SL = dtp_smem_layout(d, T, c.nseg);
  float* sm = reinterpret_cast<float*>(sm4);
  float* nbuf = sm + SL.nei;
  float* lanes = sm + SL.lanes;
  float* xcol = sm + SL.xcol;
  float* rowc = sm + SL.rowc;
  float4* s_pose = reinterpret_cast<float4*>(sm + SL.pose);
  unsigned* s_exact = reinterpret_cast<unsigned*>(sm + SL.exact);
  unsigned* s_m1 = reinterpret_cast<unsigned*>(sm + SL.m1);
  unsigned* s_queue = reinterpret_cast<unsigned*>(sm + SL.queue);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SL.misc);
  int* s_count = reinterpret_cast<int*>(sm + SL.misc + 6);
  const uint32_t bar0 = dtp_smem_u32(bars), bar_in = bar0 + 16u;
  float* s_pin = sm + SL.pin;

  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * R;
  const int rows = (a.N - n0 < R) ? a.N - n0 : R;
  const int nchunks = (K + d.KC - 1) / d.KC;

  auto issue_chunk = [&](int ch) {  // called by warp 0: lane r copies row r's run of the chunk
    const int k0 = ch * d.KC;
    const int kc = (K - k0 < d.KC) ? K - k0 : d.KC;
    const uint32_t bytes = (uint32_t)kc * T * 28u;
    const int b = ch % d.nbuf;
    const uint32_t bar = bar0 + 8u * b;
    if (tid == 0) dtp_mbar_expect_tx(bar, bytes * (uint32_t)rows);
    __syncwarp();
    for (int r = tid; r < rows; r += 32) {
      const float* src = a.neighbors + (((size_t)(n0 + r) * K + k0) * T) * 7;
      dtp_bulk_g2s(dtp_smem_u32(nbuf + ((size_t)b * R + r) * d.chunk_floats), src, bytes, bar);
    }
  };

  const int per = c.nseg * 3;
  const bool bulk_in = d.bulk_small && rows == R;
  if (tid < 32) {  // warp 0: barriers, then every copy the block needs (the other warps meet them after the block barrier)
    if (tid == 0) {
      for (int b = 0; b < d.nbuf; ++b) dtp_mbar_init(bar0 + 8u * b, 1);
      dtp_mbar_init(bar_in, 1);
      *s_count = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (bulk_in) {
      // poses (land where the (x, y, cos, sin) records will live), the three lane polylines, pSTL parameters, modes
      if (tid == 0) dtp_mbar_expect_tx(bar_in, (uint32_t)(R * T * 16 + 3 * R * per * 4 + R * 24 + R * 4));
      __syncwarp();
      if (tid == 0) dtp_bulk_g2s(dtp_smem_u32(s_pose), a.ego + (size_t)n0 * T * 4, (uint32_t)(R * T * 16), bar_in);
      if (tid >= 1 && tid <= 3)
        dtp_bulk_g2s(dtp_smem_u32(lanes + (size_t)(tid - 1) * R * per), a.lanes[tid - 1] + (size_t)n0 * per, (uint32_t)(R * per * 4), bar_in);
      if (tid == 4) dtp_bulk_g2s(dtp_smem_u32(s_pin), a.stlp + (size_t)n0 * 6, (uint32_t)(R * 24), bar_in);
      if (tid == 5) dtp_bulk_g2s(dtp_smem_u32(s_pin + dtp_align4((size_t)R * 6)), a.mode + n0, (uint32_t)(R * 4), bar_in);
    }
    issue_chunk(0);
  }
  const int r = tid / T, t = tid - r * T;
  const bool live = tid < rows * T;
  if (tid < R * T) s_exact[tid] = 0xFFFFFFFFu;
  if (tid < R) s_m1[tid] = 0xFFFFFFFFu;
  PstlPose s{0.f, 0.f, 0.f, 0.f};
  if (!bulk_in) {
    // ragged last block / unaligned shapes: plain loads, all issued before the first use
    if (live) {
      const float* e = a.ego + ((size_t)(n0 + r) * T + t) * a.ego_stride;
      s = PstlPose{e[0], e[1], e[2], e[3]};
    }
    for (int l = 0; l < 3; ++l)  // rows of one lane tensor are contiguous: coalesced
      for (int j = tid; j < rows * per; j += blockDim.x) lanes[(size_t)l * R * per + j] = __ldg(a.lanes[l] + (size_t)n0 * per + j);
    for (int e = tid; e < rows * 6; e += blockDim.x) s_pin[e] = __ldg(a.stlp + (size_t)n0 * 6 + e);
    for (int e = tid; e < rows; e += blockDim.x) s_pin[dtp_align4((size_t)R * 6) + e] = __ldg(a.mode + n0 + e);
  }
  __syncthreads();  // barriers initialised, plain-load staging visible
  if (bulk_in) {
    dtp_mbar_wait(bar_in, 0);
    if (live) {
      const float4 q = s_pose[tid];
      s = PstlPose{q.x, q.y, q.z, q.w};
    }
  }
  float sn = 0.f, cs = 1.f;
  if (live) sincosf(s.th, &sn, &cs);
  if (tid < R * T) s_pose[tid] = make_float4(s.x, s.y, cs, sn);  // in place: each lane rewrites the record it just read
  const float* s_mode = s_pin + dtp_align4((size_t)R * 6);
  int m = 4;
  if (live) {
    const float md = s_mode[r];
    m = (md == 0.f) ? 0 : (md == 1.f) ? 1 : (md == 2.f) ? 2 : (md == 3.f) ? 3 : 4;
  }
  const bool work = live && m < 3;
  const PstlPlan& pl = sp.p[work ? m : 0];
  const float* rc = rowc + (size_t)(live ? r : 0) * DTP_ROWC;
  const float* prow = s_pin + (size_t)(live ? r : 0) * 6;

  // ---- per-(row, term) constants, one thread each: the term's factor g2 and constant q/den; the row's cull margin and
  //      mode.  Read after the next block barrier (pass B, X(t), the reduction). ----
  if (tid < rows * PSTL_MAX_TERMS) {
    const int rr = tid / PSTL_MAX_TERMS, k = tid - rr * PSTL_MAX_TERMS;
    float* rcw = rowc + (size_t)rr * DTP_ROWC;
    const float md = s_mode[rr];
    const int mm = (md == 0.f) ? 0 : (md == 1.f) ? 1 : (md == 2.f) ? 2 : (md == 3.f) ? 3 : 4;
    float g2 = 0.f, qc = 0.f;
    if (mm < 3 && k < sp.p[mm].n_terms) {
      const PstlTerm& tm = sp.p[mm].terms[k];
      const float* pr = s_pin + (size_t)rr * 6;
      const float k2 = c.tau * 1.4426950408889634f;
      if (tm.pair == 0) {  // as pstl_stream_init / pstl_stream_top
        float g = tm.a.sb * k2;
        if (tm.a.den != PSTL_DEN_ONE) g = g / pstl_pred_den(tm.a.den, pr);
        g2 = (tm.inner == 0) ? g * (float)tm.outer : g;
        float q = tm.a.sp * pr[tm.a.pid];
        if (tm.a.den != PSTL_DEN_ONE) q = q / pstl_pred_den(tm.a.den, pr);
        qc = q * k2;
      }
      if (k == 0) rcw[2 * PSTL_MAX_TERMS + 6] = (sp.p[mm].nei_term == 0) ? -36.f * (1.f / g2) : INFINITY;  // 36 / |g2|, metres
    } else if (k == 0) {
      rcw[2 * PSTL_MAX_TERMS + 6] = INFINITY;
    }
    if (k == 1) rcw[2 * PSTL_MAX_TERMS + 7] = __int_as_float(mm);
    rcw[k] = g2;
    rcw[PSTL_MAX_TERMS + k] = qc;
  }

  // ---- lane predicate (nusc_api.py:693-735) ----
  float dl = 0.f, th = 0.f;
  if (work && t < pl.need_lane) {
    const float* ln = lanes + ((size_t)pl.lane * R + r) * c.nseg * 3;
    float dx = s.x - ln[0], dy = s.y - ln[1];
    float prev = pstl_sqrt_search(fmaf(dx, dx, dy * dy));
    float bestv = INFINITY;
    int bi = 0;
    for (int j = 1; j < c.nseg; ++j) {
      dx = s.x - ln[j * 3]; dy = s.y - ln[j * 3 + 1];
      const float dj = pstl_sqrt_search(fmaf(dx, dx, dy * dy));
      const float sum = prev + dj;
      if (sum < bestv) { bestv = sum; bi = j - 1; }
      prev = dj;
    }
    const float* p2 = ln + bi * 3;
    pstl_lane_finish(s.x, s.y, s.th, p2[0], p2[1], p2[2], p2[3], p2[4], c.clip_dist, bi == 0, bi == c.nseg - 2, dl, th,
                     nullptr);
  }

  // ---- neighbour clearance (utils.py:465-526, nusc_train.py:142-148), K walked in chunks ----
  // The exact clearance of one (step, neighbour) pair — 16 circle-pair distances — costs ~270 instructions, and which
  // pairs need it is data dependent, so lanes would diverge.  Instead every lane first brackets every pair with
  // two cheap centre-distance bounds  L <= term <= U  (term = clip(car_dist, -5, 20), or 100 for an invalid slot):
  //     car_dist >= |C_e - C_n| - reach_e - reach_n          (pstl_car_reach)
  //     car_dist <= |C_e - C_n| + |q_e| + |q_n| - r_e - r_n   (the two inner circles, body offsets q = (L/2 - r)/3)
  // (each widened by 2e-3 m, far above the rounding of either side) and only the pairs that can still matter go to a
  // block-wide work queue that ALL threads drain densely:
  //   (a) L < U_t = min_k U(k, t): otherwise another neighbour is at least as close at this step;
  //   (b) L < M + 36/(tau log2 e): M = min over the row's steps of U_t bounds the row's smallest clearance from above; when
  //       the clearance only feeds the soft-min  G_t(nei_t - q)  (pl.nei_term == 0) a step whose clearance exceeds the
  //       smallest one by that margin adds less than 2^-36 to a sum >= 1 — nothing in fp32 (the streaming scorer's
  //       value-aware bound, made order-free).
  // A pair that is not evaluated contributes its lower bound L, which by (a)/(b) is either >= the step's true value or
  // irrelevant to the score; equality cases (both bounds clipped to 20, invalid slots at 100) are exact.  L is parked in
  // the record's speed field (never read by the scorer) between the two passes.
  const bool nei_gated = pl.nei_term == 0;
  const bool need_nei = work && t < pl.need_nei && (!nei_gated || (t >= pl.terms[0].lo && t < pl.terms[0].hi));
  float rec_min = INFINITY;  // smallest lower bound among this step's unevaluated pairs
  float U_t = INFINITY;
  const float ego_reach = pstl_car_reach(c.ego_L, c.ego_W);
  const float ego_r = fminf(fmaxf(c.ego_L / (float)PSTL_NL / 2.f, c.ego_W / 2.f), c.ego_W / 2.f);
  const float lo_off = ego_reach + 2e-3f;
  const float up_off = fabsf(c.ego_L / 2.f - ego_r) * 0.33334f + 2e-3f - ego_r;
  for (int ch = 0; ch < nchunks; ++ch) {
    // two buffers: the copy of chunk ch+1 starts now (its buffer was released by the last barrier below); one buffer:
    // chunk ch itself starts now (smaller footprint, more blocks per SM — other blocks cover the exposed latency)
    if (d.nbuf > 1 && ch + 1 < nchunks && tid < 32) issue_chunk(ch + 1);
    if (d.nbuf == 1 && ch > 0 && tid < 32) issue_chunk(ch);
    dtp_mbar_wait(bar0 + 8u * (ch % d.nbuf), (uint32_t)((ch / d.nbuf) & 1));
    const int k0 = ch * d.KC;
    const int kc = (K - k0 < d.KC) ? K - k0 : d.KC;
    float* buf = nbuf + (size_t)(ch % d.nbuf) * R * d.chunk_floats;
    float* rec0 = buf + (size_t)r * d.chunk_floats + (size_t)t * 7;
    // pass A: both bounds of this chunk's pairs
    if (need_nei) {
      float* rec = rec0;
      for (int k = 0; k < kc; ++k, rec += (size_t)T * 7) {
        const float valid = rec[0];
        const float dx = s.x - rec[1], dy = s.y - rec[2];
        const float nL = rec[5], nW = rec[6];
        const float dc = pstl_sqrt_search(fmaf(dx, dx, dy * dy));
        const float hl = nL * 0.5f, nr = nW * 0.5f;  // r = min(max(L/8, W/2), W/2) = W/2
        float Lk = fminf(fmaxf(dc - (lo_off + fmaxf(hl, nW - hl)), -5.f), 20.f);
        float Uk = fminf(fmaxf(dc + (up_off + fabsf(hl - nr) * 0.33334f) - nr, -5.f), 20.f);
        if (valid != 1.f) {  // 0: clip(d) * 0 + (1 - 0) * 100; fractional validity: always evaluated
          Lk = (valid == 0.f) ? 100.f : -INFINITY;
          Uk = (valid == 0.f) ? 100.f : INFINITY;
        }
        rec[4] = Lk;
        U_t = fminf(U_t, Uk);
      }
      if (nei_gated) atomicMin(&s_m1[r], dtp_enc(U_t));
    }
    __syncthreads();
    // pass B: queue the pairs that can still matter
    if (need_nei) {
      float thr = fminf(U_t, dtp_dec(s_exact[tid]));
      if (nei_gated) thr = fminf(thr, dtp_dec(s_m1[r]) + rc[2 * PSTL_MAX_TERMS + 6]);
      const float* rec = rec0;
      for (int k = 0; k < kc; ++k, rec += (size_t)T * 7) {
        const float Lk = rec[4];
        if (Lk < thr) s_queue[atomicAdd(s_count, 1)] = ((unsigned)tid << 16) | (unsigned)k;
        else rec_min = fminf(rec_min, Lk);
      }
    }
    __syncthreads();
    // drain: every thread takes queued pairs, whoever queued them
    const int n_items = *s_count;
    for (int i = tid; i < n_items; i += blockDim.x) {
      const unsigned it = s_queue[i];
      const int lt = (int)(it >> 16), k = (int)(it & 0xffffu);
      const int lr = lt / T, ltt = lt - lr * T;
      const float4 ps = s_pose[lt];
      const float* rec = buf + (size_t)lr * d.chunk_floats + ((size_t)k * T + ltt) * 7;
      PstlCircles ec, cc;
      pstl_car_circles(ps.x, ps.y, ps.z, ps.w, c.ego_L, c.ego_W, ec);
      const float ncs = cosf(rec[3]), nsn = sinf(rec[3]);  // as the streaming scorer's accessors evaluate them
      pstl_car_circles(rec[1], rec[2], ncs, nsn, rec[5], rec[6], cc);
      PstlNei nb;
#pragma unroll
      for (int q = 0; q < PSTL_NL; ++q) { nb.cx[q] = cc.cx[q]; nb.cy[q] = cc.cy[q]; }
      nb.r = cc.r;
      nb.valid = rec[0];
      const float term = pstl_pair_clearance(ec, ps.z, ps.w, nb, nullptr);
      atomicMin(&s_exact[lt], dtp_enc(term));
    }
    __syncthreads();  // exact minima visible; the chunk's buffer and the queue are free again
    if (tid == 0) *s_count = 0;
  }
  const float nei = need_nei ? fminf(dtp_dec(s_exact[tid]), rec_min) : 0.f;

  // ---- X(t) of every term, exactly as pstl_stream_steps evaluates it ----
  if (work) {
    const float k2 = c.tau * 1.4426950408889634f;
    const float* p = prow;
#pragma unroll
    for (int k = 0; k < PSTL_MAX_TERMS; ++k) {
      if (k < pl.n_terms) {
        const PstlTerm& tm = pl.terms[k];
        if (t < pl.need_pose && (tm.inner != 0 || (t >= tm.lo && t < tm.hi))) {
          float x2;
          if (tm.pair == 0) {
            const float base = (tm.a.c == 0) ? s.v : (tm.a.c == 1) ? dl : (tm.a.c == 2) ? th : nei;
            x2 = base * rc[k];
          } else {
            const float g = (float)tm.pair * k2;
            const float xa = pstl_plan_leaf(tm.a, s.v, dl, th, nei, p) * g, xb = pstl_plan_leaf(tm.b, s.v, dl, th, nei, p) * g;
            x2 = (pstl_lg2(1.f + pstl_ex2(-fabsf(xa - xb))) + fmaxf(xa, xb)) * (float)tm.pair;
            if (tm.inner == 0) x2 = x2 * (float)tm.outer;
          }
          xcol[((size_t)k * R + r) * T + t] = x2;
        }
      }
    }
  }
  __syncthreads();

  // ---- one thread per (row, term): the temporal reduction, same recurrence and order as the streaming scorer; the 8 term
  //      lanes of a row sit side by side in one warp, so the ListAnd over the terms (pstl_stream_top) is a shuffle loop ----
  if (tid < ((R * PSTL_MAX_TERMS + 31) & ~31)) {
    const int rr = tid / PSTL_MAX_TERMS, k = tid - rr * PSTL_MAX_TERMS;
    const bool rlive = rr < rows;
    const float* rcr = rowc + (size_t)(rlive ? rr : 0) * DTP_ROWC;
    const int mm = rlive ? __float_as_int(rcr[2 * PSTL_MAX_TERMS + 7]) : 4;
    const PstlPlan& plr = sp.p[mm < 3 ? mm : 0];
    const bool tlive = mm < 3 && k < plr.n_terms;
    float y2 = 0.f;
    bool empty = false;
    if (tlive) {
      const PstlTerm& tm = plr.terms[k];
      const float* x = xcol + ((size_t)k * R + rr) * T;
      float am = PSTL_LSE2_INIT, as = 0.f;
      if (tm.inner == 0) {
        const int hi = tm.hi < plr.need_pose ? tm.hi : plr.need_pose;
        for (int tt = tm.lo; tt < hi; ++tt) pstl_lse2_add(am, as, x[tt]);
      } else {
        const float gi = (float)tm.inner, go = (float)(tm.inner * tm.outer);
        float mq = PSTL_LSE2_INIT, sq = 0.f;
        for (int tt = T - 1; tt >= tm.lo; --tt) {
          pstl_lse2_add(mq, sq, x[tt] * gi);
          if (tt < tm.hi) pstl_lse2_add(am, as, (pstl_lg2(sq) + mq) * go);
        }
      }
      empty = tm.hi <= tm.lo;
      y2 = (pstl_lg2(as) + am) * (float)tm.outer;  // term value * tau * log2 e
      if (tm.pair == 0) y2 = y2 + rcr[PSTL_MAX_TERMS + k];
    }
    // top level (stl_d_lib.py:97-112), terms folded in order 0..n-1 as pstl_stream_top does
    const int base_lane = (threadIdx.x & 31) & ~(PSTL_MAX_TERMS - 1);
    float top_m = PSTL_LSE2_INIT, top_s = 0.f, single = 0.f;
    bool any_empty = false;
#pragma unroll
    for (int j = 0; j < PSTL_MAX_TERMS; ++j) {
      const float yj = __shfl_sync(0xffffffffu, y2, base_lane + j);
      const int ej = __shfl_sync(0xffffffffu, (int)empty, base_lane + j);
      if (mm < 3 && j < plr.n_terms) {
        any_empty = any_empty || ej != 0;
        single = yj;
        pstl_lse2_add(top_m, top_s, -yj);
      }
    }
    if (rlive && k == 0) {
      const float back = 0.6931471805599453f / c.tau;  // ln2 / tau
      float sc = (mm == 3) ? 1.0f : 0.0f;  // nusc_train.py:322 outlier score; unknown mode selects nothing (:150-151)
      if (mm < 3) sc = any_empty ? -INFINITY : (!plr.listand ? single * back : -((pstl_lg2(top_s) + top_m) * back));
      const int nn = n0 + rr;
      if (a.scores_all) a.scores_all[nn] = sc;
      if (a.best_score) a.best_score[nn] = sc;
      if (a.best_idx) a.best_idx[nn] = 0;
    }
  }
}
#endif  // __CUDACC__
