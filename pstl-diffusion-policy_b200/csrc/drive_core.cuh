// drive_core.cuh — per-pose driving predicates and the Euler unicycle, forward values and the
// partial derivatives the fused backward needs.  Operation order mirrors the reference's fp32
// PyTorch expressions (this translation unit is compiled with -fmad=false so products and sums
// round separately, as they do upstream).
#pragma once
#include "stl_core.cuh"

struct PstlPose {
  float x, y, th, v;
};

// nusc_train.py:29-49: x += v cos(th) dt, y += v sin(th) dt, th += w dt, v += a dt (old th, v)
PSTL_HD PstlPose pstl_unicycle_step(const PstlPose& s, float w, float a, float dt, float c, float sn) {
  PstlPose n;
  n.x = s.x + (s.v * c) * dt;
  n.y = s.y + (s.v * sn) * dt;
  n.th = s.th + w * dt;
  n.v = s.v + a * dt;
  return n;
}

// torch.linspace(0,1,n)[k] in fp32 (symmetric formula of ATen's linspace kernel)
PSTL_HD float pstl_linspace01(int k, int n) {
  if (n == 1) return 0.f;
  const float step = 1.0f / (float)(n - 1);
  return (k < n / 2) ? step * (float)k : 1.0f - step * (float)(n - k - 1);
}

#define PSTL_LANE_CLIP 1    // bits of the lane flag word carried in pstl_spec_params.clip_dist
#define PSTL_LANE_INLINE 2
#define PSTL_NL 4  // --refined_nL (the only value built; checked on the host side)

// utils.py:465-497 (num_W = 1): body-axis circle centres of a car and the common radius.
struct PstlCircles {
  float cx[PSTL_NL], cy[PSTL_NL];
  float q[PSTL_NL];  // body-x offsets (needed for d centre / d heading)
  float q_y;         // body-y offset (0 for car-shaped boxes)
  float r;
};

PSTL_HD void pstl_car_circles(float x, float y, float c, float sn, float L, float W, PstlCircles& out) {
  const float r_l = L / (float)PSTL_NL / 2.f;
  const float r_w = W / 1.f / 2.f;
  const float r = fminf(fmaxf(r_l, r_w), W / 2.f);
  const float x1 = L / 2.f, x2 = -L / 2.f;
  const float y2 = W / 2.f, y3 = -(W / 2.f);
  const float ys = (y3 + r) * (1.f - 0.f) + (y2 - r) * 0.f;  // linspace(0,1,1) = [0]
  out.r = r;
  out.q_y = ys;
#pragma unroll
  for (int k = 0; k < PSTL_NL; ++k) {
    const float al = pstl_linspace01(k, PSTL_NL);
    const float xs = (x2 + r) * (1.f - al) + (x1 - r) * al;
    out.q[k] = xs;
    out.cx[k] = xs * c - ys * sn + x;
    out.cy[k] = xs * sn + ys * c + y;
  }
}

// Neighbour at one time step as the scene accessors hand it out.
struct PstlNei {
  float cx[PSTL_NL], cy[PSTL_NL];
  float r, valid;
};

// utils.py:499-510 + nusc_train.py:142-148 for ONE neighbour: clipped clearance term and, when
// grad != nullptr, d term / d (ego x, y, th).
PSTL_HD float pstl_pair_clearance(const PstlCircles& e, float ec, float es, const PstlNei& n, float* grad /*3 or null*/) {
  float best = INFINITY;
  int bi = 0, bj = 0;
#pragma unroll
  for (int i = 0; i < PSTL_NL; ++i)
#pragma unroll
    for (int j = 0; j < PSTL_NL; ++j) {
      const float dx = e.cx[i] - n.cx[j], dy = e.cy[i] - n.cy[j];
      const float d2 = fmaf(dx, dx, dy * dy);  // one rounding less than the unfused sum; used by every kernel alike
      if (d2 < best) { best = d2; bi = i; bj = j; }
    }
  const float mind = sqrtf(best);  // sqrt is monotone: min of norms == norm at the min square
  const float car = mind - e.r - n.r;
  const float clipped = fminf(fmaxf(car, -5.f), 20.f);
  const float term = clipped * n.valid + (1.f - n.valid) * 100.f;
  if (grad) {
    float gx = 0.f, gy = 0.f, gth = 0.f;
    if (car >= -5.f && car <= 20.f) {
      float ex = 0.f, ey = 0.f, nx = 0.f, ny = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < PSTL_NL; ++i) {
        if (i == bi) { ex = e.cx[i]; ey = e.cy[i]; q = e.q[i]; }
        if (i == bj) { nx = n.cx[i]; ny = n.cy[i]; }
      }
      const float ux = (ex - nx) / mind, uy = (ey - ny) / mind;  // NaN when coincident, as torch.norm's backward
      gx = ux * n.valid;
      gy = uy * n.valid;
      // centre = (x + q c - qy s, y + q s + qy c): d/dth = (-q s - qy c, q c - qy s)
      gth = (ux * (-q * es - e.q_y * ec) + uy * (q * ec - e.q_y * es)) * n.valid;
    }
    grad[0] = gx; grad[1] = gy; grad[2] = gth;
  }
  return term;
}

// Farthest a car's circles reach from its centre: centres at body offsets |q| <= |L/2 - r| with r = W/2
// (utils.py:465-497), so reach = |L/2 - W/2| + W/2 = max(L/2, W - L/2) (L/2 for car-shaped boxes, L >= W).
PSTL_HD float pstl_car_reach(float L, float W) { return fmaxf(L / 2.f, W - L / 2.f); }

// Exact cull for a neighbour with valid == 1:  car_dist >= |C_ego - C_nei| - reach_ego - reach_nei
// (pstl_car_reach).  If that bound (less a rounding margin) is not below min(best, 20) the neighbour cannot
// lower the running minimum and its clipped term is exactly 20 when the bound exceeds 20.
PSTL_HD bool pstl_cull_neighbour(float dx, float dy, float ego_half_len, float reach, float best) {
  const float R = fminf(best, 20.f) + ego_half_len + reach + 1e-3f;
  return R > 0.f && (dx * dx + dy * dy) >= R * R;
}

// The segment search only ranks d_j + d_{j+1}: a 1-ulp square root (MUFU.SQRT) changes the arg-min only
// between sums that differ by rounding; every distance that reaches the output uses the IEEE sqrtf.
// .ftz: without it every call carries a denormal guard (FSETP + two predicated FMULs around the MUFU, 3 of the 13
// instructions per lane point); a squared distance below 1.2e-38 m^2 ranks as 0 either way.
PSTL_HD float pstl_sqrt_search(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}

// nusc_api.py:693-735, after the segment search: signed lateral distance (triangle area / base) and
// heading error of pose p to the segment (x2,y2,th2)-(x3,y3).
// flags: PSTL_LANE_CLIP (--clip_dist, :732-733) | PSTL_LANE_INLINE (--inline end-caps, :716-724: a pose behind the
// first / ahead of the last segment takes the clamped point distance to that end, signed like the line distance);
// at_first / at_last: the arg-min segment is the polyline's first / last one (min_idx == 0 / nseg-2).
// part (3 floats or null): d dist/d px, d dist/d py, d ang/d pth.
PSTL_HD void pstl_lane_finish(float px, float py, float pth, float x2, float y2, float th2, float x3, float y3,
                              int flags, bool at_first, bool at_last, float& dist, float& ang, float* part) {
  const float area = px * (y2 - y3) + x2 * (y3 - py) + x3 * (py - y2);
  const float bx = x2 - x3, by = y2 - y3;
  const float base = sqrtf(bx * bx + by * by);
  const float ex = px - x2, ey = py - y2;
  const float q = ex * ex + ey * ey;
  const float l2 = sqrtf(fmaxf(q, 1e-3f));
  const float ok = (base != 0.f) ? 1.f : 0.f;
  const float den = fmaxf(base, 1e-7f);
  float d0 = ok * area / den + (1.f - ok) * l2;
  float gdx = 0.f, gdy = 0.f;
  if (part) {
    gdx = ok * (y2 - y3) / den;
    gdy = ok * (x3 - x2) / den;
    if (ok == 0.f && q >= 1e-3f) { gdx += ex / l2; gdy += ey / l2; }
  }
  if (flags & PSTL_LANE_INLINE) {
    const float fx = px - x3, fy = py - y3;
    const float qb = fx * fx + fy * fy;
    const float l2b = sqrtf(fmaxf(qb, 1e-3f));
    const bool behind = at_first && (ex * (x3 - x2) + ey * (y3 - y2) <= 0.f);
    const bool ahead = at_last && (fx * (x2 - x3) + fy * (y2 - y3) <= 0.f);
    if (behind || ahead) {
      const float sg = (d0 > 0.f) ? 1.f : (d0 < 0.f) ? -1.f : 0.f;  // torch.sign: no gradient
      const float fb = behind ? 1.f : 0.f, fa = ahead ? 1.f : 0.f;
      d0 = fb * l2 * sg + fa * l2b * sg;
      if (part) {
        gdx = 0.f; gdy = 0.f;
        if (behind && q >= 1e-3f) { gdx += sg * (ex / l2); gdy += sg * (ey / l2); }
        if (ahead && qb >= 1e-3f) { gdx += sg * (fx / l2b); gdy += sg * (fy / l2b); }
      }
    }
  }
  if (flags & PSTL_LANE_CLIP) {
    if (d0 < -5.f || d0 > 5.f) { gdx = 0.f; gdy = 0.f; }
    d0 = fminf(fmaxf(d0, -5.f), 5.f);
  }
  const float u = th2 - pth;
  dist = d0;
  ang = 1.f - cosf(u);
  if (part) {
    part[0] = gdx;
    part[1] = gdy;
    part[2] = -sinf(u);
  }
}

// nusc_api.py:693-735: closest segment by arg-min of d_j + d_{j+1} (first minimum), then pstl_lane_finish.
template <class LaneAcc>
PSTL_HD void pstl_lane_pred(float px, float py, float pth, const LaneAcc& lane, int nseg, int clip_dist,
                            float& dist, float& ang, float* part) {  // clip_dist: PSTL_LANE_* flags
  float prev = 0.f;
  float bestv = INFINITY;
  int bi = 0;
  for (int j = 0; j < nseg; ++j) {
    const float dx = px - lane(j, 0), dy = py - lane(j, 1);
    const float d = pstl_sqrt_search(dx * dx + dy * dy);
    if (j > 0) {
      const float sum = prev + d;
      if (sum < bestv) { bestv = sum; bi = j - 1; }
    }
    prev = d;
  }
  pstl_lane_finish(px, py, pth, lane(bi, 0), lane(bi, 1), lane(bi, 2), lane(bi + 1, 0), lane(bi + 1, 1), clip_dist,
                   bi == 0, bi == nseg - 2, dist, ang, part);
}

// decode helpers for PSTL_OP_PRED
PSTL_HD float pstl_pred_den(int den, const float* p) {
  switch (den) {
    case PSTL_DEN_THMAX: return p[5];
    case PSTL_DEN_VFACTOR: return fmaxf(p[1] - p[0], 0.3f);
    case PSTL_DEN_DFACTOR: return fmaxf((p[3] - p[2]) * 5.f, 0.3f);
    case PSTL_DEN_SFACTOR: return fmaxf(p[4], 0.3f);
    default: return 1.f;
  }
}
