// Pair arithmetic of the sweep kernels (kernels_sweep.cu): switching functions,
// one pair term, warp/block reductions.
#pragma once
#include "kernels.cuh"

namespace b200 {

// switching-function classes the sweep is specialised for
enum SwKind {
  K_FIX6 = 0,   // rationalfix6 : the COORDINATION default (NN=6 MM=12 D_0=0)
  K_FIXN,       // other rationalfixN (N/2 in nnf)
  K_RAT_R2,     // rationalFast / rationalSimpleFast : even powers on r^2
  K_RAT_R,      // rational / rationalSimple : needs sqrt
  K_EXP,
  K_GAUSS,
  K_FASTGAUSS,
  K_SMAP,
  K_CUBIC,
  K_TANH,
  K_COS,
  K_NATIVEQ,
  K_GHB,        // GHBFIX pairing (src/colvar/GHBFIX.cpp:186-220): piecewise polynomial, scaled by eta(type,type)
  K_DH,         // DHENERGY pairing (src/colvar/DHEnergy.cpp:130-143): screened Coulomb, scaled by q_i q_j
  K_COUNT
};

static inline int kind_of(int type) {
  switch (type) {
    case 3: return K_FIX6;
    case 0: case 1: case 2: case 4: case 5: return K_FIXN;
    case 7: case 9: return K_RAT_R2;
    case 6: case 8: return K_RAT_R;
    case 10: return K_EXP;
    case 11: return K_GAUSS;
    case 12: return K_FASTGAUSS;
    case 13: return K_SMAP;
    case 14: return K_CUBIC;
    case 15: return K_TANH;
    case 16: return K_COS;
    case 17: return K_NATIVEQ;
    case 32: return K_DH;  // B200COORD_PAIR_DHENERGY
    case 33: return K_GHB;  // B200COORD_PAIR_GHBFIX
    default: return -1;
  }
}

// overloads so that the switching functions below are written once for double and float
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
__device__ __forceinline__ float exp_t(float x) { return expf(x); }
__device__ __forceinline__ double pow_t(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float pow_t(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double tanh_t(double x) { return tanh(x); }
__device__ __forceinline__ float tanh_t(float x) { return tanhf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ void sincospi_t(double x, double* s, double* c) { sincospi(x, s, c); }
__device__ __forceinline__ void sincospi_t(float x, float* s, float* c) { sincospif(x, s, c); }
// rational<>::doRational (SwitchingFunction.cpp:258-283); res/dfn preset to preRes/preDfunc(F)
template <typename T>
__device__ __forceinline__ void rational_generic(bool simplified, T x, T secdev, int N, int M, T& res, T& dfn) {
  if (simplified) {
    const T t = ipow_dev(x, N - 1);
    res = fast_rcp(fma_t(t, x, T(1.0)));
    dfn = -(T)N * t * res * res;
  } else {
    const T hi = T(1.0 + 5.0e10 * 2.220446049250313e-16), lo = T(1.0 - 5.0e10 * 2.220446049250313e-16);
    if (!((x > lo) && (x < hi))) {
      const T tn = ipow_dev(x, N - 1);
      const T tm = ipow_dev(x, M - 1);
      const T num = fma_t(-tn, x, T(1.0));
      const T iden = fast_rcp(fma_t(-tm, x, T(1.0)));
      res = num * iden;
      dfn = (((T)M * res * tm) - ((T)N * tn)) * iden;
    } else {
      const T dx = x - T(1.0);
      res = res + dx * (dfn + T(0.5) * dx * secdev);
      dfn = dfn + dx * secdev;
    }
  }
}

// FP32 sweep: the general quotient (1-x^n)/(1-x^m) and its derivative cancel around x = 1 (1e-4 relative in float
// at |x-1| = 1e-3, and no Taylor window is both wide and accurate enough), so this rarely used kind (MM != 2 NN)
// evaluates its quotient in FP64; the simplified form 1/(1+x^n) has no cancellation and stays in FP32.
__device__ __forceinline__ void rational_generic(bool simplified, float x, float secdev, int N, int M, float& res,
                                                 float& dfn) {
  if (simplified) {
    rational_generic<float>(true, x, secdev, N, M, res, dfn);
  } else {
    double r = (double)res, d = (double)dfn;
    rational_generic<double>(false, (double)x, (double)secdev, N, M, r, d);
    res = (float)r;
    dfn = (float)d;
  }
}

// (s, df=(1/r) ds/dr) of SwitchingFunction::calculateSqr for kind K, stretch/shift and D_MAX applied
// in_dmax / above_d0 (FP32 sweep): 0 / 1 = the side of D_MAX resp. D_0 the pair is on, decided by the caller on the
// FP64 r^2 (the value or the derivative jumps there); -1 = decide here.
template <int K, bool XS = false, typename T = double>
__device__ __forceinline__ void eval_switch(const DevSwitchT<T>& p, T r2, T& s, T& df, int in_dmax = -1,
                                            int above_d0 = -1) {
  s = T(0.0);
  df = T(0.0);
  if (K == K_FIX6 || K == K_FIXN || K == K_RAT_R2) {
    if ((in_dmax >= 0) ? (in_dmax != 0) : (r2 <= p.dmax_2)) {  // fixedRational<N>::calculateSqr :203-215, rational<fast>::calculateSqr :289-303
      const T y = r2 * p.invr0_2;
      T res, d;
      if (K == K_FIX6) {
        const T t = y * y;
        res = fast_rcp(fma_t(t, y, T(1.0)));
        df = (t * res) * (res * p.fix_df);
      } else if (K == K_FIXN) {
        const T t = ipow_dev(y, p.nnf - 1);
        res = fast_rcp(fma_t(t, y, T(1.0)));
        df = (t * res) * (res * p.fix_df);
      } else {
        res = p.preRes;
        d = p.preDfuncF;
        rational_generic(p.type == 9, y, p.preSecDevF, p.nnf, p.mmf, res, d);
        df = d * p.pre_df;
      }
      s = fma_t(res, p.stretch, p.shift);
    }
  } else if (K == K_FASTGAUSS) {  // fastgaussianSwitch::calculateSqr :414-431
    if ((in_dmax >= 0) ? (in_dmax != 0) : (r2 < p.dmax_2)) {
      s = T(1.0);
      if (r2 > T(0.0)) {
        const T res = exp_t(T(-0.5) * r2);
        df = -res * p.stretch;
        s = fma_t(res, p.stretch, p.shift);
      }
    }
  } else {  // baseSwitch::calculateSqr -> calculate(sqrt(r2)) :135-149, :181-183
    T rinv, r;
    if (XS) {  // correctly rounded sqrt: comparisons of r with D_0 / D_MAX match the reference bit for bit
      r = sqrt_t(r2);
      rinv = (r2 > T(0.0)) ? T(1.0) / r : T(0.0);
    } else {
      rinv = (r2 > T(0.0)) ? fast_rsqrt(r2) : T(0.0);
      r = r2 * rinv;
    }
    if (K == K_GHB) {  // GHBFIX::pairing [* eta: caller]; C1 at D_0, at the joint and at D_MAX: no boundary patch needed
      if ((in_dmax >= 0) ? (in_dmax != 0) : !(r2 > p.dmax_2)) {
        const T rdist = r - p.d0;
        s = T(-1.0);
        if (rdist > p.c) {
          s += p.preRes + rdist * (p.preDfunc + p.preSecDev * rdist);
          df = (p.preDfunc + T(2.0) * p.preSecDev * rdist) * rinv;
        } else if (rdist > T(0.0)) {
          s += p.d * (rdist * rdist);
          df = T(2.0) * p.d * rdist * rinv;
        }
      }
    } else if (K == K_DH) {  // DHEnergy::pairing: tmp = exp(-k r)/r * constant/epsilon [* q_i q_j: caller]; dfunc = -(k+1/r) tmp / r
      const T tmp = exp_t(-p.beta * r) * rinv * p.lambda;
      s = tmp;
      df = -(p.beta + rinv) * tmp * rinv;
    } else if (K == K_NATIVEQ) {  // nativeqSwitch::calculate :524-549
      if ((in_dmax >= 0) ? (in_dmax != 0) : (r <= p.dmax)) {
        T res = T(1.0);
        if ((above_d0 >= 0) ? (above_d0 != 0) : (r > p.d0)) {
          const T e = exp_t(p.beta * (r - p.lambda * p.ref));
          res = fast_rcp(T(1.0) + e);
          df = -p.beta * fast_rcp(e + T(2.0) + fast_rcp(e)) * rinv * p.stretch;
        }
        s = fma_t(res, p.stretch, p.shift);
      }
    } else if ((in_dmax >= 0) ? (in_dmax != 0) : !(r > p.dmax)) {
      T x = (r - p.d0) * p.invr0;
      if (above_d0 > 0) x = (x > T(1e-30)) ? x : T(1e-30);  // FP64 says beyond D_0: keep the float x on that side
      if ((above_d0 >= 0) ? (above_d0 != 0) : (x > T(0.0))) {
        T f, fp;
        if (K == K_RAT_R) {
          f = p.preRes;
          fp = p.preDfunc;
          rational_generic(p.type == 8, x, p.preSecDev, p.nn, p.mm, f, fp);
        } else if (K == K_EXP) {  // :375-387
          f = exp_t(-x);
          fp = -f;
        } else if (K == K_GAUSS) {  // :389-401
          f = exp_t(T(-0.5) * x * x);
          fp = -x * f;
        } else if (K == K_SMAP) {  // :434-455
          const T sx = p.c * ipow_dev(x, p.a);
          f = pow_t(T(1.0) + sx, p.d);
          fp = -(T)p.b * sx * fast_rcp(x) * f * fast_rcp(T(1.0) + sx);
        } else if (K == K_CUBIC) {  // :457-469
          const T t1 = x - T(1.0), t2 = fma_t(T(2.0), x, T(1.0));
          fp = T(2.0) * t1 * t2 + T(2.0) * t1 * t1;
          f = t1 * t1 * t2;
        } else if (K == K_TANH) {  // :471-486
          const T t1 = tanh_t(x);
          fp = fma_t(t1, t1, T(-1.0));
          f = T(1.0) - t1;
        } else {  // K_COS :488-507
          f = T(0.0);
          fp = T(0.0);
          if (x <= T(1.0)) {
            T sn, cs;
            sincospi_t(x, &sn, &cs);
            f = T(0.5) * (cs + T(1.0));
            fp = T(-0.5 * 3.141592653589793238462643383279502884) * sn;
          }
        }
        s = fma_t(f, p.stretch, p.shift);
        df = fp * p.stretch * p.invr0 * rinv;  // applystretch :124-130
      } else {
        s = p.stretch + p.shift;
      }
    }
  }
}

// The switching functions jump (value or derivative) at D_MAX and some at D_0.  A pair whose r^2 lands within
// 1e-10 (relative) of such a boundary -- in practice only constructed inputs: lattices, round numbers, the
// reference's own rt20-switch fixtures -- is re-evaluated with the reference's exact operation sequence
// (Tools::pbc with its +100 shift, unfused r^2, correctly rounded sqrt) so that it falls on the same side.
__device__ __forceinline__ bool on_boundary(const DevSwitch& sw, double r2) {
  bool b = fabs(r2 - sw.dmax_2) <= sw.band_dmax;                 // band < 0: no D_MAX, never true
  if (sw.band_d0 >= 0.0) b |= fabs(r2 - sw.d0_2) <= sw.band_d0;  // warp-uniform: only switches with D_0 > 0
  return b;
}

// exact evaluation of one pair (cold paths only)
struct ExactPair {
  double dx, dy, dz, s, df;
};
template <int K>
__device__ __forceinline__ ExactPair exact_pair(const DevPbc& pbc, const DevSwitch& sw, double xi, double yi, double zi,
                                                double xj, double yj, double zj, bool flip) {
  double d[3];
  if (flip) {  // canonical orientation: distance = pos[i1] - pos[i0]
    d[0] = xsub(xi, xj);
    d[1] = xsub(yi, yj);
    d[2] = xsub(zi, zj);
  } else {
    d[0] = xsub(xj, xi);
    d[1] = xsub(yj, yi);
    d[2] = xsub(zj, zi);
  }
  min_image_exact(pbc, d);
  ExactPair o;
  o.dx = d[0];
  o.dy = d[1];
  o.dz = d[2];
  eval_switch<K, true>(sw, norm2_exact(d[0], d[1], d[2]), o.s, o.df);
  return o;
}

template <typename T>
struct LaneAccT {
  T val, vxx, vxy, vxz, vyy, vyz, vzz;
};
using LaneAcc = LaneAccT<double>;
using LaneAccF = LaneAccT<float>;

// one pair seen from atom i.  The reference evaluates every pair once, as distance = pos[i1]-pos[i0] with
// (i0,i1) = (GROUPA atom, GROUPB atom) resp. (lower, higher index) (NeighborList.cpp:147-166), and gives
// -dd to i0 and +dd to i1.  When two periodic images are equally close (perfect crystals: regtest rt42) the
// minimum image of -x is not minus the minimum image of x, so both ends of a pair must use the SAME vector:
// `flip` says atom i is the i1 end; the difference is then taken as r_i - r_j, the image is chosen on that
// canonical vector, and the sign goes into the derivative instead.
// `near` collects "this lane saw a pair on a D_MAX / D_0 boundary": the hot loops only set it (two compares) and the
// row is patched afterwards by row_fixup_*; kernels off the critical path pass INLINE_EXACT and fix the pair in place.
template <int K, int PBC, bool ACC, bool INLINE_EXACT = false>
__device__ __forceinline__ void pair_term(const DevPbc& pbc, const DevSwitch& sw, bool& near, double xi, double yi,
                                          double zi, const SPos& pj, bool flip, double& fx, double& fy, double& fz,
                                          LaneAcc& acc, double qq = 1.0) {
  const unsigned sgn = flip ? 0x80000000u : 0u;
  double dx = flip_sign(pj.x - xi, sgn), dy = flip_sign(pj.y - yi, sgn), dz = flip_sign(pj.z - zi, sgn);
  min_image_fast<PBC>(pbc, dx, dy, dz);
  const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
  double s, df;
  eval_switch<K>(sw, r2, s, df);
  if (K == K_DH || K == K_GHB) {  // q_i q_j (DHENERGY) / eta of the two types (GHBFIX); neither has boundary patches
    s *= qq;
    df *= qq;
  }
  if (INLINE_EXACT) {
    if (on_boundary(sw, r2)) {
      const ExactPair o = exact_pair<K>(pbc, sw, xi, yi, zi, pj.x, pj.y, pj.z, flip);
      dx = o.dx;
      dy = o.dy;
      dz = o.dz;
      s = o.s;
      df = o.df;
    }
  } else {
    near |= on_boundary(sw, r2);
  }
  const double dfs = flip_sign(df, sgn);  // deriv[i0] -= df*d ; deriv[i1] += df*d
  fx = fma(-dfs, dx, fx);
  fy = fma(-dfs, dy, fy);
  fz = fma(-dfs, dz, fz);
  if (ACC) {
    const double gx = df * dx, gy = df * dy, gz = df * dz;
    acc.val += s;
    acc.vxx = fma(gx, dx, acc.vxx);
    acc.vxy = fma(gx, dy, acc.vxy);
    acc.vxz = fma(gx, dz, acc.vxz);
    acc.vyy = fma(gy, dy, acc.vyy);
    acc.vyz = fma(gy, dz, acc.vyz);
    acc.vzz = fma(gz, dz, acc.vzz);
  }
}

// ---- opt-in FP32 sweep (B200COORD_FP32, 1e-5 parity).  The coordinate difference and the minimum image stay in
// FP64 -- a difference of box-sized coordinates rounded to float would carry ~1e-6 nm of absolute error, 1e-5 of a
// contact distance -- and the image vector is then rounded once (6e-8 relative).  r^2, the switching function, dd and
// the sums over one row run on the FP32 pipe; rows are added to the FP64 block accumulators.  Which side of D_MAX and
// D_0 a pair is on -- the derivative (without stretch also the value) jumps there, and among 1e8 pairs a few always lie
// within FP32 rounding of the jump -- is decided on the FP64 r^2 (3 more FP64 operations per pair); the 1e-10 exact
// patch of the FP64 sweep is not applied.
template <int K>
__device__ __forceinline__ bool inside_dmax(const DevSwitchT<float>& sw, double r2d) {
  return (K == K_FASTGAUSS) ? (r2d < sw.dmax_2_f64) : (r2d <= sw.dmax_2_f64);  // fastgaussian: strict (:414)
}
template <int K, int PBC, bool ACC, bool INLINE_EXACT = false>
__device__ __forceinline__ void pair_term(const DevPbc& pbc, const DevSwitchT<float>& sw, bool&, double xi, double yi,
                                          double zi, const SPos& pj, bool flip, float& fx, float& fy, float& fz,
                                          LaneAccF& acc, float qq = 1.0f) {
  const unsigned sgn = flip ? 0x80000000u : 0u;
  double ex = flip_sign(pj.x - xi, sgn), ey = flip_sign(pj.y - yi, sgn), ez = flip_sign(pj.z - zi, sgn);
  min_image_fast<PBC>(pbc, ex, ey, ez);
  const float dx = (float)ex, dy = (float)ey, dz = (float)ez;
  const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  const double r2d = fma(ez, ez, fma(ey, ey, ex * ex));
  float s, df;
  eval_switch<K>(sw, r2, s, df, inside_dmax<K>(sw, r2d), r2d > sw.d0_2_f64);
  if (K == K_DH || K == K_GHB) {
    s *= qq;
    df *= qq;
  }
  const float dfs = flip_sign(df, sgn);
  fx = fmaf(-dfs, dx, fx);
  fy = fmaf(-dfs, dy, fy);
  fz = fmaf(-dfs, dz, fz);
  if (ACC) {
    const float gx = df * dx, gy = df * dy, gz = df * dz;
    acc.val += s;
    acc.vxx = fmaf(gx, dx, acc.vxx);
    acc.vxy = fmaf(gx, dy, acc.vxy);
    acc.vxz = fmaf(gx, dz, acc.vxz);
    acc.vyy = fmaf(gy, dy, acc.vyy);
    acc.vyz = fmaf(gy, dz, acc.vyz);
    acc.vzz = fmaf(gz, dz, acc.vzz);
  }
}

// One prefetched partner record.  All four loaded doubles stay live until they are used (w = slot:abs bits is
// compared as a whole): ptxas otherwise hands a dead destination register of the pending 256-bit load to the next
// temporary, and that write-after-write waits for the load -- the latency the prefetch is there to hide.
struct RecBuf {
  double x, y, z, w;
};
__device__ __forceinline__ void load_rec(const SPos* __restrict__ p, RecBuf& b) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(b.x), "=d"(b.y), "=d"(b.z), "=d"(b.w) : "l"(p));
}

// two pairs at once (masked out when `va` / `vb` is false) as independent instruction streams; the whole warp must
// call this together (FAR votes).
// `wi` = slot:abs bits of atom i; slots are unique, so comparing the whole 64 bits orders by slot.
// FAR: the far part of a row (partners beyond D_MAX when the list was built) -- when no lane of the warp has a pair
// inside far_skip2 = D_MAX^2 + boundary band, every contribution is exactly zero and the trip ends after the
// distance test.  Returns false in that case.
template <int K, int PBC, bool ACC, bool FAR>
__device__ __forceinline__ bool pair_term2(const DevPbc& pbc, const DevSwitch& sw, bool& near, double xi, double yi,
                                           double zi, unsigned long long wi, int two_groups, bool row_is_b,
                                           const RecBuf& pa, const RecBuf& pb, bool va, bool vb, double far_skip2,
                                           double& fx, double& fy, double& fz, LaneAcc& acc, double qqa = 1.0,
                                           double qqb = 1.0) {
  const bool flipa = two_groups ? row_is_b : (wi > (unsigned long long)__double_as_longlong(pa.w));
  const bool flipb = two_groups ? row_is_b : (wi > (unsigned long long)__double_as_longlong(pb.w));
  const unsigned sga = flipa ? 0x80000000u : 0u, sgb = flipb ? 0x80000000u : 0u;
  double ax = flip_sign(pa.x - xi, sga), ay = flip_sign(pa.y - yi, sga), az = flip_sign(pa.z - zi, sga);
  double bx = flip_sign(pb.x - xi, sgb), by = flip_sign(pb.y - yi, sgb), bz = flip_sign(pb.z - zi, sgb);
  min_image_fast<PBC>(pbc, ax, ay, az);
  min_image_fast<PBC>(pbc, bx, by, bz);
  const double ra = fma(az, az, fma(ay, ay, ax * ax));
  const double rb = fma(bz, bz, fma(by, by, bx * bx));
  if (FAR) {
    if (__all_sync(0xffffffffu, (!va || ra > far_skip2) && (!vb || rb > far_skip2))) return false;
  }
  double sa, dfa, sb, dfb;
  eval_switch<K>(sw, ra, sa, dfa);
  eval_switch<K>(sw, rb, sb, dfb);
  if (K == K_DH || K == K_GHB) {
    sa *= qqa;
    dfa *= qqa;
    sb *= qqb;
    dfb *= qqb;
  }
  near |= (va & on_boundary(sw, ra)) | (vb & on_boundary(sw, rb));
  if (!va) {
    sa = 0.0;
    dfa = 0.0;
  }
  if (!vb) {
    sb = 0.0;
    dfb = 0.0;
  }
  const double da = flip_sign(dfa, sga), db = flip_sign(dfb, sgb);
  fx = fma(-da, ax, fma(-db, bx, fx));
  fy = fma(-da, ay, fma(-db, by, fy));
  fz = fma(-da, az, fma(-db, bz, fz));
  if (ACC) {
    const double gax = dfa * ax, gay = dfa * ay, gaz = dfa * az;
    const double gbx = dfb * bx, gby = dfb * by, gbz = dfb * bz;
    acc.val += sa + sb;
    acc.vxx = fma(gax, ax, fma(gbx, bx, acc.vxx));
    acc.vxy = fma(gax, ay, fma(gbx, by, acc.vxy));
    acc.vxz = fma(gax, az, fma(gbx, bz, acc.vxz));
    acc.vyy = fma(gay, ay, fma(gby, by, acc.vyy));
    acc.vyz = fma(gay, az, fma(gby, bz, acc.vyz));
    acc.vzz = fma(gaz, az, fma(gbz, bz, acc.vzz));
  }
  return true;
}

// FP32 flavour of pair_term2 (see the FP32 pair_term above)
template <int K, int PBC, bool ACC, bool FAR>
__device__ __forceinline__ bool pair_term2(const DevPbc& pbc, const DevSwitchT<float>& sw, bool&, double xi, double yi,
                                           double zi, unsigned long long wi, int two_groups, bool row_is_b,
                                           const RecBuf& pa, const RecBuf& pb, bool va, bool vb, float far_skip2,
                                           float& fx, float& fy, float& fz, LaneAccF& acc, float qqa = 1.0f,
                                           float qqb = 1.0f) {
  const bool flipa = two_groups ? row_is_b : (wi > (unsigned long long)__double_as_longlong(pa.w));
  const bool flipb = two_groups ? row_is_b : (wi > (unsigned long long)__double_as_longlong(pb.w));
  const unsigned sga = flipa ? 0x80000000u : 0u, sgb = flipb ? 0x80000000u : 0u;
  double eax = flip_sign(pa.x - xi, sga), eay = flip_sign(pa.y - yi, sga), eaz = flip_sign(pa.z - zi, sga);
  double ebx = flip_sign(pb.x - xi, sgb), eby = flip_sign(pb.y - yi, sgb), ebz = flip_sign(pb.z - zi, sgb);
  min_image_fast<PBC>(pbc, eax, eay, eaz);
  min_image_fast<PBC>(pbc, ebx, eby, ebz);
  const float ax = (float)eax, ay = (float)eay, az = (float)eaz;
  const float bx = (float)ebx, by = (float)eby, bz = (float)ebz;
  const float ra = fmaf(az, az, fmaf(ay, ay, ax * ax));
  const float rb = fmaf(bz, bz, fmaf(by, by, bx * bx));
  if (FAR) {
    if (__all_sync(0xffffffffu, (!va || ra > far_skip2) && (!vb || rb > far_skip2))) return false;
  }
  const double rad = fma(eaz, eaz, fma(eay, eay, eax * eax));
  const double rbd = fma(ebz, ebz, fma(eby, eby, ebx * ebx));
  float sa, dfa, sb, dfb;
  eval_switch<K>(sw, ra, sa, dfa, inside_dmax<K>(sw, rad), rad > sw.d0_2_f64);
  eval_switch<K>(sw, rb, sb, dfb, inside_dmax<K>(sw, rbd), rbd > sw.d0_2_f64);
  if (K == K_DH || K == K_GHB) {
    sa *= qqa;
    dfa *= qqa;
    sb *= qqb;
    dfb *= qqb;
  }
  if (!va) {
    sa = 0.0f;
    dfa = 0.0f;
  }
  if (!vb) {
    sb = 0.0f;
    dfb = 0.0f;
  }
  const float da = flip_sign(dfa, sga), db = flip_sign(dfb, sgb);
  fx = fmaf(-da, ax, fmaf(-db, bx, fx));
  fy = fmaf(-da, ay, fmaf(-db, by, fy));
  fz = fmaf(-da, az, fmaf(-db, bz, fz));
  if (ACC) {
    const float gax = dfa * ax, gay = dfa * ay, gaz = dfa * az;
    const float gbx = dfb * bx, gby = dfb * by, gbz = dfb * bz;
    acc.val += sa + sb;
    acc.vxx = fmaf(gax, ax, fmaf(gbx, bx, acc.vxx));
    acc.vxy = fmaf(gax, ay, fmaf(gbx, by, acc.vxy));
    acc.vxz = fmaf(gax, az, fmaf(gbx, bz, acc.vxz));
    acc.vyy = fmaf(gay, ay, fmaf(gby, by, acc.vyy));
    acc.vyz = fmaf(gay, az, fmaf(gby, bz, acc.vyz));
    acc.vzz = fmaf(gaz, az, fmaf(gbz, bz, acc.vzz));
  }
  return true;
}

// the accumulator a row's pair terms go to: the block's FP64 lane accumulators, or in FP32 mode a per-row FP32
// set that is added to them when the row is done
__device__ __forceinline__ LaneAcc& row_acc(LaneAcc& block, LaneAcc&) { return block; }
__device__ __forceinline__ LaneAccF& row_acc(LaneAcc&, LaneAccF& row) { return row; }
__device__ __forceinline__ void flush_row_acc(LaneAcc&, const LaneAcc&) {}
__device__ __forceinline__ void flush_row_acc(LaneAcc& block, const LaneAccF& row) {
  block.val += (double)row.val;
  block.vxx += (double)row.vxx;
  block.vxy += (double)row.vxy;
  block.vxz += (double)row.vxz;
  block.vyy += (double)row.vyy;
  block.vyz += (double)row.vyz;
  block.vzz += (double)row.vzz;
}

// Patch of one row in which some lane saw a boundary pair (rare: constructed inputs).  Walks the row again, and for
// every pair on a boundary returns (exact contribution - fast contribution).  Out of line and by value so that the
// hot loop keeps its registers; parameters come from global memory.
struct RowFix {
  double fx, fy, fz, val, vxx, vxy, vxz, vyy, vyz, vzz;
};
template <int K, int PBC>
__device__ __forceinline__ void fix_one(const DevPbc& pbc, const DevSwitch& sw, double xi, double yi, double zi,
                                        double xj, double yj, double zj, const double* __restrict__ ri,
                                        const double* __restrict__ rj, bool flip, RowFix& f) {
  const unsigned sgn = flip ? 0x80000000u : 0u;
  double dx = flip_sign(xj - xi, sgn), dy = flip_sign(yj - yi, sgn), dz = flip_sign(zj - zi, sgn);
  min_image_fast<PBC>(pbc, dx, dy, dz);
  const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
  if (!on_boundary(sw, r2)) return;
  double s, df;
  eval_switch<K>(sw, r2, s, df);
  // the exact evaluation starts from the caller's own positions (the records may hold continuous coordinates)
  const ExactPair o = exact_pair<K>(pbc, sw, ri[0], ri[1], ri[2], rj[0], rj[1], rj[2], flip);
  const double dfs = flip_sign(df, sgn), odfs = flip ? -o.df : o.df;
  f.fx += -odfs * o.dx + dfs * dx;
  f.fy += -odfs * o.dy + dfs * dy;
  f.fz += -odfs * o.dz + dfs * dz;
  f.val += o.s - s;
  f.vxx += o.df * o.dx * o.dx - df * dx * dx;
  f.vxy += o.df * o.dx * o.dy - df * dx * dy;
  f.vxz += o.df * o.dx * o.dz - df * dx * dz;
  f.vyy += o.df * o.dy * o.dy - df * dy * dy;
  f.vyz += o.df * o.dy * o.dz - df * dy * dz;
  f.vzz += o.df * o.dz * o.dz - df * dz * dz;
}
template <int K, int PBC>
__device__ __noinline__ RowFix row_fixup_list(const DevPbc* __restrict__ pbc_g, const DevSwitch* __restrict__ sw_g,
                                              const SPos* __restrict__ spos, PosSrc pos,
                                              uint32_t idx_mask, const uint32_t* __restrict__ row, unsigned cnt,
                                              const uint32_t* __restrict__ far_row, unsigned far_cnt, unsigned k,
                                              unsigned lane, int two_groups, bool row_is_b) {
  RowFix f = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const SPos pi = spos[k];
  for (unsigned e = lane; e < cnt + far_cnt; e += 32) {
    const SPos pj = spos[(e < cnt ? row[e] : far_row[e - cnt]) & idx_mask];
    fix_one<K, PBC>(*pbc_g, *sw_g, pi.x, pi.y, pi.z, pj.x, pj.y, pj.z, pos_at(pos, pi.slot), pos_at(pos, pj.slot),
                    two_groups ? row_is_b : (pi.slot > pj.slot), f);
  }
  return f;
}
__device__ __forceinline__ void apply_fix(const RowFix& f, bool accumulate, double& fx, double& fy, double& fz, LaneAcc& acc) {
  fx += f.fx;
  fy += f.fy;
  fz += f.fz;
  if (accumulate) {
    acc.val += f.val;
    acc.vxx += f.vxx;
    acc.vxy += f.vxy;
    acc.vxz += f.vxz;
    acc.vyy += f.vyy;
    acc.vyz += f.vyz;
    acc.vzz += f.vzz;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kSweepThreads = 256;
constexpr int kSweepWarps = kSweepThreads / 32;

static inline unsigned pick_rows_per_block(unsigned rows) {
  // aim for >= 4 resident blocks on each of the 148 SMs; a warp always owns whole rows
  unsigned rpb = rows / (148u * 4u);
  rpb = (rpb / kSweepWarps) * kSweepWarps;
  if (rpb < (unsigned)kSweepWarps) rpb = kSweepWarps;
  if (rpb > 128u) rpb = 128u;  // 16 rows per warp: 3 % faster than 64 (fewer block tails), no gain beyond
  return rpb;
}

// block epilogue: reduce the lane accumulators of all warps and store one partial record
// {value, c[3][3]} with c = sum df d (x) d (symmetric here; the image sweep stores a general 3x3 sum)
__device__ __forceinline__ void block_store_partials(const LaneAcc& a, unsigned long long evals, double* partials,
                                                     unsigned long long* evals_out, unsigned record = 0xffffffffu) {
  if (record == 0xffffffffu) record = blockIdx.x;
  constexpr int kMaxWarps = 32;  // blocks of up to 1024 threads
  __shared__ double sm[kMaxWarps][8];
  __shared__ unsigned long long sev[kMaxWarps];
  const int nwarps = (int)(blockDim.x >> 5);
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double v0 = warp_sum(a.val), v1 = warp_sum(a.vxx), v2 = warp_sum(a.vxy), v3 = warp_sum(a.vxz),
               v4 = warp_sum(a.vyy), v5 = warp_sum(a.vyz), v6 = warp_sum(a.vzz);
  if (lane == 0) {
    sm[wid][0] = v0;
    sm[wid][1] = v1;
    sm[wid][2] = v2;
    sm[wid][3] = v3;
    sm[wid][4] = v4;
    sm[wid][5] = v5;
    sm[wid][6] = v6;
    sev[wid] = evals;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    // record slot -> symmetric component: value | xx xy xz | xy yy yz | xz yz zz
    const int map[10] = {0, 1, 2, 3, 2, 4, 5, 3, 5, 6};
    const int src = map[threadIdx.x];
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) t += sm[w][src];
    partials[(size_t)record * kPartialStride + threadIdx.x] = t;
  }
  if (threadIdx.x == 32) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) t += sev[w];
    if (t) atomicAdd(evals_out, t);
  }
}


}  // namespace b200
