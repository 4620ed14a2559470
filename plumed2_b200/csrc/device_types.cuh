// Device-side parameter blocks and small math helpers shared by the build and sweep kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// ---- periodic box, passed by value to kernels (lives in the constant bank as a kernel parameter)
struct DevPbc {
  int type;          // 0 none (plain delta), 1 orthorhombic, 2 generic  (Pbc.h:52)
  double box[9];     // ortho path uses box[4k] and inv_box[4k]   (Pbc.cpp:377-379)
  double inv_box[9];
  double reduced[9]; // generic path (Pbc.cpp:381-411)
  double inv_reduced[9];
  int nshift[8];
  double shifts[8][6][3];
};

// ---- cell grid used for binning (either the reference's LinkCells grid or our own margin grid)
struct DevGrid {
  int bbox;            // 1: no box -> bounding box around the atoms, origin subtracted (LinkCells.cpp:49-73, :279-281)
  int stencil_pbc;     // 1: neighbour stencil wraps, 0: clamps (the usePbc argument of addRequiredCells, :195-239)
  int radius;          // stencil half-width in cells: 1 = the reference's 27 cells; 2 = our finer NLIST search grid
  int n[3];            // cells per direction
  int ncell;           // n0*n1*n2
  double inv_box_t[9]; // transpose(invBox): fpos = inv_box_t * pos   (Pbc::realToScaled, Pbc.cpp:472-474)
  double origin[3];
};

// ---- switching function (mirrors b200coord_switch / switchContainers::Data).  T = double, or float for the
// opt-in FP32 sweep (B200COORD_FP32), whose copy is rounded once on the host (to_f32).
template <typename T>
struct DevSwitchT {
  int type;
  T d0, dmax, dmax_2, invr0, invr0_2, stretch, shift;
  int nn, mm;
  T preRes, preDfunc, preSecDev;
  int nnf, mmf;
  T preDfuncF, preSecDevF;
  int a, b;
  T c, d;
  T beta, lambda, ref;
  // derived on the host (to_dev_switch): constant factors of the r^2 fast paths
  T d0_2;       // d0^2
  T band_dmax;  // |r^2 - dmax^2| below this -> exact re-evaluation (0 when there is no D_MAX)
  T band_d0;    // same around d0^2 (negative when d0 == 0: never)
  T pre_df;   // 2*invr0_2*stretch
  T fix_df;   // -(N/2)*pre_df for rationalfixN
  // FP32 sweep only: D_MAX^2 and D_0^2 in double -- which side of a jump a pair is on is decided on the FP64 r^2
  double dmax_2_f64, d0_2_f64;
};
using DevSwitch = DevSwitchT<double>;

static inline float clamp_f32(double v) {  // "infinite" D_MAX (DBL_MAX) and friends stay finite-comparable in float
  if (v > 3.0e38) return 3.0e38f;
  if (v < -3.0e38) return -3.0e38f;
  return (float)v;
}
static inline DevSwitchT<float> to_f32(const DevSwitch& s) {
  DevSwitchT<float> f;
  f.type = s.type;
  f.d0 = clamp_f32(s.d0); f.dmax = clamp_f32(s.dmax); f.dmax_2 = clamp_f32(s.dmax_2); f.invr0 = clamp_f32(s.invr0);
  f.invr0_2 = clamp_f32(s.invr0_2); f.stretch = clamp_f32(s.stretch); f.shift = clamp_f32(s.shift);
  f.nn = s.nn; f.mm = s.mm;
  f.preRes = clamp_f32(s.preRes); f.preDfunc = clamp_f32(s.preDfunc); f.preSecDev = clamp_f32(s.preSecDev);
  f.nnf = s.nnf; f.mmf = s.mmf;
  f.preDfuncF = clamp_f32(s.preDfuncF); f.preSecDevF = clamp_f32(s.preSecDevF);
  f.a = s.a; f.b = s.b;
  f.c = clamp_f32(s.c); f.d = clamp_f32(s.d);
  f.beta = clamp_f32(s.beta); f.lambda = clamp_f32(s.lambda); f.ref = clamp_f32(s.ref);
  f.d0_2 = clamp_f32(s.d0_2);
  f.band_dmax = -1.0f;  // no boundary patch in FP32 mode
  f.band_d0 = -1.0f;
  f.pre_df = clamp_f32(s.pre_df); f.fix_df = clamp_f32(s.fix_df);
  f.dmax_2_f64 = s.dmax_2;
  f.d0_2_f64 = s.d0_2;
  return f;
}

// sorted atom record: position + slot bookkeeping in one 32-byte sector
struct __align__(32) SPos {
  double x, y, z;
  uint32_t abs_index;  // absolute atom index (self-pair skip, CoordinationBase.cpp:183)
  uint32_t slot;       // index in the caller's position array
};

// one sorted record with a single 256-bit read-only load (LDG.E.256 on sm_100a): a gather costs one L1
// wavefront per distinct 128-byte line instead of two
__device__ __forceinline__ SPos load_spos(const SPos* __restrict__ p) {
  double x, y, z, w;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(p));
  SPos r;
  r.x = x;
  r.y = y;
  r.z = z;
  const unsigned long long bits = (unsigned long long)__double_as_longlong(w);
  r.abs_index = (uint32_t)(bits & 0xffffffffull);
  r.slot = (uint32_t)(bits >> 32);
  return r;
}

// ------------------------------------------------------------------ exact (never contracted) arithmetic
// Used wherever the result feeds a comparison that must reproduce the reference's x86-64 non-FMA build
// bit for bit: cell assignment and the r^2 <= cutoff^2 neighbour test.
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }

// Tools::pbc (Tools.h:557-564): x+=100; x - int(x +/- 0.5)  with C truncation
__device__ __forceinline__ double tools_pbc_exact(double x) {
  x = xadd(x, 100.0);
  const double h = (x >= 0.0) ? xadd(x, 0.5) : xsub(x, 0.5);
  return xsub(x, (double)__double2int_rz(h));
}

// modulo2 (LoopUnroller.h:146-152): (x*x + y*y) + z*z
__device__ __forceinline__ double norm2_exact(double x, double y, double z) {
  return xadd(xadd(xmul(x, x), xmul(y, y)), xmul(z, z));
}

// row-vector x matrix with accumulation from 0 in j order (Tensor.h:451-458)
__device__ __forceinline__ void vecmat_exact(const double a[3], const double* __restrict__ m, double out[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double t = xadd(0.0, xmul(a[0], m[i]));  // the reference accumulates from 0 (matters for -0.0)
    t = xadd(t, xmul(a[1], m[3 + i]));
    t = xadd(t, xmul(a[2], m[6 + i]));
    out[i] = t;
  }
}

// Pbc::distance (Pbc.cpp:362-415) reproduced operation by operation; d = p1 - p0 on entry
__device__ __forceinline__ void min_image_exact(const DevPbc& pbc, double d[3]) {
  if (pbc.type == 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = xmul(tools_pbc_exact(xmul(d[k], pbc.inv_box[4 * k])), pbc.box[4 * k]);
  } else if (pbc.type == 2) {
    double s[3];
    vecmat_exact(d, pbc.inv_reduced, s);
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] = tools_pbc_exact(s[k]);
    vecmat_exact(s, pbc.reduced, d);
    if (xadd(xadd(fabs(s[0]), fabs(s[1])), fabs(s[2])) > 0.5) {
      const int o = 4 * (s[0] > 0 ? 1 : 0) + 2 * (s[1] > 0 ? 1 : 0) + (s[2] > 0 ? 1 : 0);
      double best[3] = {d[0], d[1], d[2]};
      double lbest = norm2_exact(d[0], d[1], d[2]);
      const int ns = pbc.nshift[o];
      for (int i = 0; i < ns; ++i) {
        const double t0 = xadd(d[0], pbc.shifts[o][i][0]);
        const double t1 = xadd(d[1], pbc.shifts[o][i][1]);
        const double t2 = xadd(d[2], pbc.shifts[o][i][2]);
        const double lt = norm2_exact(t0, t1, t2);
        if (lt < lbest) {
          lbest = lt;
          best[0] = t0;
          best[1] = t1;
          best[2] = t2;
        }
      }
      d[0] = best[0];
      d[1] = best[1];
      d[2] = best[2];
    }
  }
}

// conditional negation on the integer pipe (a DADD/DMUL would spend an FP64 issue slot on a sign flip)
__device__ __forceinline__ double flip_sign(double x, unsigned sign_mask /*0 or 0x80000000*/) {
  return __hiloint2double(__double2hiint(x) ^ (int)sign_mask, __double2loint(x));
}

__device__ __forceinline__ float flip_sign(float x, unsigned sign_mask /*0 or 0x80000000*/) {
  return __int_as_float(__float_as_int(x) ^ (int)sign_mask);
}

// ------------------------------------------------------------------ fast arithmetic for the sweep
// floor(x) with two FP64 adds: the first add rounds toward -inf onto the integer grid of the magic constant
// (valid for |x| < 2^51).  Callers pass x = t + 0.5 (folded into an FMA) to get the round-half-up of
// Tools::pbc (Tools.h:557-564: x - int(x+0.5) after the +100 shift), including its behaviour at exact ties.
__device__ __forceinline__ double fast_floor(double x) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52
  return __dsub_rn(__dadd_rd(x, magic), magic);
}

// 1/a: hardware seed (MUFU.RCP64H) + two Newton steps -> ~1 ulp, no slow path, no division by zero care
__device__ __forceinline__ double fast_rcp(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  x = fma(x, e, x);
  e = fma(-a, x, 1.0);
  x = fma(x, e, x);
  return x;
}

// 1/sqrt(a) for a>0: MUFU.RSQ64H seed + two Newton steps
__device__ __forceinline__ double fast_rsqrt(double a) {
  double x;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  const double ha = 0.5 * a;
  double e = fma(-ha * x, x, 0.5);
  x = fma(x, e, x);
  e = fma(-ha * x, x, 0.5);
  x = fma(x, e, x);
  return x;
}

// FP32 sweep: MUFU seed + one Newton step (~1 ulp)
__device__ __forceinline__ float fast_rcp(float a) {
  float x;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(a));
  const float e = fmaf(-a, x, 1.0f);
  return fmaf(x, e, x);
}
__device__ __forceinline__ float fast_rsqrt(float a) {
  float x;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(a));
  const float e = fmaf(-0.5f * a * x, x, 0.5f);
  return fmaf(x, e, x);
}

template <typename T>
__device__ __forceinline__ T ipow_dev(T base, int e) {  // Tools::fastpow (Tools.h:581-595)
  if (e < 0) {
    e = -e;
    base = fast_rcp(base);
  }
  T r = T(1.0);
  while (e) {
    if (e & 1) r *= base;
    e >>= 1;
    base *= base;
  }
  return r;
}

// minimum image for the sweep (1e-10 parity, not bit parity): same algorithm, fused arithmetic
template <int PBC>
__device__ __forceinline__ void min_image_fast(const DevPbc& pbc, double& dx, double& dy, double& dz) {
  if (PBC == 1) {
    dx = fma(-fast_floor(fma(dx, pbc.inv_box[0], 0.5)), pbc.box[0], dx);
    dy = fma(-fast_floor(fma(dy, pbc.inv_box[4], 0.5)), pbc.box[4], dy);
    dz = fma(-fast_floor(fma(dz, pbc.inv_box[8], 0.5)), pbc.box[8], dz);
  } else if (PBC == 2) {
    const double* ir = pbc.inv_reduced;
    const double* rd = pbc.reduced;
    double s0 = fma(dz, ir[6], fma(dy, ir[3], dx * ir[0]));
    double s1 = fma(dz, ir[7], fma(dy, ir[4], dx * ir[1]));
    double s2 = fma(dz, ir[8], fma(dy, ir[5], dx * ir[2]));
    s0 -= fast_floor(s0 + 0.5);
    s1 -= fast_floor(s1 + 0.5);
    s2 -= fast_floor(s2 + 0.5);
    dx = fma(s2, rd[6], fma(s1, rd[3], s0 * rd[0]));
    dy = fma(s2, rd[7], fma(s1, rd[4], s0 * rd[1]));
    dz = fma(s2, rd[8], fma(s1, rd[5], s0 * rd[2]));
    if (fabs(s0) + fabs(s1) + fabs(s2) > 0.5) {
      const int o = 4 * (s0 > 0 ? 1 : 0) + 2 * (s1 > 0 ? 1 : 0) + (s2 > 0 ? 1 : 0);
      double bx = dx, by = dy, bz = dz;
      double lbest = fma(dz, dz, fma(dy, dy, dx * dx));
      const int ns = pbc.nshift[o];
      for (int i = 0; i < ns; ++i) {
        const double tx = dx + pbc.shifts[o][i][0];
        const double ty = dy + pbc.shifts[o][i][1];
        const double tz = dz + pbc.shifts[o][i][2];
        const double lt = fma(tz, tz, fma(ty, ty, tx * tx));
        if (lt < lbest) {
          lbest = lt;
          bx = tx;
          by = ty;
          bz = tz;
        }
      }
      dx = bx;
      dy = by;
      dz = bz;
    }
  }
}

}  // namespace b200
