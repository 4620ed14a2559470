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

// rational<>::doRational (SwitchingFunction.cpp:258-283); res/dfn preset to preRes/preDfunc(F)
__device__ __forceinline__ void rational_generic(bool simplified, double x, double secdev, int N, int M, double& res,
                                                 double& dfn) {
  if (simplified) {
    const double t = ipow_dev(x, N - 1);
    res = fast_rcp(fma(t, x, 1.0));
    dfn = -(double)N * t * res * res;
  } else {
    const double hi = 1.0 + 5.0e10 * 2.220446049250313e-16, lo = 1.0 - 5.0e10 * 2.220446049250313e-16;
    if (!((x > lo) && (x < hi))) {
      const double tn = ipow_dev(x, N - 1);
      const double tm = ipow_dev(x, M - 1);
      const double num = fma(-tn, x, 1.0);
      const double iden = fast_rcp(fma(-tm, x, 1.0));
      res = num * iden;
      dfn = (((double)M * res * tm) - ((double)N * tn)) * iden;
    } else {
      const double dx = x - 1.0;
      res = res + dx * (dfn + 0.5 * dx * secdev);
      dfn = dfn + dx * secdev;
    }
  }
}

// (s, df=(1/r) ds/dr) of SwitchingFunction::calculateSqr for kind K, stretch/shift and D_MAX applied
template <int K, bool XS = false>
__device__ __forceinline__ void eval_switch(const DevSwitch& p, double r2, double& s, double& df) {
  s = 0.0;
  df = 0.0;
  if (K == K_FIX6 || K == K_FIXN || K == K_RAT_R2) {
    if (r2 <= p.dmax_2) {  // fixedRational<N>::calculateSqr :203-215, rational<fast>::calculateSqr :289-303
      const double y = r2 * p.invr0_2;
      double res, d;
      if (K == K_FIX6) {
        const double t = y * y;
        res = fast_rcp(fma(t, y, 1.0));
        df = (t * res) * (res * p.fix_df);
      } else if (K == K_FIXN) {
        const double t = ipow_dev(y, p.nnf - 1);
        res = fast_rcp(fma(t, y, 1.0));
        df = (t * res) * (res * p.fix_df);
      } else {
        res = p.preRes;
        d = p.preDfuncF;
        rational_generic(p.type == 9, y, p.preSecDevF, p.nnf, p.mmf, res, d);
        df = d * p.pre_df;
      }
      s = fma(res, p.stretch, p.shift);
    }
  } else if (K == K_FASTGAUSS) {  // fastgaussianSwitch::calculateSqr :414-431
    if (r2 < p.dmax_2) {
      s = 1.0;
      if (r2 > 0.0) {
        const double res = exp(-0.5 * r2);
        df = -res * p.stretch;
        s = fma(res, p.stretch, p.shift);
      }
    }
  } else {  // baseSwitch::calculateSqr -> calculate(sqrt(r2)) :135-149, :181-183
    double rinv, r;
    if (XS) {  // correctly rounded sqrt: comparisons of r with D_0 / D_MAX match the reference bit for bit
      r = sqrt(r2);
      rinv = (r2 > 0.0) ? 1.0 / r : 0.0;
    } else {
      rinv = (r2 > 0.0) ? fast_rsqrt(r2) : 0.0;
      r = r2 * rinv;
    }
    if (K == K_GHB) {  // GHBFIX::pairing [* eta: caller]; C1 at D_0, at the joint and at D_MAX: no boundary patch needed
      if (!(r2 > p.dmax_2)) {
        const double rdist = r - p.d0;
        s = -1.0;
        if (rdist > p.c) {
          s += p.preRes + rdist * (p.preDfunc + p.preSecDev * rdist);
          df = (p.preDfunc + 2.0 * p.preSecDev * rdist) * rinv;
        } else if (rdist > 0.0) {
          s += p.d * (rdist * rdist);
          df = 2.0 * p.d * rdist * rinv;
        }
      }
    } else if (K == K_DH) {  // DHEnergy::pairing: tmp = exp(-k r)/r * constant/epsilon [* q_i q_j: caller]; dfunc = -(k+1/r) tmp / r
      const double tmp = exp(-p.beta * r) * rinv * p.lambda;
      s = tmp;
      df = -(p.beta + rinv) * tmp * rinv;
    } else if (K == K_NATIVEQ) {  // nativeqSwitch::calculate :524-549
      if (r <= p.dmax) {
        double res = 1.0;
        if (r > p.d0) {
          const double e = exp(p.beta * (r - p.lambda * p.ref));
          res = fast_rcp(1.0 + e);
          df = -p.beta * fast_rcp(e + 2.0 + fast_rcp(e)) * rinv * p.stretch;
        }
        s = fma(res, p.stretch, p.shift);
      }
    } else if (!(r > p.dmax)) {
      const double x = (r - p.d0) * p.invr0;
      if (x > 0.0) {
        double f, fp;
        if (K == K_RAT_R) {
          f = p.preRes;
          fp = p.preDfunc;
          rational_generic(p.type == 8, x, p.preSecDev, p.nn, p.mm, f, fp);
        } else if (K == K_EXP) {  // :375-387
          f = exp(-x);
          fp = -f;
        } else if (K == K_GAUSS) {  // :389-401
          f = exp(-0.5 * x * x);
          fp = -x * f;
        } else if (K == K_SMAP) {  // :434-455
          const double sx = p.c * ipow_dev(x, p.a);
          f = pow(1.0 + sx, p.d);
          fp = -(double)p.b * sx * fast_rcp(x) * f * fast_rcp(1.0 + sx);
        } else if (K == K_CUBIC) {  // :457-469
          const double t1 = x - 1.0, t2 = fma(2.0, x, 1.0);
          fp = 2.0 * t1 * t2 + 2.0 * t1 * t1;
          f = t1 * t1 * t2;
        } else if (K == K_TANH) {  // :471-486
          const double t1 = tanh(x);
          fp = fma(t1, t1, -1.0);
          f = 1.0 - t1;
        } else {  // K_COS :488-507
          f = 0.0;
          fp = 0.0;
          if (x <= 1.0) {
            double sn, cs;
            sincospi(x, &sn, &cs);
            f = 0.5 * (cs + 1.0);
            fp = -0.5 * 3.141592653589793238462643383279502884 * sn;
          }
        }
        s = fma(f, p.stretch, p.shift);
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

struct LaneAcc {
  double val, vxx, vxy, vxz, vyy, vyz, vzz;
};

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

// Patch of one row in which some lane saw a boundary pair (rare: constructed inputs).  Walks the row again, and for
// every pair on a boundary returns (exact contribution - fast contribution).  Out of line and by value so that the
// hot loop keeps its registers; parameters come from global memory.
struct RowFix {
  double fx, fy, fz, val, vxx, vxy, vxz, vyy, vyz, vzz;
};
template <int K, int PBC>
__device__ __forceinline__ void fix_one(const DevPbc& pbc, const DevSwitch& sw, double xi, double yi, double zi,
                                        double xj, double yj, double zj, bool flip, RowFix& f) {
  const unsigned sgn = flip ? 0x80000000u : 0u;
  double dx = flip_sign(xj - xi, sgn), dy = flip_sign(yj - yi, sgn), dz = flip_sign(zj - zi, sgn);
  min_image_fast<PBC>(pbc, dx, dy, dz);
  const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
  if (!on_boundary(sw, r2)) return;
  double s, df;
  eval_switch<K>(sw, r2, s, df);
  const ExactPair o = exact_pair<K>(pbc, sw, xi, yi, zi, xj, yj, zj, flip);
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
                                              const SPos* __restrict__ spos, const uint32_t* __restrict__ row, unsigned cnt,
                                              const uint32_t* __restrict__ far_row, unsigned far_cnt, unsigned k,
                                              unsigned lane, int two_groups, bool row_is_b) {
  RowFix f = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const SPos pi = spos[k];
  for (unsigned e = lane; e < cnt + far_cnt; e += 32) {
    const SPos pj = spos[e < cnt ? row[e] : far_row[e - cnt]];
    fix_one<K, PBC>(*pbc_g, *sw_g, pi.x, pi.y, pi.z, pj.x, pj.y, pj.z, two_groups ? row_is_b : (pi.slot > pj.slot), f);
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

constexpr int kSweepThreads = 256;
constexpr int kSweepWarps = kSweepThreads / 32;

// block epilogue: reduce the lane accumulators of all warps and store one partial record
__device__ __forceinline__ void block_store_partials(const LaneAcc& a, unsigned long long evals, double* partials,
                                                     unsigned long long* evals_out) {
  constexpr int kMaxWarps = 32;  // blocks of up to 1024 threads
  __shared__ double sm[kMaxWarps][kPartialStride];
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
  if (threadIdx.x < 7) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) t += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * kPartialStride + threadIdx.x] = t;
  }
  if (threadIdx.x == 32) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) t += sev[w];
    if (t) atomicAdd(evals_out, t);
  }
}


}  // namespace b200
