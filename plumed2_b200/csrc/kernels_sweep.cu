// The pair sweep: value + 3N derivatives + virial of CoordinationBase::calculate
// (src/colvar/CoordinationBase.cpp:142-232) as hand-written sm_100a kernels.
//
// Design (see DESIGN.md "pair sweep"):
//   * Every atom i owns a ROW of partners j and is processed by one warp: lanes stride over the row,
//     accumulate -df*d for atom i in registers, and a 5-step shuffle tree reduces the three components.
//     A pair (i,j) is therefore evaluated from both ends -- twice the FP64 work of the reference's half
//     list, but no atomics, no write conflicts and a bit-reproducible summation order.  The metric counts
//     each pair once.
//   * value and virial are accumulated per lane across all rows of a block, reduced once per block
//     (shuffles + shared memory) and written as one partial record per block; k_finalize adds the partials
//     in index order (deterministic) and applies the 1/2 for doubly visited pairs.
//   * FP64 pipe is the bound: the reciprocal / reciprocal square root use the MUFU seed + 2 Newton steps
//     instead of the IEEE division slow path; the minimum image uses 2-add rounding instead of F2I/I2F.
//   * Rows come either from the stream-compacted CSR list (NLIST) or, for NLISTCELLS / no list, from the
//     contiguous sorted ranges of the <=27 stencil cells (no index traffic, coalesced 32-byte records).
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
  K_COUNT
};

static int kind_of(int type) {
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
template <int K>
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
    const double rinv = (r2 > 0.0) ? fast_rsqrt(r2) : 0.0;
    const double r = r2 * rinv;
    if (K == K_NATIVEQ) {  // nativeqSwitch::calculate :524-549
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

struct LaneAcc {
  double val, vxx, vxy, vxz, vyy, vyz, vzz;
};

// one pair seen from atom i.  The reference evaluates every pair once, as distance = pos[i1]-pos[i0] with
// (i0,i1) = (GROUPA atom, GROUPB atom) resp. (lower, higher index) (NeighborList.cpp:147-166), and gives
// -dd to i0 and +dd to i1.  When two periodic images are equally close (perfect crystals: regtest rt42) the
// minimum image of -x is not minus the minimum image of x, so both ends of a pair must use the SAME vector:
// `flip` says atom i is the i1 end; the difference is then taken as r_i - r_j, the image is chosen on that
// canonical vector, and the sign goes into the derivative instead.
template <int K, int PBC, bool ACC>
__device__ __forceinline__ void pair_term(const DevPbc& pbc, const DevSwitch& sw, double xi, double yi, double zi,
                                          const SPos& pj, bool flip, double& fx, double& fy, double& fz, LaneAcc& acc) {
  const unsigned sgn = flip ? 0x80000000u : 0u;
  double dx = flip_sign(pj.x - xi, sgn), dy = flip_sign(pj.y - yi, sgn), dz = flip_sign(pj.z - zi, sgn);
  min_image_fast<PBC>(pbc, dx, dy, dz);
  const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
  double s, df;
  eval_switch<K>(sw, r2, s, df);
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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kSweepThreads = 256;
constexpr int kMaxRowsPerBlock = 64;
constexpr int kSweepWarps = kSweepThreads / 32;

// block epilogue: reduce the lane accumulators of all warps and store one partial record
__device__ __forceinline__ void block_store_partials(const LaneAcc& a, unsigned long long evals, double* partials,
                                                     unsigned long long* evals_out) {
  __shared__ double sm[kSweepWarps][kPartialStride];
  __shared__ unsigned long long sev[kSweepWarps];
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
    for (int w = 0; w < kSweepWarps; ++w) t += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * kPartialStride + threadIdx.x] = t;
  }
  if (threadIdx.x == 32) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < kSweepWarps; ++w) t += sev[w];
    if (t) atomicAdd(evals_out, t);
  }
}

// ------------------------------------------------------------------------------------------------
// rows from the CSR list (classic NLIST)
template <int K, int PBC, bool ACC>
__global__ void __launch_bounds__(kSweepThreads)
    k_sweep_list(SweepArgs a, DevPbc pbc, DevSwitch sw, unsigned rows_per_block, unsigned seg_begin, unsigned seg_end) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  LaneAcc acc = {0, 0, 0, 0, 0, 0, 0};
  unsigned long long evals = 0;
  const unsigned first = seg_begin + blockIdx.x * rows_per_block;
  const unsigned last = min(first + rows_per_block, seg_end);
  __shared__ double s_rows[3 * kMaxRowsPerBlock];  // this block's finished rows, pushed to the peers in one piece
  for (unsigned k = first + wid; k < last; k += kSweepWarps) {
    const SPos pi = load_spos(a.spos + k);
    const unsigned long long base = a.row_start[k - a.row_begin];
    const unsigned cnt = a.row_count[k - a.row_begin];
    const bool row_is_b = (k >= a.n_a);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    const uint32_t* __restrict__ row = a.nbr + base;
    // software pipeline: the neighbour index is fetched two iterations ahead and the 32-byte record one
    // iteration ahead, so the dependent index -> record -> FP64 chain of one pair overlaps the arithmetic of
    // the previous one (the kernel is otherwise bound by the latency of that chain, see profiles/)
    unsigned e = lane;
    uint32_t j_next = (e < cnt) ? __ldg(row + e) : 0u;
    uint32_t j_next2 = (e + 32 < cnt) ? __ldg(row + e + 32) : 0u;
    SPos p_next = load_spos(a.spos + j_next);
    for (; e < cnt; e += 32) {
      const SPos pj = p_next;
      p_next = load_spos(a.spos + j_next2);
      j_next2 = (e + 64 < cnt) ? __ldg(row + e + 64) : 0u;
      const bool flip = a.two_groups ? row_is_b : (pi.slot > pj.slot);
      pair_term<K, PBC, ACC>(pbc, sw, pi.x, pi.y, pi.z, pj, flip, fx, fy, fz, acc);
    }
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (lane == 0) {
      a.sderiv[3 * (size_t)k] = fx;
      a.sderiv[3 * (size_t)k + 1] = fy;
      a.sderiv[3 * (size_t)k + 2] = fz;
      if (a.npeers) {
        s_rows[3 * (k - first)] = fx;
        s_rows[3 * (k - first) + 1] = fy;
        s_rows[3 * (k - first) + 2] = fz;
      }
      evals += cnt;
    }
  }
  if (a.npeers) {
    // fused exchange: the block's rows are contiguous in every rank's row buffer -> one warp per peer streams
    // them over NVLink with coalesced stores while other blocks keep computing
    __syncthreads();
    if ((int)wid < a.npeers && last > first) {
      double* __restrict__ q = a.peers[wid] + 3 * (size_t)first;
      const unsigned m = 3u * (last - first);
      for (unsigned t = lane; t < m; t += 32) q[t] = s_rows[t];
    }
  }
  if (ACC)
    block_store_partials(acc, evals, a.partials, a.evals);
  else if (lane == 0 && evals)
    atomicAdd(a.evals, evals);
}

// ------------------------------------------------------------------------------------------------
// rows from the sorted ranges of the stencil cells (NLISTCELLS superset, or a single 1x1x1 "cell" = no NL)
template <int K, int PBC, bool ACC>
__global__ void __launch_bounds__(kSweepThreads)
    k_sweep_cells(SweepArgs a, DevPbc pbc, DevSwitch sw, unsigned rows_per_block, unsigned seg_begin, unsigned seg_end) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  LaneAcc acc = {0, 0, 0, 0, 0, 0, 0};
  unsigned long long evals = 0;
  const DevGrid& g = a.grid;
  const unsigned first = seg_begin + blockIdx.x * rows_per_block;
  const unsigned last = min(first + rows_per_block, seg_end);
  __shared__ double s_rows[3 * kMaxRowsPerBlock];  // this block's finished rows, pushed to the peers in one piece
  for (unsigned k = first + wid; k < last; k += kSweepWarps) {
    const SPos pi = load_spos(a.spos + k);
    const unsigned my_grp = (k < a.n_a) ? 0u : 1u;
    const unsigned other = a.two_groups ? (1u - my_grp) : 0u;
    int c[3];
    cell_coords(g, (int)a.scell[k], c);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    unsigned cnt = 0;
    for_each_stencil_range(g, c, other * (unsigned)g.ncell, a.cstart, a.ccount, [&](uint32_t s0, uint32_t m, int, int, int) {
      cnt += m;
#pragma unroll 2
      for (uint32_t e = lane; e < m; e += 32) {
        const uint32_t j = s0 + e;
        const SPos pj = load_spos(a.spos + j);
        const bool valid = (j != k) && (!a.check_abs || pj.abs_index != pi.abs_index);
        const bool flip = a.two_groups ? (my_grp == 1u) : (pi.slot > pj.slot);
        if (valid) pair_term<K, PBC, ACC>(pbc, sw, pi.x, pi.y, pi.z, pj, flip, fx, fy, fz, acc);
      }
    });
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (lane == 0) {
      a.sderiv[3 * (size_t)k] = fx;
      a.sderiv[3 * (size_t)k + 1] = fy;
      a.sderiv[3 * (size_t)k + 2] = fz;
      if (a.npeers) {
        s_rows[3 * (k - first)] = fx;
        s_rows[3 * (k - first) + 1] = fy;
        s_rows[3 * (k - first) + 2] = fz;
      }
      evals += cnt;
    }
  }
  if (a.npeers) {
    // fused exchange: the block's rows are contiguous in every rank's row buffer -> one warp per peer streams
    // them over NVLink with coalesced stores while other blocks keep computing
    __syncthreads();
    if ((int)wid < a.npeers && last > first) {
      double* __restrict__ q = a.peers[wid] + 3 * (size_t)first;
      const unsigned m = 3u * (last - first);
      for (unsigned t = lane; t < m; t += 32) q[t] = s_rows[t];
    }
  }
  if (ACC)
    block_store_partials(acc, evals, a.partials, a.evals);
  else if (lane == 0 && evals)
    atomicAdd(a.evals, evals);
}

// ------------------------------------------------------------------------------------------------
// PAIR style: pair k = (k, k+n_a) (NeighborList.cpp:150-152); each atom slot occurs in exactly one pair
template <int K, int PBC>
__global__ void __launch_bounds__(kSweepThreads)
    k_sweep_pairs(const double* __restrict__ pos, const uint32_t* __restrict__ abs_index, const uint8_t* __restrict__ active,
                  unsigned n_a, unsigned pair_begin, unsigned pair_end, DevPbc pbc, DevSwitch sw, double* __restrict__ out,
                  double* partials, unsigned long long* evals_out) {
  LaneAcc acc = {0, 0, 0, 0, 0, 0, 0};
  unsigned long long evals = 0;
  const unsigned k = pair_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (k < pair_end) {
    const size_t ia = 3 * (size_t)k, ib = 3 * (size_t)(k + n_a);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    const bool on = (!active || active[k]) && abs_index[k] != abs_index[k + n_a];
    if (on) {
      SPos pj;
      pj.x = pos[ib];
      pj.y = pos[ib + 1];
      pj.z = pos[ib + 2];
      pair_term<K, PBC, true>(pbc, sw, pos[ia], pos[ia + 1], pos[ia + 2], pj, false, fx, fy, fz, acc);
      evals = 1;
    }
    out[ia] = fx;  // deriv[i0] -= dd
    out[ia + 1] = fy;
    out[ia + 2] = fz;
    out[ib] = -fx;  // deriv[i1] += dd
    out[ib + 1] = -fy;
    out[ib + 2] = -fz;
  }
  evals = (unsigned long long)warp_sum((double)evals);
  block_store_partials(acc, (threadIdx.x & 31) == 0 ? evals : 0ull, partials, evals_out);
}

// ------------------------------------------------------------------------------------------------
// fixed-order sum of the block partials; virial = -weight * sum(df d(x)d), value = weight * sum(s).
// 1024 threads = 128 groups x 8 components; group g adds records g, g+128, ... and the 128 group sums are
// added in index order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(1024) k_finalize(const double* __restrict__ partials, int nblocks, double weight,
                                                   double* __restrict__ tail) {
  __shared__ double sm[128][kPartialStride];
  const int comp = threadIdx.x & 7, grp = threadIdx.x >> 3;
  double t = 0.0;
  for (int b = grp; b < nblocks; b += 128) t += partials[(size_t)b * kPartialStride + comp];
  sm[grp][comp] = t;
  __syncthreads();
  if (threadIdx.x < 7) {
    double r = 0.0;
    for (int g2 = 0; g2 < 128; ++g2) r += sm[g2][threadIdx.x];
    sm[0][threadIdx.x] = r * weight;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double val = sm[0][0], xx = -sm[0][1], xy = -sm[0][2], xz = -sm[0][3], yy = -sm[0][4], yz = -sm[0][5],
                 zz = -sm[0][6];
    tail[0] = xx; tail[1] = xy; tail[2] = xz;
    tail[3] = xy; tail[4] = yy; tail[5] = yz;
    tail[6] = xz; tail[7] = yz; tail[8] = zz;
    tail[9] = val;
  }
}

__global__ void k_unsort_derivs(const double* __restrict__ sderiv, const SPos* __restrict__ spos, unsigned n,
                                double* __restrict__ out) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const size_t o = 3 * (size_t)spos[k].slot;
  out[o] = sderiv[3 * (size_t)k];
  out[o + 1] = sderiv[3 * (size_t)k + 1];
  out[o + 2] = sderiv[3 * (size_t)k + 2];
}

// ------------------------------------------------------------------------------------------------
// dispatch
static unsigned pick_rows_per_block(unsigned rows) {
  // aim for >= 4 resident blocks on each of the 148 SMs; a warp always owns whole rows
  unsigned rpb = rows / (148u * 4u);
  rpb = (rpb / kSweepWarps) * kSweepWarps;
  if (rpb < (unsigned)kSweepWarps) rpb = kSweepWarps;
  if (rpb > 64u) rpb = 64u;
  return rpb;
}

template <int K, int PBC, bool LIST>
static int run_sweep(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st) {
  // rows that accumulate value+virial: SingleList -> all; TwoList -> only the A rows
  const unsigned acc_end = a.two_groups ? min(a.row_end, a.n_a) : a.row_end;
  int nblocks = 0;
  if (a.row_begin < acc_end) {
    const unsigned rows = acc_end - a.row_begin;
    const unsigned rpb = pick_rows_per_block(rows);
    nblocks = (int)((rows + rpb - 1) / rpb);
    if (LIST)
      k_sweep_list<K, PBC, true><<<nblocks, kSweepThreads, 0, st>>>(a, pbc, sw, rpb, a.row_begin, acc_end);
    else
      k_sweep_cells<K, PBC, true><<<nblocks, kSweepThreads, 0, st>>>(a, pbc, sw, rpb, a.row_begin, acc_end);
  }
  const unsigned b_begin = max(a.row_begin, acc_end);
  if (b_begin < a.row_end) {
    const unsigned rows = a.row_end - b_begin;
    const unsigned rpb = pick_rows_per_block(rows);
    const int nb = (int)((rows + rpb - 1) / rpb);
    if (LIST)
      k_sweep_list<K, PBC, false><<<nb, kSweepThreads, 0, st>>>(a, pbc, sw, rpb, b_begin, a.row_end);
    else
      k_sweep_cells<K, PBC, false><<<nb, kSweepThreads, 0, st>>>(a, pbc, sw, rpb, b_begin, a.row_end);
  }
  return nblocks;
}

template <int K, bool LIST>
static int run_sweep_pbc(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st) {
  switch (pbc.type) {
    case 0: return run_sweep<K, 0, LIST>(a, pbc, sw, st);
    case 1: return run_sweep<K, 1, LIST>(a, pbc, sw, st);
    default: return run_sweep<K, 2, LIST>(a, pbc, sw, st);
  }
}

template <bool LIST>
static int run_sweep_kind(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st) {
  switch (kind_of(sw.type)) {
    case K_FIX6: return run_sweep_pbc<K_FIX6, LIST>(a, pbc, sw, st);
    case K_FIXN: return run_sweep_pbc<K_FIXN, LIST>(a, pbc, sw, st);
    case K_RAT_R2: return run_sweep_pbc<K_RAT_R2, LIST>(a, pbc, sw, st);
    case K_RAT_R: return run_sweep_pbc<K_RAT_R, LIST>(a, pbc, sw, st);
    case K_EXP: return run_sweep_pbc<K_EXP, LIST>(a, pbc, sw, st);
    case K_GAUSS: return run_sweep_pbc<K_GAUSS, LIST>(a, pbc, sw, st);
    case K_FASTGAUSS: return run_sweep_pbc<K_FASTGAUSS, LIST>(a, pbc, sw, st);
    case K_SMAP: return run_sweep_pbc<K_SMAP, LIST>(a, pbc, sw, st);
    case K_CUBIC: return run_sweep_pbc<K_CUBIC, LIST>(a, pbc, sw, st);
    case K_TANH: return run_sweep_pbc<K_TANH, LIST>(a, pbc, sw, st);
    case K_COS: return run_sweep_pbc<K_COS, LIST>(a, pbc, sw, st);
    case K_NATIVEQ: return run_sweep_pbc<K_NATIVEQ, LIST>(a, pbc, sw, st);
    default: return -1;
  }
}

int launch_sweep_list(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st) {
  return run_sweep_kind<true>(a, pbc, sw, st);
}
int launch_sweep_cells(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st) {
  return run_sweep_kind<false>(a, pbc, sw, st);
}

template <int K>
static int run_pairs(const double* pos, const uint32_t* abs_index, const uint8_t* active, unsigned n_a, unsigned pb,
                     unsigned pe, const DevPbc& pbc, const DevSwitch& sw, double* out, double* partials,
                     unsigned long long* evals, cudaStream_t st) {
  const int nblocks = (int)((pe - pb + kSweepThreads - 1) / kSweepThreads);
  if (nblocks == 0) return 0;
  switch (pbc.type) {
    case 0: k_sweep_pairs<K, 0><<<nblocks, kSweepThreads, 0, st>>>(pos, abs_index, active, n_a, pb, pe, pbc, sw, out, partials, evals); break;
    case 1: k_sweep_pairs<K, 1><<<nblocks, kSweepThreads, 0, st>>>(pos, abs_index, active, n_a, pb, pe, pbc, sw, out, partials, evals); break;
    default: k_sweep_pairs<K, 2><<<nblocks, kSweepThreads, 0, st>>>(pos, abs_index, active, n_a, pb, pe, pbc, sw, out, partials, evals); break;
  }
  return nblocks;
}

int launch_sweep_pairs(const double* pos, const uint32_t* abs_index, const uint8_t* active, unsigned n_a,
                       unsigned pair_begin, unsigned pair_end, const DevPbc& pbc, const DevSwitch& sw, double* out,
                       double* partials, unsigned long long* evals, cudaStream_t st) {
#define B200_PAIR_CASE(KK) \
  case KK: return run_pairs<KK>(pos, abs_index, active, n_a, pair_begin, pair_end, pbc, sw, out, partials, evals, st);
  switch (kind_of(sw.type)) {
    B200_PAIR_CASE(K_FIX6)
    B200_PAIR_CASE(K_FIXN)
    B200_PAIR_CASE(K_RAT_R2)
    B200_PAIR_CASE(K_RAT_R)
    B200_PAIR_CASE(K_EXP)
    B200_PAIR_CASE(K_GAUSS)
    B200_PAIR_CASE(K_FASTGAUSS)
    B200_PAIR_CASE(K_SMAP)
    B200_PAIR_CASE(K_CUBIC)
    B200_PAIR_CASE(K_TANH)
    B200_PAIR_CASE(K_COS)
    B200_PAIR_CASE(K_NATIVEQ)
    default: return -1;
  }
#undef B200_PAIR_CASE
}

void launch_finalize(const double* partials, int nblocks, double weight, double* out_tail, cudaStream_t st) {
  k_finalize<<<1, 1024, 0, st>>>(partials, nblocks, weight, out_tail);
}

void launch_unsort_derivs(const double* sderiv, const SPos* spos, unsigned n, double* out, cudaStream_t st) {
  if (n) k_unsort_derivs<<<(n + 255) / 256, 256, 0, st>>>(sderiv, spos, n, out);
}

}  // namespace b200
