// The list and cell sweeps as templates on the arithmetic type (double: kernels_sweep.cu; float, the opt-in
// B200COORD_FP32 mode: kernels_sweep_f32.cu -- two translation units so that they compile in parallel).
#pragma once
#include <cstdlib>
#include <type_traits>

#include "sweep_math.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// rows from the neighbour list (classic NLIST)
//
// 128 registers / 2 blocks per SM on purpose: with fewer registers ptxas sinks the record loads of the next trip
// down to their first use (and spills the accumulators), which exposes the L2 latency of the gather in every trip.
// one block's worth of rows: rows [seg_begin + vb * rows_per_block, +rows_per_block), partial record vb
template <int K, int PBC, bool ACC, typename T>
__device__ __forceinline__ void sweep_list_block(const SweepArgs& a, const DevPbc& pbc, const DevSwitchT<T>& sw,
                                                 unsigned rows_per_block, unsigned seg_begin, unsigned seg_end, unsigned vb) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  LaneAcc acc = {0, 0, 0, 0, 0, 0, 0};
  unsigned long long evals = 0, execd = 0;
  const unsigned first = seg_begin + vb * rows_per_block;
  const unsigned last = min(first + rows_per_block, seg_end);
  unsigned fixmask = 0u;  // rows of this warp with a pair on a D_MAX / D_0 boundary (rows_per_block <= 32 warps' worth)
  const T far_skip2 = (T)a.far_skip2;
  const bool far_on = a.force_far || !(__longlong_as_double((long long)*a.disp2_bits) < a.far_disp2_max);
  // Row lookahead: a row's first list entries sit behind its metadata, both in HBM, and a near part is only ~4 loop
  // trips long -- so the metadata is requested two rows ahead and the first six entries per lane one row ahead.
  const unsigned k0 = first + wid;
  unsigned long long base1 = 0ull, base2 = 0ull;  // row k, row k + kSweepWarps
  unsigned cnt1 = 0u, cnt2 = 0u;
  uint32_t h0 = 0u, h1 = 0u, h2 = 0u, h3 = 0u, h4 = 0u, h5 = 0u;  // entries lane + 32 i of row k's near part
  if (k0 < last) {
    base1 = a.row_start[k0 - a.row_begin];
    cnt1 = a.row_count[k0 - a.row_begin];
    if (k0 + kSweepWarps < last) {
      base2 = a.row_start[k0 + kSweepWarps - a.row_begin];
      cnt2 = a.row_count[k0 + kSweepWarps - a.row_begin];
    }
    const uint32_t* __restrict__ r = a.nbr + base1 + lane;
    h0 = (lane < cnt1) ? __ldg(r) : 0u;
    h1 = (lane + 32 < cnt1) ? __ldg(r + 32) : 0u;
    h2 = (lane + 64 < cnt1) ? __ldg(r + 64) : 0u;
    h3 = (lane + 96 < cnt1) ? __ldg(r + 96) : 0u;
    h4 = (lane + 128 < cnt1) ? __ldg(r + 128) : 0u;
    h5 = (lane + 160 < cnt1) ? __ldg(r + 160) : 0u;
  }
  for (unsigned k = k0; k < last; k += kSweepWarps) {
    const SPos pi = load_spos(a.spos + k);
    const unsigned long long wi = ((unsigned long long)pi.slot << 32) | pi.abs_index;
    const unsigned long long base = base1;
    const unsigned cnt_near = cnt1, cnt_far = a.row_far_cnt[k - a.row_begin];
    const unsigned far_off = a.row_far_off[k - a.row_begin];
    const uint32_t j0 = h0, j1 = h1, j2 = h2, j3 = h3, j4 = h4, j5 = h5;
    {  // next row: its entries now (metadata is here), the metadata of the row after it
      const unsigned kn = k + kSweepWarps, knn = k + 2 * kSweepWarps;
      base1 = base2;
      cnt1 = cnt2;
      if (kn < last) {
        const uint32_t* __restrict__ r = a.nbr + base1 + lane;
        h0 = (lane < cnt1) ? __ldg(r) : 0u;
        h1 = (lane + 32 < cnt1) ? __ldg(r + 32) : 0u;
        h2 = (lane + 64 < cnt1) ? __ldg(r + 64) : 0u;
        h3 = (lane + 96 < cnt1) ? __ldg(r + 96) : 0u;
        h4 = (lane + 128 < cnt1) ? __ldg(r + 128) : 0u;
        h5 = (lane + 160 < cnt1) ? __ldg(r + 160) : 0u;
      }
      if (knn < last) {
        base2 = a.row_start[knn - a.row_begin];
        cnt2 = a.row_count[knn - a.row_begin];
      }
    }
    const bool row_is_b = (k >= a.n_a);
    const T qi = (K == K_DH) ? (T)__ldg(a.sq + k) : T(1.0);
    const uint32_t ti = (K == K_GHB) ? __ldg(a.stype + k) : 0u;
    T fx = T(0.0), fy = T(0.0), fz = T(0.0);
    LaneAccT<T> row_sums = {T(0.0), T(0.0), T(0.0), T(0.0), T(0.0), T(0.0), T(0.0)};  // FP32 mode only
    bool near = false;
    // Two pairs per lane and trip, evaluated as two independent straight-line chains: one pair is a ~25-deep
    // chain of dependent FP64 operations, so a single chain per warp leaves the FP64 pipe half idle.
    auto part = [&](const uint32_t* __restrict__ row, unsigned cnt, auto far_tag) {
      constexpr bool FAR = decltype(far_tag)::value;  // the near part's first six entries were prefetched (j0..j5)
      unsigned e = lane;
      uint32_t ja = FAR ? ((e < cnt) ? __ldg(row + e) : 0u) : j0;
      uint32_t jb = FAR ? ((e + 32 < cnt) ? __ldg(row + e + 32) : 0u) : j1;
      RecBuf pa, pb;
      load_rec(a.spos + (ja & a.idx_mask), pa);
      load_rec(a.spos + (jb & a.idx_mask), pb);
      uint32_t ia = ja, ib = jb;  // entries of the records in flight (DHENERGY / GHBFIX look charges / types up)
      // list entries run two loop trips ahead of the records (HBM latency), the records one trip ahead of the math
      ja = FAR ? ((e + 64 < cnt) ? __ldg(row + e + 64) : 0u) : j2;
      jb = FAR ? ((e + 96 < cnt) ? __ldg(row + e + 96) : 0u) : j3;
      uint32_t na = FAR ? ((e + 128 < cnt) ? __ldg(row + e + 128) : 0u) : j4;
      uint32_t nb = FAR ? ((e + 160 < cnt) ? __ldg(row + e + 160) : 0u) : j5;
      for (unsigned e0 = 0; e0 < cnt; e0 += 64, e += 64) {  // warp-uniform trip count: the far part votes
        const RecBuf ca = pa, cb = pb;
        T qqa = T(1.0), qqb = T(1.0);
        if (K == K_DH || K == K_GHB) {
          if (K == K_DH) {
            qqa = qi * (T)__ldg(a.sq + (ia & a.idx_mask));
            qqb = qi * (T)__ldg(a.sq + (ib & a.idx_mask));
          } else {  // eta[type of the pair's first atom][type of its second atom], GHBFIX.cpp:189-197
            const uint32_t ta = __ldg(a.stype + (ia & a.idx_mask)), tb = __ldg(a.stype + (ib & a.idx_mask));
            const bool fa = a.two_groups ? row_is_b : (wi > (unsigned long long)__double_as_longlong(ca.w));
            const bool fb = a.two_groups ? row_is_b : (wi > (unsigned long long)__double_as_longlong(cb.w));
            qqa = (T)__ldg(a.etas + (fa ? ta * a.ntypes + ti : ti * a.ntypes + ta));
            qqb = (T)__ldg(a.etas + (fb ? tb * a.ntypes + ti : ti * a.ntypes + tb));
          }
          ia = ja;
          ib = jb;
        }
        load_rec(a.spos + (ja & a.idx_mask), pa);
        load_rec(a.spos + (jb & a.idx_mask), pb);
        ja = na;
        jb = nb;
        na = (e + 192 < cnt) ? __ldg(row + e + 192) : 0u;
        nb = (e + 224 < cnt) ? __ldg(row + e + 224) : 0u;
        pair_term2<K, PBC, ACC, FAR>(pbc, sw, near, pi.x, pi.y, pi.z, wi, a.two_groups, row_is_b, ca, cb, e < cnt,
                                     e + 32 < cnt, far_skip2, fx, fy, fz, row_acc(acc, row_sums), qqa, qqb);
      }
    };
    if (cnt_near) part(a.nbr + base, cnt_near, std::false_type{});
    if (cnt_far && far_on) part(a.nbr + base + far_off, cnt_far, std::true_type{});
    // a pair of this row sits on a D_MAX / D_0 boundary: the row is patched after the loop (the cold call is kept
    // out of it so that nothing is spilled around it)
    if (__any_sync(0xffffffffu, near)) fixmask |= 1u << ((k - first) / kSweepWarps);
    if (ACC) flush_row_acc(acc, row_sums);
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (lane == 0) {
      a.sderiv[3 * (size_t)k] = (double)fx;
      a.sderiv[3 * (size_t)k + 1] = (double)fy;
      a.sderiv[3 * (size_t)k + 2] = (double)fz;
      evals += cnt_near + cnt_far;
      execd += cnt_near + (far_on ? cnt_far : 0u);
    }
  }
  if (lane == 0 && execd) atomicAdd(a.executed, execd);
  if constexpr (std::is_same<T, double>::value)
  while (fixmask) {
    const unsigned m = (unsigned)__ffs((int)fixmask) - 1u;
    fixmask &= fixmask - 1u;
    const unsigned kf = first + wid + kSweepWarps * m;
    const uint32_t* __restrict__ row = a.nbr + a.row_start[kf - a.row_begin];
    const RowFix f = row_fixup_list<K, PBC>(a.pbc_g, a.sw_g, a.spos, a.pos, a.idx_mask, row, a.row_count[kf - a.row_begin],
                                            row + a.row_far_off[kf - a.row_begin], a.row_far_cnt[kf - a.row_begin], kf, lane,
                                            a.two_groups, kf >= a.n_a);
    double gx = 0.0, gy = 0.0, gz = 0.0;
    apply_fix(f, ACC, gx, gy, gz, acc);
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    gz = warp_sum(gz);
    if (lane == 0) {  // the same lane stored these three values above
      a.sderiv[3 * (size_t)kf] += gx;
      a.sderiv[3 * (size_t)kf + 1] += gy;
      a.sderiv[3 * (size_t)kf + 2] += gz;
    }
  }
  if (ACC)
    block_store_partials(acc, evals, a.partials, a.evals, vb);
  else if (lane == 0 && evals)
    atomicAdd(a.evals, evals);
}

// In image mode this kernel is launched every step next to k_sweep_img and one of the two returns at once (device-side
// gate on the displacement).  The one that returns should cost nothing: the launch then has only as many blocks as
// fit the machine at once, and a block that does run walks its share of the `nvb` virtual blocks.
template <int K, int PBC, bool ACC, typename T>
__global__ void __launch_bounds__(kSweepThreads, 2)
    k_sweep_list(SweepArgs a, DevPbc pbc, DevSwitchT<T> sw, unsigned rows_per_block, unsigned seg_begin, unsigned seg_end,
                 unsigned nvb) {
  // image mode: k_sweep_img does the step while its displacement bound holds (img_disp2_max = 0: it never runs)
  if (__longlong_as_double((long long)*a.disp2_bits) < a.img_disp2_max) return;
  for (unsigned vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
    sweep_list_block<K, PBC, ACC, T>(a, pbc, sw, rows_per_block, seg_begin, seg_end, vb);
    __syncthreads();  // the shared arrays of the block epilogue are reused by the next virtual block
  }
}

// ------------------------------------------------------------------------------------------------
// rows from the sorted ranges of the stencil cells (NLISTCELLS superset, or a single 1x1x1 "cell" = no NL)
template <int K, int PBC, bool ACC, typename T>
__global__ void __launch_bounds__(kSweepThreads)
    k_sweep_cells(SweepArgs a, DevPbc pbc, DevSwitchT<T> sw, unsigned rows_per_block, unsigned seg_begin, unsigned seg_end) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  LaneAcc acc = {0, 0, 0, 0, 0, 0, 0};
  unsigned long long evals = 0;
  const DevGrid& g = a.grid;
  const unsigned first = seg_begin + blockIdx.x * rows_per_block;
  const unsigned last = min(first + rows_per_block, seg_end);
  for (unsigned k = first + wid; k < last; k += kSweepWarps) {
    const SPos pi = load_spos(a.spos + k);
    const unsigned my_grp = (k < a.n_a) ? 0u : 1u;
    const unsigned other = a.two_groups ? (1u - my_grp) : 0u;
    int c[3];
    cell_coords(g, (int)a.scell[k], c);
    const T qi = (K == K_DH) ? (T)__ldg(a.sq + k) : T(1.0);
    const uint32_t ti = (K == K_GHB) ? __ldg(a.stype + k) : 0u;
    T fx = T(0.0), fy = T(0.0), fz = T(0.0);
    LaneAccT<T> row_sums = {T(0.0), T(0.0), T(0.0), T(0.0), T(0.0), T(0.0), T(0.0)};  // FP32 mode only
    unsigned cnt = 0;
    bool unused_near = false;
    for_each_stencil_range(g, c, other * (unsigned)g.ncell, a.cstart, a.ccount, [&](uint32_t s0, uint32_t m, int, int, int) {
      cnt += m;
#pragma unroll 2
      for (uint32_t e = lane; e < m; e += 32) {
        const uint32_t j = s0 + e;
        const SPos pj = load_spos(a.spos + j);
        const bool valid = (j != k) && (!a.check_abs || pj.abs_index != pi.abs_index);
        const bool flip = a.two_groups ? (my_grp == 1u) : (pi.slot > pj.slot);
        if (valid) {
          T qq = T(1.0);
          if (K == K_DH) qq = qi * (T)__ldg(a.sq + j);
          if (K == K_GHB) {
            const uint32_t tj = __ldg(a.stype + j);
            qq = (T)__ldg(a.etas + (flip ? tj * a.ntypes + ti : ti * a.ntypes + tj));
          }
          pair_term<K, PBC, ACC, true>(pbc, sw, unused_near, pi.x, pi.y, pi.z, pj, flip, fx, fy, fz, row_acc(acc, row_sums), qq);
        }
      }
    });
    if (ACC) flush_row_acc(acc, row_sums);
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (lane == 0) {
      a.sderiv[3 * (size_t)k] = (double)fx;
      a.sderiv[3 * (size_t)k + 1] = (double)fy;
      a.sderiv[3 * (size_t)k + 2] = (double)fz;
      evals += cnt;
    }
  }
  if (lane == 0 && evals) atomicAdd(a.executed, evals);
  if (ACC)
    block_store_partials(acc, evals, a.partials, a.evals);
  else if (lane == 0 && evals)
    atomicAdd(a.evals, evals);
}

// ------------------------------------------------------------------------------------------------
// dispatch
// blocks to launch for `nblocks` virtual blocks: all of them normally; in image mode (the kernel is the fall-back that
// usually returns at once) one wave, 2 resident blocks on each of the 148 SMs
static inline int list_grid(const SweepArgs& a, int nblocks) {
  return (a.img_disp2_max > 0.0 && nblocks > 296) ? 296 : nblocks;
}

template <int K, int PBC, bool LIST, typename T>
static int run_sweep(const SweepArgs& a, const DevPbc& pbc, const DevSwitchT<T>& sw, cudaStream_t st) {
  // rows that accumulate value+virial: SingleList -> all; TwoList -> only the A rows
  const unsigned acc_end = a.two_groups ? min(a.row_end, a.n_a) : a.row_end;
  int nblocks = 0;
  if (a.row_begin < acc_end) {
    const unsigned rows = acc_end - a.row_begin;
    const unsigned rpb = a.rows_per_block ? a.rows_per_block : pick_rows_per_block(rows);
    nblocks = (int)((rows + rpb - 1) / rpb);
    if (LIST)
      k_sweep_list<K, PBC, true, T><<<list_grid(a, nblocks), kSweepThreads, 0, st>>>(a, pbc, sw, rpb, a.row_begin, acc_end,
                                                                                      (unsigned)nblocks);
    else
      k_sweep_cells<K, PBC, true, T><<<nblocks, kSweepThreads, 0, st>>>(a, pbc, sw, rpb, a.row_begin, acc_end);
  }
  const unsigned b_begin = max(a.row_begin, acc_end);
  if (b_begin < a.row_end) {
    const unsigned rows = a.row_end - b_begin;
    const unsigned rpb = a.rows_per_block ? a.rows_per_block : pick_rows_per_block(rows);
    const int nb = (int)((rows + rpb - 1) / rpb);
    if (LIST)
      k_sweep_list<K, PBC, false, T><<<list_grid(a, nb), kSweepThreads, 0, st>>>(a, pbc, sw, rpb, b_begin, a.row_end, (unsigned)nb);
    else
      k_sweep_cells<K, PBC, false, T><<<nb, kSweepThreads, 0, st>>>(a, pbc, sw, rpb, b_begin, a.row_end);
  }
  return nblocks;
}

template <int K, bool LIST, typename T>
static int run_sweep_pbc(const SweepArgs& a, const DevPbc& pbc, const DevSwitchT<T>& sw, cudaStream_t st) {
  switch (pbc.type) {
    case 0: return run_sweep<K, 0, LIST, T>(a, pbc, sw, st);
    case 1: return run_sweep<K, 1, LIST, T>(a, pbc, sw, st);
    default: return run_sweep<K, 2, LIST, T>(a, pbc, sw, st);
  }
}

template <bool LIST, typename T>
static int run_sweep_kind(const SweepArgs& a, const DevPbc& pbc, const DevSwitchT<T>& sw, cudaStream_t st) {
  switch (kind_of(sw.type)) {
    case K_FIX6: return run_sweep_pbc<K_FIX6, LIST, T>(a, pbc, sw, st);
    case K_FIXN: return run_sweep_pbc<K_FIXN, LIST, T>(a, pbc, sw, st);
    case K_RAT_R2: return run_sweep_pbc<K_RAT_R2, LIST, T>(a, pbc, sw, st);
    case K_RAT_R: return run_sweep_pbc<K_RAT_R, LIST, T>(a, pbc, sw, st);
    case K_EXP: return run_sweep_pbc<K_EXP, LIST, T>(a, pbc, sw, st);
    case K_GAUSS: return run_sweep_pbc<K_GAUSS, LIST, T>(a, pbc, sw, st);
    case K_FASTGAUSS: return run_sweep_pbc<K_FASTGAUSS, LIST, T>(a, pbc, sw, st);
    case K_SMAP: return run_sweep_pbc<K_SMAP, LIST, T>(a, pbc, sw, st);
    case K_CUBIC: return run_sweep_pbc<K_CUBIC, LIST, T>(a, pbc, sw, st);
    case K_TANH: return run_sweep_pbc<K_TANH, LIST, T>(a, pbc, sw, st);
    case K_COS: return run_sweep_pbc<K_COS, LIST, T>(a, pbc, sw, st);
    case K_NATIVEQ: return run_sweep_pbc<K_NATIVEQ, LIST, T>(a, pbc, sw, st);
    case K_DH: return run_sweep_pbc<K_DH, LIST, T>(a, pbc, sw, st);
    case K_GHB: return run_sweep_pbc<K_GHB, LIST, T>(a, pbc, sw, st);
    default: return -1;
  }
}

}  // namespace b200
