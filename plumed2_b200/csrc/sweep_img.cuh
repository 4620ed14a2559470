// The list sweep on CONTINUOUS coordinates with the periodic image stored in the list ("image mode").
//
// What the reference does per pair (CoordinationBase.cpp:177-208): distance = pbcDistance(pos[i0], pos[i1]) -- a full
// minimum-image search -- then the switching function, then  deriv[i0] -= dd, deriv[i1] += dd, virial -= dd (x) d.
// Here the minimum-image search has already happened: when the neighbour list is built, every entry records the
// periodic image (wx,wy,wz) its partner was found through (6 bits above the 26-bit sorted index), and the per-step
// gather (k_gather_u) writes continuous coordinates u = wrapped position at the sort + minimum-image displacement
// since.  The vector of a listed pair is then  e = u_j - u_i + S[image]  -- three subtractions, and three more
// additions on the few trips that touch a box face.  It equals the reference's minimum image while
// NL_CUTOFF + 2 * (largest displacement since the rebuild) < half the smallest box height (the image of a pair that
// close is unique); the kernel checks that bound on the device and leaves the step to the general kernel
// (k_sweep_list, minimum image per pair) otherwise.
//
// Virial from positions.  With e = u_j - u_i + S:
//   -sum_pairs dd (x) d  =  -sum_atoms deriv_a (x) u_a  -  w * sum_rows sum_entries g (x) S,     g = df * e,
// w = 1/2 when every pair is seen from both ends (one group), 1 when only GROUPA rows count (two groups).
// The first sum costs nine FMAs per ROW; the second only runs on trips with a shifted entry.  That removes ten of
// the ~34 FP64 operations per pair end.
//
// One flat software pipeline per warp across rows and row parts: list entries are requested two trips ahead
// (HBM stream), partner records one trip ahead (L1/L2 gather), the next row's own record and metadata earlier
// still -- a row is only ~4 trips long, so a per-row pipeline would expose two memory round trips per row.
#pragma once
#include <cstdlib>
#include <type_traits>

#include "sweep_math.cuh"

namespace b200 {

struct ImgShifts {
  double v[64][3];  // indexed by the xor-encoded image code (kernels.cuh: super_image); 27 of 64 used
};

// rationalfixN: value = stretch * res + shift.  Accumulating res and counting the pairs inside D_MAX replaces one
// FP64 FMA per pair by one predicated integer add.
template <int K>
struct Unstretched {
  static constexpr bool value = (K == K_FIX6 || K == K_FIXN);
};

// 1/a to full precision: MUFU seed (~2^-20) and ONE cubic step, x1 = x0 (1 + e + e^2), error e^3
__device__ __forceinline__ double rcp_cubic(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  const double e = fma(-a, x, 1.0);
  const double p = fma(e, e, e);
  return fma(x, p, x);
}

// record `idx` of the sorted array: address = one IMAD.WIDE (index * 32 + base), one 256-bit load
__device__ __forceinline__ void load_rec_idx(const SPos* __restrict__ base, uint32_t idx, RecBuf& b) {
  asm volatile(
      "{\n\t.reg .u64 p;\n\tmad.wide.u32 p, %4, 32, %5;\n\tld.global.nc.v4.f64 {%0,%1,%2,%3}, [p];\n\t}"
      : "=d"(b.x), "=d"(b.y), "=d"(b.z), "=d"(b.w)
      : "r"(idx), "l"(base));
}

// (s or res, df) of one pair; ok = listed entry (not padding).  Branch-free for the rationalfix kinds.
template <int K>
__device__ __forceinline__ void img_eval(const DevSwitch& sw, double r2, bool ok, double& s, double& df, unsigned& nin) {
  if (K == K_FIX6 || K == K_FIXN) {
    const double y = r2 * sw.invr0_2;
    const double t = (K == K_FIX6) ? y * y : ipow_dev(y, sw.nnf - 1);
    double res = rcp_cubic(fma(t, y, 1.0));
    const bool in = ok && (r2 <= sw.dmax_2);
    res = in ? res : 0.0;
    nin += in ? 1u : 0u;
    df = (t * res) * (res * sw.fix_df);
    s = res;
  } else {
    eval_switch<K>(sw, r2, s, df);
    if (!ok) {
      s = 0.0;
      df = 0.0;
    }
  }
}

// coarse "is r^2 near a boundary of the switching function" on the high word of r^2 (integer pipe): the buckets
// around D_MAX^2 and D_0^2 are ~1e-6 wide; the exact patch re-checks with its 1e-10 band.
struct NearBands {
  uint32_t lo[2], span[2];  // high words; span < 0x80000000.  Unused band: lo = 0xffffffff, span = 0
};
__device__ __forceinline__ bool near_band0(const NearBands& nb, double r2) {
  return ((uint32_t)__double2hiint(r2) - nb.lo[0]) <= nb.span[0];
}
__device__ __forceinline__ bool near_band1(const NearBands& nb, double r2) {
  return ((uint32_t)__double2hiint(r2) - nb.lo[1]) <= nb.span[1];
}

// ------------------------------------------------------------------------------------------------
// exact patch of one row (cold; see sweep_math.cuh for why it exists).  For every pair of the row within 1e-10 of a
// boundary: (exact contribution) - (what the hot loop added).
struct RowFixImg {
  double fx, fy, fz, val, c[9];
};
template <int K>
__device__ __noinline__ RowFixImg row_fixup_img(const DevPbc* __restrict__ pbc_g, const DevSwitch* __restrict__ sw_g,
                                                const SPos* __restrict__ spos, PosSrc pos,
                                                const uint32_t* __restrict__ nbr, uint4 m, bool far_on,
                                                const double* shifts /*shared*/, unsigned k, unsigned lane, int two_groups,
                                                bool row_is_b, bool acc, double* scatter /*null or the derivative rows*/) {
  RowFixImg f;
  f.fx = f.fy = f.fz = f.val = 0.0;
#pragma unroll
  for (int q = 0; q < 9; ++q) f.c[q] = 0.0;
  const DevSwitch& sw = *sw_g;
  const SPos pi = spos[k];
  const uint32_t* __restrict__ row = nbr + 4ull * m.x;
  const unsigned total = m.y + (far_on ? m.w : 0u);
  for (unsigned e = lane; e < total; e += 32) {
    const uint32_t ent = (e < m.y) ? row[e] : row[m.z + (e - m.y)];
    const SPos pj = spos[ent & kSuperIndexMask];
    const double* S = shifts + 3 * (ent >> 26);
    const double ex = (pj.x - pi.x) + S[0], ey = (pj.y - pi.y) + S[1], ez = (pj.z - pi.z) + S[2];
    const double r2 = fma(ez, ez, fma(ey, ey, ex * ex));
    if (!on_boundary(sw, r2)) continue;
    double s, df;
    eval_switch<K>(sw, r2, s, df);  // what the hot loop added (same formulas; rounding differences are 1e-16)
    const bool flip = two_groups ? row_is_b : (pi.slot > pj.slot);
    const double* ri = pos_at(pos, pi.slot);
    const double* rj = pos_at(pos, pj.slot);
    const ExactPair o = exact_pair<K>(*pbc_g, sw, ri[0], ri[1], ri[2], rj[0], rj[1], rj[2], flip);
    // the exact vector in this row's orientation (from i to the image of j), g = df * e
    const double sg = flip ? -1.0 : 1.0;
    const double gx = o.df * sg * o.dx - df * ex, gy = o.df * sg * o.dy - df * ey, gz = o.df * sg * o.dz - df * ez;
    f.fx -= gx;
    f.fy -= gy;
    f.fz -= gz;
    if (scatter) {  // the partner's row is not swept: it gets its share of the correction here
      double* dj = scatter + 3 * (size_t)(ent & kSuperIndexMask);
      atomicAdd(dj, gx);
      atomicAdd(dj + 1, gy);
      atomicAdd(dj + 2, gz);
    }
    if (acc) {
      f.val += o.s - s;
      f.c[0] += gx * S[0]; f.c[1] += gx * S[1]; f.c[2] += gx * S[2];
      f.c[3] += gy * S[0]; f.c[4] += gy * S[1]; f.c[5] += gy * S[2];
      f.c[6] += gz * S[0]; f.c[7] += gz * S[1]; f.c[8] += gz * S[2];
    }
  }
  return f;
}

// ------------------------------------------------------------------------------------------------
// Registers are what limits the resident warps of this kernel, so everything that is not touched by every pair
// lives in shared memory: the block's own records and row metadata (staged once, coalesced), the 27 shift vectors,
// and the per-thread sums of g (x) S (only trips with a shifted entry touch them) and of the patch values.
// The position part of the virial needs one register pair: lane q < 9 accumulates component q of deriv (x) u.
constexpr int kImgMaxRows = 128;  // rows per block (pick_rows_per_block never exceeds it)
constexpr int kImgTripCap = 208;  // trips per warp the table holds: 16 rows x 13 trips (img_rows_per_block adapts)
constexpr unsigned kImgMaxRow = 65535u;  // longest row the 16-bit count of a trip descriptor can describe

template <int K, bool ACC, int MINB>
__global__ void __launch_bounds__(kSweepThreads, 2)
    k_sweep_img(SweepArgs a, DevSwitch sw, ImgShifts sh, NearBands nb, unsigned rows_per_block, unsigned seg_begin,
                unsigned seg_end) {
  const double disp2 = __longlong_as_double((long long)*a.disp2_bits);
  if (!(disp2 < a.img_disp2_max)) return;  // images may no longer be the minimum images: k_sweep_list takes the step
  const bool far_on = a.force_far || !(disp2 < a.far_disp2_max);
  __shared__ double s_shift[64 * 3];
  __shared__ double s_c[10][kSweepThreads];  // [0..8] g (x) S, [9] value corrections of the exact patch
  __shared__ RecBuf s_own[kImgMaxRows];      // the block's own records {u, slot:abs bits}
  __shared__ uint4 s_meta[kImgMaxRows];      // {row start / 4, near count, far offset, far count}
  const unsigned first = seg_begin + blockIdx.x * rows_per_block;
  const unsigned last = min(first + rows_per_block, seg_end);
  for (unsigned t = threadIdx.x; t < 192u; t += kSweepThreads) s_shift[t] = (&sh.v[0][0])[t];
#pragma unroll
  for (int q = 0; q < 10; ++q) s_c[q][threadIdx.x] = 0.0;
  for (unsigned t = threadIdx.x; t < last - first; t += kSweepThreads) {
    RecBuf own;
    load_rec(a.spos + first + t, own);
    s_own[t] = own;
    s_meta[t] = __ldg(a.row_meta + (first + t - a.row_begin));
  }
  __syncthreads();

  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned nrows = last - first;
  const double winv = a.two_groups ? 1.0 : 2.0;  // 1 / w

  double val = 0.0;    // sum of s (Unstretched kinds: of the unstretched value)
  double cpos = 0.0;   // lane q < 9: component q of sum (deriv / w) (x) u
  unsigned nin = 0u;   // pairs inside D_MAX (Unstretched kinds)
  unsigned executed = 0u, fixmask = 0u;

  // ---- trip table.  A warp owns the local rows wid, wid+8, ... (at most 16); per row the near part and, when visited,
  // the far part, each cut into trips of 64 entries.  The table is built once per block (lane l lists the trips of the
  // warp's l-th row), so the hot loop's "iterator" is one 8-byte shared-memory load: no branches, no state.
  //   .x = first entry of the trip relative to the block's first row, .y = entries left in the part | row slot << 16 | flags
  constexpr unsigned kOk = 0x80000000u, kFar = 0x40000000u, kRem = 0xffffu;
  __shared__ uint2 s_trip[kSweepWarps][kImgTripCap + 4];
  const uint32_t base4 = nrows ? s_meta[0].x : 0u;
  const uint32_t* __restrict__ blk = a.nbr + 4ull * base4 + lane;  // rows of a block are stored in order
  {
    const unsigned rl = wid + kSweepWarps * lane;  // this lane's row of the warp (lanes >= 16 never have one)
    uint4 m = make_uint4(0u, 0u, 0u, 0u);
    if (lane < 16u && rl < nrows) m = s_meta[rl];
    const unsigned tn = (m.y + 63u) >> 6, tf = far_on ? ((m.w + 63u) >> 6) : 0u;
    uint32_t total;
    unsigned at = warp_exclusive_scan(tn + tf, lane, total);
    const uint32_t off0 = 4u * (m.x - base4);
    for (unsigned t = 0; t < tn; ++t) s_trip[wid][at++] = make_uint2(off0 + 64u * t, (m.y - 64u * t) | (lane << 16) | kOk);
    for (unsigned t = 0; t < tf; ++t)
      s_trip[wid][at++] = make_uint2(off0 + m.z + 64u * t, (m.w - 64u * t) | (lane << 16) | kOk | kFar);
    if (lane < 4u) s_trip[wid][total + lane] = make_uint2(0u, 0u);  // the pipeline reads up to four trips past the end
    unsigned mine = m.y + (far_on ? m.w : 0u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    executed = mine;
    __syncwarp();
  }
  unsigned it = 0u;  // next trip of the table
  uint2 it_d = s_trip[wid][0];
  // entries of the table's next trip (0 = sorted atom 0 through the home image: a valid record, masked out later)
  auto entry = [&](unsigned off) -> uint32_t { return (off + lane < (it_d.y & kRem)) ? __ldg(blk + it_d.x + off) : 0u; };

  // Two register sets, A and B, alternate: while the trip of one set is evaluated, the records of the other set's
  // trip are in flight and the first set's entry registers are refilled with the trip after that.  No copies.
  struct Set {
    uint32_t ea, eb;   // entries (lane, lane + 32) of the set's trip
    unsigned m;        // entries left in its part | row slot << 16 | flags
    RecBuf pa, pb;     // partner records
    double qa, qb;     // DHENERGY: their charges
    uint32_t ta, tb;   // GHBFIX: their types
  };
  auto request = [&](Set& s) {
    load_rec_idx(a.spos, s.ea & kSuperIndexMask, s.pa);
    load_rec_idx(a.spos, s.eb & kSuperIndexMask, s.pb);
    if (K == K_DH) {
      s.qa = __ldg(a.sq + (s.ea & kSuperIndexMask));
      s.qb = __ldg(a.sq + (s.eb & kSuperIndexMask));
    }
    if (K == K_GHB) {
      s.ta = __ldg(a.stype + (s.ea & kSuperIndexMask));
      s.tb = __ldg(a.stype + (s.eb & kSuperIndexMask));
    }
  };
  auto refill = [&](Set& s) {  // descriptor and entries of the table's next trip
    s.m = it_d.y;
    s.ea = entry(0u);
    s.eb = entry(32u);
    it_d = s_trip[wid][++it];
    // The entries stream from HBM, one 256-byte piece per warp and trip from ~2400 places at once; measured, a trip
    // (~1300 cycles) is not always enough for that round trip.  Ask L2 for the piece of the trip two further down
    // the table (8 bytes per lane cover its 256 bytes); the load above then finds its piece in L2.
    const uint2 pd = s_trip[wid][it + 2];
    asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + pd.x + lane));
  };
  Set A, B;
  A.qa = A.qb = B.qa = B.qb = 1.0;
  A.ta = A.tb = B.ta = B.tb = 0u;
  refill(A);
  refill(B);
  request(A);

  unsigned cur_r = 0xffffffffu;
  double xi = 0.0, yi = 0.0, zi = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
  unsigned long long wi = 0ull;
  double qi = 1.0;
  uint32_t ti = 0u;
  uint32_t near0 = 0xffffffffu, near1 = 0xffffffffu;  // smallest (high word of r^2) - (start of the band) of the row
  bool row_is_b = false;
  const bool two_bands = (nb.lo[1] != 0xffffffffu);

  // Row epilogue.  The three sums are reduced together: after two exchange rounds every lane owns ONE of them
  // (lanes with bits 4,3 = 00: x, 10: y, x1: z), three more rounds finish -- 6 shuffles instead of 15.  Lanes 0,1,2 /
  // 16,17,18 / 8,9,10 then add their component of (deriv / w) (x) u; lanes 0, 16, 8 store the row.
  const bool b4 = (lane & 16u) != 0u, b3 = (lane & 8u) != 0u;
  const unsigned col = lane & 7u;
  auto row_done = [&]() {
    const bool flagged = (near0 <= nb.span[0]) || (two_bands && near1 <= nb.span[1]);
    if (__any_sync(0xffffffffu, flagged)) fixmask |= 1u << (cur_r / kSweepWarps);
    double keep = b4 ? fy : fx;
    keep += __shfl_xor_sync(0xffffffffu, b4 ? fx : fy, 16);
    const double z = fz + __shfl_xor_sync(0xffffffffu, fz, 16);
    double own = b3 ? z : keep;
    own += __shfl_xor_sync(0xffffffffu, b3 ? keep : z, 8);
    own += __shfl_xor_sync(0xffffffffu, own, 4);
    own += __shfl_xor_sync(0xffffffffu, own, 2);
    own += __shfl_xor_sync(0xffffffffu, own, 1);
    const size_t k = (size_t)first + cur_r;
    if (col == 0u && lane < 24u) a.sderiv[3 * k + (b3 ? 2u : (b4 ? 1u : 0u))] = own;
    const double uq = (col == 0u) ? xi : ((col == 1u) ? yi : zi);
    if (col < 3u) cpos = fma(own * winv, uq, cpos);
  };

  // one trip: `cur` holds its entries, descriptor and (arrived) records, `nxt` the entries of the next trip
  auto step = [&](Set& cur, Set& nxt) {
    const uint32_t ia = cur.ea, ib = cur.eb;
    const unsigned m = cur.m;
    const unsigned r = wid + kSweepWarps * ((m >> 16) & 15u);
    if (r != cur_r) {  // a new row starts with this trip
      if (cur_r != 0xffffffffu) row_done();
      cur_r = r;
      const RecBuf own = s_own[r];
      xi = own.x;
      yi = own.y;
      zi = own.z;
      wi = (unsigned long long)__double_as_longlong(own.w);
      fx = fy = fz = 0.0;
      near0 = near1 = 0xffffffffu;
      row_is_b = (first + r >= a.n_a);
      if (K == K_DH) qi = __ldg(a.sq + first + r);
      if (K == K_GHB) ti = __ldg(a.stype + first + r);
    }
    // Order matters: a warp has six scoreboards for its loads in flight, so a wait for an old load also waits for
    // any younger load that shares its scoreboard.  Everything this trip has to WAIT for -- its records, then the
    // entries of the next trip -- is therefore consumed before this trip ISSUES anything new.
    double ax = cur.pa.x - xi, ay = cur.pa.y - yi, az = cur.pa.z - zi;
    double bx = cur.pb.x - xi, by = cur.pb.y - yi, bz = cur.pb.z - zi;
    const unsigned long long wa = (unsigned long long)__double_as_longlong(cur.pa.w);
    const unsigned long long wb = (unsigned long long)__double_as_longlong(cur.pb.w);
    const double qa = cur.qa, qb = cur.qb;
    const uint32_t ta = cur.ta, tb = cur.tb;
    // The issue side of the pipeline: records of the next trip (L1/L2 gather, one trip ahead), entries of the trip
    // after next (HBM stream, two trips ahead).  ~35 integer / load instructions that depend on nothing computed here.
    auto issue_side = [&]() {
      request(nxt);
      refill(cur);
    };
    // Measured (1 M atoms): issuing up front 1.075 ms, issuing between the two FP64 chains (variant 3, A/B switch
    // B200COORD_IMG_VARIANT=3) 1.126 ms -- the interleaved loads get their turn later and the next trip waits for them.
    constexpr bool kLateIssue = (MINB == 3);
    if (!kLateIssue) issue_side();
    const unsigned rem = m & kRem;
    const bool va = lane < rem, vb = lane + 32u < rem;
    // everything after the vector: same for both flavours of the trip
    auto finish = [&](auto shifted_tag) {
      constexpr bool SHIFTED = decltype(shifted_tag)::value;
      const double ra = fma(az, az, fma(ay, ay, ax * ax));
      const double rb = fma(bz, bz, fma(by, by, bx * bx));
      if (m & kFar) {  // far part: every pair of the trip beyond D_MAX -> nothing to add
        if (__all_sync(0xffffffffu, (!va || ra > a.far_skip2) && (!vb || rb > a.far_skip2))) {
          if (kLateIssue) issue_side();
          return;
        }
      }
      if (kLateIssue) issue_side();
      double sa, dfa, sb, dfb;
      img_eval<K>(sw, ra, va, sa, dfa, nin);
      img_eval<K>(sw, rb, vb, sb, dfb, nin);
      if (K == K_DH) {
        const double qqa = qi * qa, qqb = qi * qb;
        sa *= qqa; dfa *= qqa;
        sb *= qqb; dfb *= qqb;
      }
      if (K == K_GHB) {  // eta[type of the pair's first atom][type of its second atom], GHBFIX.cpp:189-197
        const bool fa = a.two_groups ? row_is_b : (wi > wa);
        const bool fb = a.two_groups ? row_is_b : (wi > wb);
        const double qqa = __ldg(a.etas + (fa ? ta * a.ntypes + ti : ti * a.ntypes + ta));
        const double qqb = __ldg(a.etas + (fb ? tb * a.ntypes + ti : ti * a.ntypes + tb));
        sa *= qqa; dfa *= qqa;
        sb *= qqb; dfb *= qqb;
      }
      // padding lanes take part too: a false alarm only makes the (exact) patch look at the row
      near0 = min(near0, min((uint32_t)__double2hiint(ra) - nb.lo[0], (uint32_t)__double2hiint(rb) - nb.lo[0]));
      if (two_bands)
        near1 = min(near1, min((uint32_t)__double2hiint(ra) - nb.lo[1], (uint32_t)__double2hiint(rb) - nb.lo[1]));
      fx = fma(-dfa, ax, fma(-dfb, bx, fx));
      fy = fma(-dfa, ay, fma(-dfb, by, fy));
      fz = fma(-dfa, az, fma(-dfb, bz, fz));
      if (ACC && a.scatter_b) {
        // few GROUPA atoms in a sea of GROUPB atoms: a GROUPB row holds a handful of entries, and sweeping a million of
        // them costs ten times the GROUPA rows.  Instead this (GROUPA) row hands +dd to its partners: deriv[i1] += dd
        // (CoordinationBase.cpp:201), three reductions per contributing pair, no second evaluation.
        if (dfa != 0.0) {
          double* dj = a.sderiv + 3 * (size_t)(ia & kSuperIndexMask);
          atomicAdd(dj, dfa * ax);
          atomicAdd(dj + 1, dfa * ay);
          atomicAdd(dj + 2, dfa * az);
        }
        if (dfb != 0.0) {
          double* dj = a.sderiv + 3 * (size_t)(ib & kSuperIndexMask);
          atomicAdd(dj, dfb * bx);
          atomicAdd(dj + 1, dfb * by);
          atomicAdd(dj + 2, dfb * bz);
        }
      }
      if (ACC) {
        val += sa + sb;
        if (SHIFTED) {  // g (x) S of this trip, g = df * e
          const double* Sa = s_shift + 3 * (ia >> 26);
          const double* Sb = s_shift + 3 * (ib >> 26);
          const double g[3][2] = {{dfa * ax, dfb * bx}, {dfa * ay, dfb * by}, {dfa * az, dfb * bz}};
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
              s_c[3 * rr + cc][threadIdx.x] = fma(g[rr][0], Sa[cc], fma(g[rr][1], Sb[cc], s_c[3 * rr + cc][threadIdx.x]));
        }
      }
    };
    if (__any_sync(0xffffffffu, (ia | ib) > kSuperIndexMask)) {  // some partner of this trip sits across a box face
      const double* Sa = s_shift + 3 * (ia >> 26);
      const double* Sb = s_shift + 3 * (ib >> 26);
      ax += Sa[0]; ay += Sa[1]; az += Sa[2];
      bx += Sb[0]; by += Sb[1]; bz += Sb[2];
      finish(std::true_type{});
    } else {
      finish(std::false_type{});
    }
  };
  for (;;) {
    if (!(A.m & kOk)) break;
    step(A, B);
    if (!(B.m & kOk)) break;
    step(B, A);
  }
  if (cur_r != 0xffffffffu) row_done();

  // ---- rows with a pair on a D_MAX / D_0 boundary: exact patch (cold)
  while (fixmask) {
    const unsigned m = (unsigned)__ffs((int)fixmask) - 1u;
    fixmask &= fixmask - 1u;
    const unsigned rf = wid + kSweepWarps * m;
    const unsigned kf = first + rf;
    const RowFixImg f = row_fixup_img<K>(a.pbc_g, a.sw_g, a.spos, a.pos, a.nbr, s_meta[rf], far_on, s_shift, kf, lane,
                                         a.two_groups, kf >= a.n_a, ACC, (ACC && a.scatter_b) ? a.sderiv : nullptr);
    const RecBuf own = s_own[rf];
    const double gx = f.fx * winv, gy = f.fy * winv, gz = f.fz * winv;  // per lane: the sums are linear
    double* sc = &s_c[0][threadIdx.x];
    sc[0 * kSweepThreads] += gx * own.x + f.c[0]; sc[1 * kSweepThreads] += gx * own.y + f.c[1]; sc[2 * kSweepThreads] += gx * own.z + f.c[2];
    sc[3 * kSweepThreads] += gy * own.x + f.c[3]; sc[4 * kSweepThreads] += gy * own.y + f.c[4]; sc[5 * kSweepThreads] += gy * own.z + f.c[5];
    sc[6 * kSweepThreads] += gz * own.x + f.c[6]; sc[7 * kSweepThreads] += gz * own.y + f.c[7]; sc[8 * kSweepThreads] += gz * own.z + f.c[8];
    sc[9 * kSweepThreads] += f.val;
    const double hx = warp_sum(f.fx), hy = warp_sum(f.fy), hz = warp_sum(f.fz);
    if (lane == 0) {  // rows are owned by one warp: nobody else touches these three values
      a.sderiv[3 * (size_t)kf] += hx;
      a.sderiv[3 * (size_t)kf + 1] += hy;
      a.sderiv[3 * (size_t)kf + 2] += hz;
    }
  }
  // ---- block epilogue: one partial record {value, c[9]}
  double val_lane = 0.0;
  if (ACC) {
    val_lane = Unstretched<K>::value ? fma(sw.stretch, val, sw.shift * (double)nin) : val;
    val_lane += s_c[9][threadIdx.x];
  }
  __shared__ double s_red[kSweepWarps][10];
  __shared__ unsigned s_exec[kSweepWarps];
  {
    const double v0 = warp_sum(val_lane);
    if (lane == 0) s_red[wid][0] = v0;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      // component q of (deriv / w) (x) u sits in lane 0,1,2 (row x), 16,17,18 (row y), 8,9,10 (row z)
      const unsigned owner = (q < 3 ? 0u : (q < 6 ? 16u : 8u)) + (unsigned)(q % 3);
      const double t = warp_sum(s_c[q][threadIdx.x] + ((lane == owner) ? cpos : 0.0));
      if (lane == 0) s_red[wid][1 + q] = t;
    }
    if (lane == 0) s_exec[wid] = executed;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kSweepWarps; ++w) t += s_red[w][threadIdx.x];
    a.partials[(size_t)blockIdx.x * kPartialStride + threadIdx.x] = t;
  }
  if (threadIdx.x == 32) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < kSweepWarps; ++w) t += s_exec[w];
    if (t) atomicAdd(a.executed, t);
  }
}

// position part of the virial for rows that were not swept (scatter_b): one partial record per block with
// c = sum deriv_k (x) u_k over the block's share of rows [begin, end); gated like k_sweep_img
__global__ void __launch_bounds__(256)
    k_posvir_rows(SweepArgs a, unsigned begin, unsigned end, unsigned record0) {
  if (!(__longlong_as_double((long long)*a.disp2_bits) < a.img_disp2_max)) return;
  double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (unsigned k = begin + blockIdx.x * blockDim.x + threadIdx.x; k < end; k += gridDim.x * blockDim.x) {
    const SPos p = load_spos(a.spos + k);
    const double dx = a.sderiv[3 * (size_t)k], dy = a.sderiv[3 * (size_t)k + 1], dz = a.sderiv[3 * (size_t)k + 2];
    c[0] = fma(dx, p.x, c[0]); c[1] = fma(dx, p.y, c[1]); c[2] = fma(dx, p.z, c[2]);
    c[3] = fma(dy, p.x, c[3]); c[4] = fma(dy, p.y, c[4]); c[5] = fma(dy, p.z, c[5]);
    c[6] = fma(dz, p.x, c[6]); c[7] = fma(dz, p.y, c[7]); c[8] = fma(dz, p.z, c[8]);
  }
  __shared__ double sm[8][9];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    const double t = warp_sum(c[q]);
    if (lane == 0) sm[wid][q] = t;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double t = 0.0;
    if (threadIdx.x > 0)
      for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x - 1];
    a.partials[(size_t)(record0 + blockIdx.x) * kPartialStride + threadIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// dispatch.  Same grid shape as run_sweep (sweep_kernels.cuh): whichever of the two kernels takes the step fills
// the same partial records.
// rows per block for the image sweep: pick_rows_per_block, reduced until the trips of a warp's rows fit its table
// (a row of r entries in two parts is at most ceil(r / 64) + 1 trips).  0: rows too long for this kernel.
static inline unsigned img_rows_per_block(unsigned rows, unsigned max_row) {
  if (max_row > kImgMaxRow) return 0u;
  const unsigned trips_per_row = (max_row + 63u) / 64u + 1u;
  unsigned per_warp = (unsigned)kImgTripCap / trips_per_row;
  if (per_warp == 0u) return 0u;
  if (per_warp > 16u) per_warp = 16u;
  unsigned rpb = pick_rows_per_block(rows);
  if (rpb > per_warp * kSweepWarps) rpb = per_warp * kSweepWarps;
  return rpb;
}

template <int K, int MINB>
static int run_sweep_img(const SweepArgs& a, const DevSwitch& sw, const ImgShifts& sh, const NearBands& nb, cudaStream_t st) {
  const unsigned acc_end = a.two_groups ? min(a.row_end, a.n_a) : a.row_end;
  int nblocks = 0;
  if (a.row_begin < acc_end) {
    const unsigned rows = acc_end - a.row_begin;
    const unsigned rpb = a.rows_per_block;
    nblocks = (int)((rows + rpb - 1) / rpb);
    k_sweep_img<K, true, MINB><<<nblocks, kSweepThreads, 0, st>>>(a, sw, sh, nb, rpb, a.row_begin, acc_end);
  }
  const unsigned b_begin = max(a.row_begin, acc_end);
  if (b_begin < a.row_end && a.scatter_b) {
    // the GROUPA rows have added +dd to their partners: only the position part of the virial is left to do
    const int nb2 = 148;
    k_posvir_rows<<<nb2, 256, 0, st>>>(a, b_begin, a.row_end, (unsigned)nblocks);
    nblocks += nb2;
  } else if (b_begin < a.row_end) {
    const unsigned rows = a.row_end - b_begin;
    const unsigned rpb = a.rows_per_block;
    const int nb2 = (int)((rows + rpb - 1) / rpb);
    SweepArgs b = a;
    b.partials = a.partials + (size_t)nblocks * kPartialStride;  // B rows: position part of the virial only
    k_sweep_img<K, false, MINB><<<nb2, kSweepThreads, 0, st>>>(b, sw, sh, nb, rpb, b_begin, a.row_end);
    nblocks += nb2;
  }
  return nblocks;
}

}  // namespace b200
