// Neighbour-list rebuild on the device: cell binning, deterministic counting sort, and the
// stream-compacted classic (distance-filtered) pair list.
//
//   reference                                            here
//   LinkCells::findMyCell/findCell  LinkCells.cpp:277-315   k_bin_atoms        (bit-exact cell of every atom)
//   LinkCells::resetCollection      LinkCells.cpp:138-181   k_bin_atoms + k_scan_cells + k_place_atoms + k_order_cells
//   LinkCells::addRequiredCells     LinkCells.cpp:183-239   stencil_bounds / wrap_cell (device_types + here)
//   NeighborList::update (classic)  NeighborList.cpp:237-308 k_nl_rows<false> (count) + scan + k_nl_rows<true> (fill)
//
// All of this is HBM-bound integer/byte work except the distance test in k_nl_rows, which repeats the
// reference's arithmetic exactly (no FMA contraction) so that the pair SET is identical bit for bit.
#include "kernels.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// bounding box of the positions (only when there is no periodic box): min/max per axis
__global__ void k_bbox(const double* __restrict__ pos, unsigned n, double* __restrict__ out /*[6]: min xyz, max xyz*/,
                       unsigned long long* __restrict__ scratch /*[6] ordered-int encodings*/) {
  // encode doubles so that unsigned comparison == floating comparison
  auto enc = [](double v) -> unsigned long long {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
  };
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double v = pos[3 * (size_t)i + k];
      mn[k] = fmin(mn[k], v);
      mx[k] = fmax(mx[k], v);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atomicMin(&scratch[k], enc(mn[k]));
      atomicMax(&scratch[3 + k], enc(mx[k]));
    }
  }
  (void)out;
}

__global__ void k_bbox_decode(const unsigned long long* __restrict__ scratch, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k < 6) {
    unsigned long long e = scratch[k];
    unsigned long long b = (e & 0x8000000000000000ull) ? (e & 0x7fffffffffffffffull) : ~e;
    out[k] = __longlong_as_double((long long)b);
  }
}

// ------------------------------------------------------------------------------------------------
// LinkCells::findCell : f = invBox^T * (pos - origin); c_k = floor((Tools::pbc(f_k)+0.5)*n_k)
__device__ __forceinline__ int cell_of(const DevGrid& g, double px, double py, double pz) {
  double p[3] = {px, py, pz};
  if (g.bbox) {
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = xsub(p[k], g.origin[k]);
  }
  int c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double f = xadd(0.0, xmul(g.inv_box_t[3 * i], p[0]));  // Tensor.h:440-447, accumulation from 0
    f = xadd(f, xmul(g.inv_box_t[3 * i + 1], p[1]));
    f = xadd(f, xmul(g.inv_box_t[3 * i + 2], p[2]));
    const double w = xmul(xadd(tools_pbc_exact(f), 0.5), (double)g.n[i]);
    int ci = (int)floor(w);
    ci = max(0, min(g.n[i] - 1, ci));  // the reference asserts this range (LinkCells.cpp:286-290)
    c[i] = ci;
  }
  return c[0] + c[1] * g.n[0] + c[2] * g.n[0] * g.n[1];
}

// one thread per atom slot: cell id + histogram per (group, cell)
__global__ void k_bin_atoms(const double* __restrict__ pos, unsigned n, unsigned n_a, DevGrid g,
                            uint32_t* __restrict__ cell_of_slot, uint32_t* __restrict__ cell_count /*[ngroups*ncell]*/) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cell_of(g, pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2]);
  cell_of_slot[i] = (uint32_t)c;
  const unsigned grp = (i < n_a) ? 0u : 1u;
  atomicAdd(&cell_count[grp * (unsigned)g.ncell + (unsigned)c], 1u);
}

// single-block exclusive scan over `m` counters (m = ngroups*ncell; group 1 simply continues after group 0,
// which is exactly the [A sorted | B sorted] layout).  Also zeroes the placement cursors.
__global__ void k_scan_cells(const uint32_t* __restrict__ cnt, uint32_t* __restrict__ start, uint32_t* __restrict__ cursor,
                             unsigned m) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (unsigned base = 0; base < m; base += blockDim.x) {
    const unsigned i = base + threadIdx.x;
    const uint32_t v = (i < m) ? cnt[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= (unsigned)o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t t = (lane < nw) ? warp_tot[lane] : 0u;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= (unsigned)o) t += y;
      }
      warp_tot[lane] = t;  // inclusive totals of warps
    }
    __syncthreads();
    const uint32_t before = carry + (wid ? warp_tot[wid - 1] : 0u) + (x - v);
    if (i < m) {
      start[i] = before;
      cursor[i] = 0u;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + v;
    __syncthreads();
  }
}

// scatter slots into their cell segment (order inside a segment is arbitrary here, fixed by k_order_cells)
__global__ void k_place_atoms(unsigned n, unsigned n_a, int ncell, const uint32_t* __restrict__ cell_of_slot,
                              const uint32_t* __restrict__ start, uint32_t* __restrict__ cursor,
                              uint32_t* __restrict__ tmp) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned grp = (i < n_a) ? 0u : 1u;
  const unsigned cc = grp * (unsigned)ncell + cell_of_slot[i];
  const uint32_t k = start[cc] + atomicAdd(&cursor[cc], 1u);
  tmp[k] = i;
}

// one warp per (group,cell) segment: order its slots ascending (what the reference's serial counting sort
// produces, LinkCells.cpp:175-180) by rank counting.  O(m^2/32) per segment, m = atoms in the cell -- the
// pair work on the same cell is O(27 m^2), so this never dominates.
__global__ void k_order_cells(unsigned nseg, int ncell, const uint32_t* __restrict__ start, const uint32_t* __restrict__ cnt,
                              const uint32_t* __restrict__ tmp, uint32_t* __restrict__ perm, uint32_t* __restrict__ scell) {
  const unsigned seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (seg >= nseg) return;
  const uint32_t s = start[seg], m = cnt[seg];
  const uint32_t cell = seg % (unsigned)ncell;
  for (uint32_t e = lane; e < m; e += 32) {
    const uint32_t mine = tmp[s + e];
    uint32_t rank = 0;
    for (uint32_t q = 0; q < m; ++q) rank += (tmp[s + q] < mine) ? 1u : 0u;
    perm[s + rank] = mine;
    scell[s + rank] = cell;
  }
}

// per step: sorted, 32-byte records from the caller's AoS positions.
// TRACK 1 (rebuild step): also remember the positions the list was built from; TRACK 2 (other steps): largest squared
// minimum-image displacement of any atom since then -> disp2 (bits of a non-negative double order like integers).
// The sweep uses it for the Verlet-skin argument: while 2 * max displacement < skin, no partner of the far parts of
// the rows (r > D_MAX + skin at build time) can have come inside D_MAX.
template <int TRACK, int PBC>
__global__ void __launch_bounds__(256)
    k_gather_sorted(const double* __restrict__ pos, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ abs_index,
                    unsigned n, SPos* __restrict__ spos, double* __restrict__ bpos, DevPbc pbc,
                    unsigned long long* __restrict__ disp2) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (k < n) {
    const uint32_t slot = perm[k];
    SPos r;
    r.x = pos[3 * (size_t)slot];
    r.y = pos[3 * (size_t)slot + 1];
    r.z = pos[3 * (size_t)slot + 2];
    r.abs_index = abs_index[slot];
    r.slot = slot;
    spos[k] = r;
    if (TRACK == 1) {
      bpos[3 * (size_t)k] = r.x;
      bpos[3 * (size_t)k + 1] = r.y;
      bpos[3 * (size_t)k + 2] = r.z;
    } else if (TRACK == 2) {
      double dx = r.x - bpos[3 * (size_t)k], dy = r.y - bpos[3 * (size_t)k + 1], dz = r.z - bpos[3 * (size_t)k + 2];
      min_image_fast<PBC>(pbc, dx, dy, dz);
      d2 = fma(dz, dz, fma(dy, dy, dx * dx));
      if (!(d2 >= 0.0)) d2 = INFINITY;  // NaN positions: never skip anything
    }
  }
  if (TRACK == 2) {
    __shared__ double sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = d2;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = sm[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) m = fmax(m, sm[w]);
      if (m > 0.0) atomicMax(disp2, (unsigned long long)__double_as_longlong(m));
    }
  }
}

__global__ void k_identity_perm(unsigned n, uint32_t* __restrict__ perm, uint32_t* __restrict__ scell) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    perm[k] = k;
    scell[k] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// classic list rows.  One warp per row (sorted atom k in [row_begin,row_end)); candidates = atoms of the
// partner group in the stencil cells; a candidate is kept iff modulo2(Pbc::distance) <= cutoff^2 evaluated
// exactly as NeighborList.cpp:246-259 does.  FILL=false counts, FILL=true writes the sorted indices j in
// stencil order (deterministic).  Self-pairs (same absolute index) are not stored: the sweep would skip
// them anyway (CoordinationBase.cpp:183); they are accounted for in the reported list size on the host.
template <bool FILL, int PBC>
__global__ void __launch_bounds__(256)
    k_nl_rows(const SPos* __restrict__ spos, const uint32_t* __restrict__ scell, const uint32_t* __restrict__ cstart,
              const uint32_t* __restrict__ ccount, DevGrid g, DevPbc pbc, double cutoff2, unsigned n_a, int two_groups,
              unsigned row_begin, unsigned row_end, uint32_t* __restrict__ row_count,
              const unsigned long long* __restrict__ row_start, uint32_t* __restrict__ nbr) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  const unsigned k = row_begin + warp;
  if (k >= row_end) return;
  const SPos pi = load_spos(spos + k);
  const unsigned my_grp = (k < n_a) ? 0u : 1u;
  const unsigned other = two_groups ? (1u - my_grp) : 0u;
  int c[3];
  cell_coords(g, (int)scell[k], c);
  // Pre-filter in fused arithmetic: a candidate whose r^2 is clearly outside / inside the cutoff needs no
  // bit-exact evaluation (the two evaluations differ by < 1e-11 relative); only the thin band around the
  // cutoff runs the reference's exact operation sequence, so the kept SET is still the reference's.
  const double band = 1e-9 * cutoff2;
  const double c2_hi = cutoff2 + band, c2_lo = cutoff2 - band;
  unsigned total = 0;
  const unsigned long long base = FILL ? row_start[k - row_begin] : 0ull;
  for_each_stencil_range(g, c, other * (unsigned)g.ncell, cstart, ccount, [&](uint32_t s, uint32_t m, int, int, int) {
    for (uint32_t e0 = 0; e0 < m; e0 += 32) {
      const uint32_t e = e0 + lane;
      bool keep = false;
      uint32_t j = 0;
      if (e < m) {
        j = s + e;
        const SPos pj = load_spos(spos + j);
        if (j != k && pj.abs_index != pi.abs_index) {
          double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
          min_image_fast<PBC>(pbc, dx, dy, dz);
          const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
          if (r2 < c2_lo) {
            keep = true;
          } else if (r2 <= c2_hi) {
            // the reference evaluates the pair as (index0,index1) = (A atom, B atom) resp. (lower, higher
            // slot): distance = pos[index1]-pos[index0]   (NeighborList.cpp:247-254)
            const bool i_first = two_groups ? (my_grp == 0u) : (pi.slot < pj.slot);
            double d[3];
            if (i_first) {
              d[0] = xsub(pj.x, pi.x);
              d[1] = xsub(pj.y, pi.y);
              d[2] = xsub(pj.z, pi.z);
            } else {
              d[0] = xsub(pi.x, pj.x);
              d[1] = xsub(pi.y, pj.y);
              d[2] = xsub(pi.z, pj.z);
            }
            min_image_exact(pbc, d);
            keep = norm2_exact(d[0], d[1], d[2]) <= cutoff2;
          }
        }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (FILL && keep) nbr[base + total + __popc(mask & ((1u << lane) - 1u))] = j;
      total += __popc(mask);
    }
  });
  if (FILL) {  // rows are padded to 4 entries (k_scan_*): make the padding a harmless index
    const unsigned pad = ((total + 3u) & ~3u) - total;
    if (lane < pad) nbr[base + total + lane] = 0u;
  }
  if (!FILL && lane == 0) row_count[k - row_begin] = total;
}


// The reference's neighbour test on the caller's unmodified positions (NeighborList.cpp:246-259): the pair is
// evaluated as (index0,index1) = (A atom, B atom) resp. (lower, higher slot), distance = pos[index1]-pos[index0],
// kept iff modulo2(distance) <= cutoff^2, every operation as the reference's non-FMA build does it.
// k, j: sorted indices; perm: sorted -> slot.
// pbc_g lives in GLOBAL memory: a reference to the 1.4 KB kernel parameter would make every thread of the caller
// copy it to its stack.
__device__ __noinline__ bool exact_within(const double* __restrict__ pos, const uint32_t* __restrict__ perm,
                                          const DevPbc* __restrict__ pbc_g, uint32_t k, uint32_t j, int two_groups,
                                          bool k_in_a, double cutoff2) {
  const uint32_t sk = perm[k], sj = perm[j];
  const double* pk = pos + 3 * (size_t)sk;
  const double* pj = pos + 3 * (size_t)sj;
  const bool i_first = two_groups ? k_in_a : (sk < sj);
  double d[3];
  if (i_first) {
    d[0] = xsub(pj[0], pk[0]);
    d[1] = xsub(pj[1], pk[1]);
    d[2] = xsub(pj[2], pk[2]);
  } else {
    d[0] = xsub(pk[0], pj[0]);
    d[1] = xsub(pk[1], pj[1]);
    d[2] = xsub(pk[2], pj[2]);
  }
  min_image_exact(*pbc_g, d);
  return norm2_exact(d[0], d[1], d[2]) <= cutoff2;
}

// ------------------------------------------------------------------------------------------------
// FP32 candidate search.  The sorted atoms are copied once per rebuild into float4 records holding the
// position WRAPPED into the cell the atom was binned to (relative to the box centre; bounding-box mode:
// relative to the origin), so that a candidate in a stencil cell reached through the periodic boundary is
// seen through the image (wx,wy,wz) of that cell: d = (l_j + wx*a + wy*b + wz*c) - l_i, no per-candidate
// minimum-image arithmetic and all of it on the FP32 pipe.  For cells at least one cutoff wide (and >=
// 2*radius+1 of them per periodic direction) that image is the minimum image of every pair within the
// cutoff.  FP32 rounding moves r^2 by < band_rel*cutoff^2; candidates inside that band -- a ~1e-5 fraction
// -- are decided by the reference's exact FP64 operation sequence on the original positions, so the kept
// SET is the reference's bit for bit.
// After a re-sort (image mode, see sweep_img.cuh): everything the following steps continue from.
//   braw  = the caller's position of sorted atom k now,
//   wpos  = that position wrapped into the cell the atom was binned to (same arithmetic as cell_of()),
//   lpos  = float4 copy of wpos for the FP32 search (w = absolute index bits),
//   spos  = the record the sweep reads: u = wpos under PBC (the periodic images stored with the list entries refer
//           to these coordinates), the caller's position otherwise; ubuild = u (displacement bound).
__global__ void k_sort_init(const double* __restrict__ pos, const uint32_t* __restrict__ perm,
                            const uint32_t* __restrict__ abs_index, unsigned n, DevGrid g, DevPbc box, int use_wrapped,
                            float4* __restrict__ lpos, double* __restrict__ wpos, double* __restrict__ braw,
                            SPos* __restrict__ spos, double* __restrict__ ubuild) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t slot = perm[k];
  const double q[3] = {pos[3 * (size_t)slot], pos[3 * (size_t)slot + 1], pos[3 * (size_t)slot + 2]};
  double out[3];
  if (g.bbox) {
#pragma unroll
    for (int a = 0; a < 3; ++a) out[a] = xsub(q[a], g.origin[a]);
  } else {
    double fw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // same arithmetic as cell_of(): the wrapped scaled coordinate the cell came from
      double f = xadd(0.0, xmul(g.inv_box_t[3 * i], q[0]));
      f = xadd(f, xmul(g.inv_box_t[3 * i + 1], q[1]));
      f = xadd(f, xmul(g.inv_box_t[3 * i + 2], q[2]));
      fw[i] = tools_pbc_exact(f);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) out[a] = fw[0] * box.box[a] + fw[1] * box.box[3 + a] + fw[2] * box.box[6 + a];
  }
  const uint32_t ab = abs_index[slot];
  lpos[k] = make_float4((float)out[0], (float)out[1], (float)out[2], __uint_as_float(ab));
  SPos r;
  r.x = use_wrapped ? out[0] : q[0];
  r.y = use_wrapped ? out[1] : q[1];
  r.z = use_wrapped ? out[2] : q[2];
  r.abs_index = ab;
  r.slot = slot;
  spos[k] = r;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    wpos[3 * (size_t)k + a] = out[a];
    braw[3 * (size_t)k + a] = q[a];
  }
  ubuild[3 * (size_t)k] = r.x;
  ubuild[3 * (size_t)k + 1] = r.y;
  ubuild[3 * (size_t)k + 2] = r.z;
}

// Per step in image mode: u = wpos + minimum-image displacement since the sort -- continuous coordinates: the MD
// engine may re-wrap atoms at will, the periodic image stored with every list entry stays the right one.
// BUILD (a rebuild that keeps the sort and filters the super-list): also the float4 copy for the FP32 test, the
// largest squared displacement since the sort (validity of the super-list), and ubuild = u.
// !BUILD: the largest squared displacement since the list was built (far-part skip, validity of the images).
template <int PBC, bool BUILD>
__global__ void __launch_bounds__(256)
    k_gather_u(PosSrc pos, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ abs_index, IdxRanges rng,
               const double* __restrict__ wpos, const double* __restrict__ braw, DevPbc pbc,
               SPos* __restrict__ spos, double* __restrict__ ubuild, float4* __restrict__ lpos,
               unsigned long long* __restrict__ disp2) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (t < rng.total) {
    unsigned k = 0, rem = t;
    bool found = false;
#pragma unroll
    for (int q = 0; q < 6; ++q) {  // thread t -> the t-th sorted index of the intervals
      if (!found && q < (int)rng.n) {
        if (rem < rng.len[q]) {
          k = rng.lo[q] + rem;
          found = true;
        } else {
          rem -= rng.len[q];
        }
      }
    }
    const uint32_t slot = perm[k];
    const double* __restrict__ q3 = pos_at(pos, slot);
    const double qx = q3[0], qy = q3[1], qz = q3[2];
    double dx = qx - braw[3 * (size_t)k], dy = qy - braw[3 * (size_t)k + 1], dz = qz - braw[3 * (size_t)k + 2];
    min_image_fast<PBC>(pbc, dx, dy, dz);
    const double wx = wpos[3 * (size_t)k] + dx, wy = wpos[3 * (size_t)k + 1] + dy, wz = wpos[3 * (size_t)k + 2] + dz;
    SPos r;
    r.x = PBC ? wx : qx;
    r.y = PBC ? wy : qy;
    r.z = PBC ? wz : qz;
    const uint32_t ab = abs_index[slot];
    r.abs_index = ab;
    r.slot = slot;
    spos[k] = r;
    if (BUILD) {
      lpos[k] = make_float4((float)wx, (float)wy, (float)wz, __uint_as_float(ab));
      ubuild[3 * (size_t)k] = r.x;
      ubuild[3 * (size_t)k + 1] = r.y;
      ubuild[3 * (size_t)k + 2] = r.z;
      d2 = fma(dz, dz, fma(dy, dy, dx * dx));
    } else {
      const double ex = r.x - ubuild[3 * (size_t)k], ey = r.y - ubuild[3 * (size_t)k + 1], ez = r.z - ubuild[3 * (size_t)k + 2];
      d2 = fma(ez, ez, fma(ey, ey, ex * ex));
    }
    if (!(d2 >= 0.0)) d2 = INFINITY;  // NaN positions: never skip anything
  }
  __shared__ double sm[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(0xffffffffu, d2, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = d2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = sm[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmax(m, sm[w]);
    if (m > 0.0) atomicMax(disp2, (unsigned long long)__double_as_longlong(m));
  }
}

// one 16-byte record per row for the image sweep: {row start / 4, near count, far offset, far count}; *listed += all
// entries of these rows (what the reference's list holds, counted from both ends)
__global__ void __launch_bounds__(256)
    k_pack_meta(unsigned rows, const unsigned long long* __restrict__ row_start, const uint32_t* __restrict__ row_count,
                const uint32_t* __restrict__ far_off, const uint32_t* __restrict__ far_cnt, uint4* __restrict__ meta,
                unsigned long long* __restrict__ listed) {
  const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned mine = 0u;
  if (r < rows) {
    const uint4 m = make_uint4((uint32_t)(row_start[r] >> 2), row_count[r], far_off[r], far_cnt[r]);
    meta[r] = m;
    mine = m.y + m.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  __shared__ unsigned sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w];
    if (t) atomicAdd(listed, t);
  }
}

// FP32 constants of the candidate search
struct SearchF32 {
  float box[9];        // box rows
  float c2_hi, c2_lo;  // cutoff^2 * (1 +- band): outside -> decided in FP32, inside -> exact FP64 test
};

// SUPER: the rows of the super-list (cutoff + delta, kernels.cuh): entries carry the periodic image they were found
// through in their top 6 bits, no near/far split.
// float4 record `idx`: address = one IMAD.WIDE (index * 16 + base), one 128-bit load
__device__ __forceinline__ float4 load_f4_idx(const float4* __restrict__ base, uint32_t idx) {
  float4 v;
  asm volatile("{\n\t.reg .u64 p;\n\tmad.wide.u32 p, %4, 16, %5;\n\tld.global.nc.v4.f32 {%0,%1,%2,%3}, [p];\n\t}"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(idx), "l"(base));
  return v;
}
// 32-bit word `idx` of a block-uniform array: one IMAD.WIDE + load, no per-lane 64-bit pointer to keep alive
__device__ __forceinline__ uint32_t load_u32_idx(const uint32_t* __restrict__ base, uint32_t idx) {
  uint32_t v;
  asm volatile("{\n\t.reg .u64 p;\n\tmad.wide.u32 p, %1, 4, %2;\n\tld.global.nc.u32 %0, [p];\n\t}" : "=r"(v) : "r"(idx), "l"(base));
  return v;
}
__device__ __forceinline__ void prefetch_l2_idx(const uint32_t* __restrict__ base, uint32_t idx) {
  asm volatile("{\n\t.reg .u64 p;\n\tmad.wide.u32 p, %0, 4, %1;\n\tprefetch.global.L2 [p];\n\t}" ::"r"(idx), "l"(base));
}

template <bool FILL, bool CAPPED, bool SUPER, bool IMAGES, int MINB>
__global__ void __launch_bounds__(256, MINB)
    k_nl_rows_f32(const double* __restrict__ pos, const uint32_t* __restrict__ perm, const float4* __restrict__ lpos,
                  const uint32_t* __restrict__ scell,
                  const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ ccount, DevGrid g,
                  const DevPbc* __restrict__ pbc_g, SearchF32 f, double cutoff2, unsigned n_a, int two_groups, unsigned row_begin,
                  unsigned row_end, uint32_t* __restrict__ row_count, const unsigned long long* __restrict__ row_start,
                  uint32_t* __restrict__ nbr, unsigned row_cap, unsigned* __restrict__ cap_info /*[0] max count, [1] overflow*/,
                  float far2, uint32_t* __restrict__ row_far_off, uint32_t* __restrict__ row_far_cnt) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  const unsigned k = row_begin + warp;
  if (k >= row_end) return;
  const float4 li = lpos[k];
  const unsigned my_abs = __float_as_uint(li.w);
  const unsigned my_grp = (k < n_a) ? 0u : 1u;
  const unsigned group_off = (two_groups ? (1u - my_grp) : 0u) * (unsigned)g.ncell;
  int c[3], lo[3], hi[3];
  cell_coords(g, (int)scell[k], c);
  stencil_bounds(g, c, lo, hi);
  const float c2_hi = f.c2_hi, c2_lo = f.c2_lo;  // FP32 constants come ready from the host: converting them here costs
                                                  // XU-pipe instructions in every stencil column

  // ---- range table, one (y,z) stencil column per lane (<= 25): its x-run is one contiguous sorted range, or
  // two when it wraps around the box.  All the dependent cstart/ccount loads of a row are in flight at once.
  const int ny_n = hi[1] - lo[1], nz_n = hi[2] - lo[2];
  const int ncol = ny_n * nz_n;
  uint32_t sA = 0, mA = 0, sB = 0, mB = 0;
  int wxA = 0, wxB = 0, wy = 0, wz = 0;
  if ((int)lane < ncol) {
    const int ny = lo[1] + (int)lane / nz_n, nz = lo[2] + (int)lane % nz_n;
    wy = wrap_count(ny, g.n[1]);
    wz = wrap_count(nz, g.n[2]);
    const unsigned cbase = group_off + (unsigned)(wrap_cell(ny, g.n[1]) * g.n[0] + wrap_cell(nz, g.n[2]) * g.n[0] * g.n[1]);
    int x = lo[0];
    {
      const int xw = wrap_cell(x, g.n[0]);
      const int run = min(hi[0] - x, g.n[0] - xw);
      const unsigned first = cbase + (unsigned)xw, last = first + (unsigned)run - 1u;
      sA = cstart[first];
      mA = cstart[last] + ccount[last] - sA;
      wxA = wrap_count(x, g.n[0]);
      x += run;
    }
    if (x < hi[0]) {  // the part of the run on the other side of the periodic boundary
      const int xw = wrap_cell(x, g.n[0]);
      const int run = min(hi[0] - x, g.n[0] - xw);
      const unsigned first = cbase + (unsigned)xw, last = first + (unsigned)run - 1u;
      sB = cstart[first];
      mB = cstart[last] + ccount[last] - sB;
      wxB = wrap_count(x, g.n[0]);
    }
  }
  const float ax = f.box[0], ay = f.box[1], az = f.box[2];
  const float bx = f.box[3], by = f.box[4], bz = f.box[5];
  const float cx = f.box[6], cy = f.box[7], cz = f.box[8];

  unsigned total = 0, total_far = 0;
  // CAPPED: single-pass build into fixed-capacity rows (capacity learnt from the previous rebuild)
  const unsigned long long base = CAPPED ? (unsigned long long)(k - row_begin) * row_cap : (FILL ? row_start[k - row_begin] : 0ull);
  // 32-bit list: the row's allocation is filled from both ends, near partners (r^2 <= far2 now) forwards from its
  // start, far ones backwards from its end (two-pass build: row_count still holds the total of the count pass)
  const unsigned alloc = CAPPED ? row_cap : (FILL ? ((row_count[k - row_begin] + 3u) & ~3u) : 0u);

  // One batch = 32 candidates, one per lane, without divergent branches (the kernel is bound by instruction issue):
  // predicates, two ballots, one predicated store.  A candidate inside the FP32 rounding band of the cutoff takes the
  // warp through the exact FP64 decision (NeighborList.cpp:246-259).
  unsigned lt;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
  uint32_t* __restrict__ rowp = FILL ? nbr + base : nbr;
  auto batch = [&](bool keep, float r2, uint32_t j, uint32_t img) {
    const bool band = keep & (r2 > c2_lo);
    if (__any_sync(0xffffffffu, band)) {
      if (band) keep = exact_within(pos, perm, pbc_g, k, j, two_groups, my_grp == 0u, cutoff2);
    }
    const bool far = !SUPER && (r2 > far2);  // only orders the row: any classification gives the same results
    const unsigned mk = __ballot_sync(0xffffffffu, keep);
    const unsigned mf = SUPER ? 0u : __ballot_sync(0xffffffffu, keep & far);
    const unsigned mn = mk & ~mf;
    if (FILL) {
      const unsigned at = (far ? total_far : total) + __popc((far ? mf : mn) & lt);
      const unsigned idx = far ? alloc - 1u - at : at;
      if (keep & (at < alloc)) rowp[idx] = j | img;
    }
    total += __popc(mn);
    if (!SUPER) total_far += __popc(mf);
  };
  // one contiguous range: two 32-candidate batches per trip so that two loads are in flight per lane
  auto scan_range = [&](uint32_t s, uint32_t m, int wx, int wyy, int wzz) {
    const float ox = li.x - ((float)wx * ax + (float)wyy * bx + (float)wzz * cx);
    const float oy = li.y - ((float)wx * ay + (float)wyy * by + (float)wzz * cy);
    const float oz = li.z - ((float)wx * az + (float)wyy * bz + (float)wzz * cz);
    const uint32_t img = IMAGES ? super_image(wx, wyy, wzz) : 0u;
    for (uint32_t e0 = 0; e0 < m; e0 += 64) {
      const uint32_t e1 = e0 + lane, e2 = e1 + 32;
      const bool in1 = e1 < m, in2 = e2 < m;
      const uint32_t j1 = s + e1, j2 = s + e2;
      const float4 l1 = load_f4_idx(lpos, in1 ? j1 : k);
      const float4 l2 = load_f4_idx(lpos, in2 ? j2 : k);
      const float dx1 = l1.x - ox, dy1 = l1.y - oy, dz1 = l1.z - oz;
      const float dx2 = l2.x - ox, dy2 = l2.y - oy, dz2 = l2.z - oz;
      const float r1 = fmaf(dz1, dz1, fmaf(dy1, dy1, dx1 * dx1));
      const float r2 = fmaf(dz2, dz2, fmaf(dy2, dy2, dx2 * dx2));
      batch(in1 & (j1 != k) & (__float_as_uint(l1.w) != my_abs) & (r1 < c2_hi), r1, j1, img);
      if (e0 + 32 < m) batch(in2 & (j2 != k) & (__float_as_uint(l2.w) != my_abs) & (r2 < c2_hi), r2, j2, img);
    }
  };
  for (int col = 0; col < ncol; ++col) {
    const uint32_t s1 = __shfl_sync(0xffffffffu, sA, col), m1 = __shfl_sync(0xffffffffu, mA, col);
    const uint32_t s2 = __shfl_sync(0xffffffffu, sB, col), m2 = __shfl_sync(0xffffffffu, mB, col);
    const int w1 = __shfl_sync(0xffffffffu, wxA, col), w2 = __shfl_sync(0xffffffffu, wxB, col);
    const int wyy = __shfl_sync(0xffffffffu, wy, col), wzz = __shfl_sync(0xffffffffu, wz, col);
    if (m1) scan_range(s1, m1, w1, wyy, wzz);
    if (m2) scan_range(s2, m2, w2, wyy, wzz);
  }
  const unsigned all = total + total_far;
  if (lane == 0) {
    if (FILL) {
      row_count[k - row_begin] = min(total, alloc);
      if (!SUPER) {
        row_far_cnt[k - row_begin] = min(total_far, alloc);
        row_far_off[k - row_begin] = alloc - min(total_far, alloc);
      }
    } else {
      row_count[k - row_begin] = all;
    }
    if (CAPPED || !FILL) {
      if (all > cap_info[0]) atomicMax(&cap_info[0], all);
    }
    if (CAPPED && all > row_cap) atomicExch(&cap_info[1], 1u);
  }
}

// ------------------------------------------------------------------------------------------------
// Rebuild by FILTERING the super-list: the candidates of row k are the entries of its super-list row (every atom that
// was within NL_CUTOFF + delta when the super-list was built; valid while 2 * max displacement < delta, checked by
// the host), seen through the periodic image stored with each entry.  Same FP32 test, same exact FP64 decision
// inside the rounding band and the same two-ended near/far rows as k_nl_rows_f32 -- ~1/3 of its candidates, no
// cell tables, no re-sort.
template <bool FILL, bool CAPPED, bool IMAGES>
__global__ void __launch_bounds__(256, 3)
    k_nl_filter(const double* __restrict__ pos, const uint32_t* __restrict__ perm, const float4* __restrict__ lpos,
                const unsigned long long* __restrict__ srow_start,
                const uint32_t* __restrict__ srow_count, const uint32_t* __restrict__ snbr, const DevPbc* __restrict__ pbc_g,
                SearchF32 f,
                double cutoff2, unsigned n_a, int two_groups, unsigned row_begin, unsigned row_end,
                uint32_t* __restrict__ row_count, const unsigned long long* __restrict__ row_start, uint32_t* __restrict__ nbr,
                unsigned row_cap, unsigned* __restrict__ cap_info, float far2, uint32_t* __restrict__ row_far_off,
                uint32_t* __restrict__ row_far_cnt) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  const unsigned k = row_begin + warp;
  if (k >= row_end) return;
  const float4 li = lpos[k];
  const unsigned my_abs = __float_as_uint(li.w);
  const unsigned my_grp = (k < n_a) ? 0u : 1u;
  const float c2_hi = f.c2_hi, c2_lo = f.c2_lo;
  const unsigned long long sbase = srow_start[k - row_begin];
  const unsigned m = srow_count[k - row_begin];
  unsigned total = 0, total_far = 0;
  const unsigned long long base = CAPPED ? (unsigned long long)(k - row_begin) * row_cap : (FILL ? row_start[k - row_begin] : 0ull);
  const unsigned alloc = CAPPED ? row_cap : (FILL ? ((row_count[k - row_begin] + 3u) & ~3u) : 0u);
  auto exact_keep = [&](uint32_t j) -> bool { return exact_within(pos, perm, pbc_g, k, j, two_groups, my_grp == 0u, cutoff2); };
  // `shifted`: some entry of this trip is seen through a periodic image (warp-uniform; rare away from the box faces)
  auto test = [&](bool in, uint32_t entry, const float4 lj, bool shifted, bool& far) -> bool {
    far = false;
    if (!in) return false;
    float dx = lj.x - li.x, dy = lj.y - li.y, dz = lj.z - li.z;
    if (shifted) {
      const uint32_t code = (entry >> 26) ^ kImageCentre;
      const float wx = (float)((int)(code & 3u) - 1), wy = (float)((int)((code >> 2) & 3u) - 1),
                  wz = (float)((int)((code >> 4) & 3u) - 1);
      dx = lj.x - (li.x - (wx * f.box[0] + wy * f.box[3] + wz * f.box[6]));
      dy = lj.y - (li.y - (wx * f.box[1] + wy * f.box[4] + wz * f.box[7]));
      dz = lj.z - (li.z - (wx * f.box[2] + wy * f.box[5] + wz * f.box[8]));
    }
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    bool keep = (__float_as_uint(lj.w) != my_abs) && (r2 < c2_hi);  // j != k already holds in the super-list
    if (keep && r2 > c2_lo) keep = exact_keep(entry & kSuperIndexMask);
    far = r2 > far2;
    return keep;
  };
  auto emit = [&](bool keep, bool far, uint32_t j) {
    const unsigned below = (1u << lane) - 1u;
    const unsigned mn = __ballot_sync(0xffffffffu, keep && !far), mf = __ballot_sync(0xffffffffu, keep && far);
    if (FILL && keep) {
      const unsigned at = far ? total_far + __popc(mf & below) : total + __popc(mn & below);
      if (at < alloc) nbr[base + (far ? alloc - 1u - at : at)] = j;
    }
    total += __popc(mn);
    total_far += __popc(mf);
  };
  // Two 32-candidate batches per trip, software-pipelined: the super-list entries (streamed from HBM) run two trips
  // ahead, the float4 records they point to (L1/L2 gather) one trip ahead of the test.
  const uint32_t* __restrict__ srow = snbr + sbase + lane;
  const uint32_t centre = 0u;  // xor-encoded images: the home cell is code 0
  auto entry_at = [&](uint32_t e) -> uint32_t { return (e + lane < m) ? __ldg(srow + e) : centre; };
  uint32_t c1 = entry_at(0), c2 = entry_at(32);
  float4 l1 = __ldg(lpos + (c1 & kSuperIndexMask)), l2 = __ldg(lpos + (c2 & kSuperIndexMask));
  uint32_t n1 = entry_at(64), n2 = entry_at(96);
  for (uint32_t e0 = 0; e0 < m; e0 += 64) {
    // the entries stream from HBM and a warp only lives for ~9 trips: ask L2 for the piece four trips down the row
    // (8 bytes per lane cover its 256 bytes; past the end of the row the request is harmless)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(srow + e0 + 256u + lane));
    const uint32_t a1 = c1, a2 = c2;
    const float4 p1 = l1, p2 = l2;
    c1 = n1;
    c2 = n2;
    l1 = __ldg(lpos + (c1 & kSuperIndexMask));  // entry 0 (a valid atom) when past the end of the row
    l2 = __ldg(lpos + (c2 & kSuperIndexMask));
    n1 = entry_at(e0 + 128);
    n2 = entry_at(e0 + 160);
    const bool in1 = e0 + lane < m, in2 = e0 + 32 + lane < m;
    const bool shifted = __any_sync(0xffffffffu, ((a1 & ~kSuperIndexMask) != centre) || ((a2 & ~kSuperIndexMask) != centre));
    bool f1, f2;
    const bool k1 = test(in1, a1, p1, shifted, f1);
    const bool k2 = test(in2, a2, p2, shifted, f2);
    emit(k1, f1, IMAGES ? a1 : (a1 & kSuperIndexMask));
    if (e0 + 32 < m) emit(k2, f2, IMAGES ? a2 : (a2 & kSuperIndexMask));
  }
  const unsigned all = total + total_far;
  if (lane == 0) {
    if (FILL) {
      row_count[k - row_begin] = min(total, alloc);
      row_far_cnt[k - row_begin] = min(total_far, alloc);
      row_far_off[k - row_begin] = alloc - min(total_far, alloc);
    } else {
      row_count[k - row_begin] = all;
    }
    if (CAPPED || !FILL) {
      if (all > cap_info[0]) atomicMax(&cap_info[0], all);
    }
    if (CAPPED && all > row_cap) atomicExch(&cap_info[1], 1u);
  }
}

// ------------------------------------------------------------------------------------------------
// The same filter with ONE pipeline per warp that runs across rows.  A super-list row is ~9 trips long, and the
// kernel above pays three dependent memory round trips (row header -> entries -> records) at the start of every row.
// Here a block stages the headers of its <= 128 rows once (coalesced), every warp lists the trips of its <= 16 rows
// in a shared-memory table (as k_sweep_img does), and the loop runs over that table: entries two trips ahead,
// records one trip ahead, across row boundaries.  A row switch costs the write of the finished row's three counters.
constexpr int kFltRowsPerWarp = 16;
constexpr int kFltTripCap = 208;
constexpr int kFltMinBlocks = 2;   // resident blocks per SM the register allocation is sized for  // trips per warp the table holds (launch_nl_filter sizes rows_per_block for it)

template <bool FILL, bool CAPPED, bool IMAGES, int MINB>
__global__ void __launch_bounds__(256, MINB)
    k_nl_filter_flat(const double* __restrict__ pos, const uint32_t* __restrict__ perm, const float4* __restrict__ lpos,
                     const unsigned long long* __restrict__ srow_start, const uint32_t* __restrict__ srow_count,
                     const uint32_t* __restrict__ snbr, const DevPbc* __restrict__ pbc_g, SearchF32 f, double cutoff2,
                     unsigned n_a, int two_groups, unsigned row_begin, unsigned row_end, uint32_t* __restrict__ row_count,
                     const unsigned long long* __restrict__ row_start, uint32_t* __restrict__ nbr, unsigned row_cap,
                     unsigned* __restrict__ cap_info, float far2, uint32_t* __restrict__ row_far_off,
                     uint32_t* __restrict__ row_far_cnt, unsigned rows_per_block) {
  // trip descriptor: .x = first entry of the trip relative to the block's first super-list row,
  //                  .y = entries left in the row | row slot of the warp << 16 | kNew (first trip of a row) | kOk
  constexpr unsigned kOk = 0x80000000u, kNew = 0x40000000u, kRem = 0xffffu;
  constexpr int kRows = 8 * kFltRowsPerWarp;
  __shared__ float4 s_li[kRows];
  __shared__ uint32_t s_m[kRows], s_off[kRows], s_alloc[kRows];
  __shared__ unsigned long long s_base[kRows];
  __shared__ uint2 s_trip[8][kFltTripCap + 6];
  const unsigned first = row_begin + blockIdx.x * rows_per_block;
  if (first >= row_end) return;
  const unsigned nrows = min(rows_per_block, row_end - first);
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned long long sb0 = srow_start[first - row_begin];
  for (unsigned t = threadIdx.x; t < nrows; t += 256u) {
    const unsigned rr = first + t - row_begin;
    s_li[t] = lpos[first + t];
    const uint32_t m = srow_count[rr];
    s_m[t] = m;
    s_off[t] = (uint32_t)(srow_start[rr] - sb0);
    const uint32_t alloc = CAPPED ? row_cap : (FILL ? ((row_count[rr] + 3u) & ~3u) : 0u);
    s_alloc[t] = alloc;
    s_base[t] = CAPPED ? (unsigned long long)rr * row_cap : (FILL ? row_start[rr] : 0ull);
    if (m == 0u) {  // no trips: nobody will come by to write this row's counters
      row_count[rr] = 0u;
      if (FILL) {
        row_far_cnt[rr] = 0u;
        row_far_off[rr] = alloc;
      }
    }
  }
  __syncthreads();
  {
    const unsigned rl = wid + 8u * lane;  // this lane's row of the warp (lanes >= 16 never have one)
    uint32_t m = 0u, off0 = 0u;
    if (lane < (unsigned)kFltRowsPerWarp && rl < nrows) {
      m = s_m[rl];
      off0 = s_off[rl];
    }
    const unsigned tn = (m + 63u) >> 6;
    uint32_t total;
    unsigned at = warp_exclusive_scan(tn, lane, total);
    for (unsigned t = 0; t < tn; ++t)
      s_trip[wid][at++] = make_uint2(off0 + 64u * t, (m - 64u * t) | (lane << 16) | kOk | (t == 0u ? kNew : 0u));
    if (lane < 6u) s_trip[wid][total + lane] = make_uint2(0u, 0u);  // the pipeline looks up to five trips past the end
    __syncwarp();
  }
  const float c2_hi = f.c2_hi, c2_lo = f.c2_lo;
  unsigned lt;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
  // row state
  unsigned k = 0u, my_abs = 0u, alloc = 0u, total = 0u, total_far = 0u;
  bool grp_a = true, have_row = false;
  float lix = 0.f, liy = 0.f, liz = 0.f;
  uint32_t* __restrict__ rowp = nbr;  // start of the current row's allocation
  auto flush = [&]() {
    if (lane == 0) {
      const unsigned rr = k - row_begin;
      const unsigned all = total + total_far;
      if (FILL) {
        row_count[rr] = min(total, alloc);
        row_far_cnt[rr] = min(total_far, alloc);
        row_far_off[rr] = alloc - min(total_far, alloc);
      } else {
        row_count[rr] = all;
      }
      if (CAPPED || !FILL) {
        if (all > cap_info[0]) atomicMax(&cap_info[0], all);
      }
      if (CAPPED && all > row_cap) atomicExch(&cap_info[1], 1u);
    }
  };
  // One batch = 32 candidates, one per lane.  The kernel is bound by instruction issue (ncu, per-row kernel: 72 % of
  // the issue slots, 206 instructions per 64-candidate trip), so a batch has no divergent branch: predicates, two
  // ballots, one predicated store.  Only a candidate inside the FP32 rounding band of the cutoff (a few per 10^4)
  // takes the warp through the exact FP64 decision.
  auto batch = [&](bool keep, float r2, uint32_t entry) {
    const bool band = keep & (r2 > c2_lo);
    if (__any_sync(0xffffffffu, band)) {
      if (band) keep = exact_within(pos, perm, pbc_g, k, entry & kSuperIndexMask, two_groups, grp_a, cutoff2);
    }
    const bool far = r2 > far2;
    const unsigned mk = __ballot_sync(0xffffffffu, keep);
    const unsigned mf = __ballot_sync(0xffffffffu, keep & far);
    const unsigned mn = mk & ~mf;
    if (FILL) {
      const unsigned at = (far ? total_far : total) + __popc((far ? mf : mn) & lt);
      const unsigned idx = far ? alloc - 1u - at : at;
      if (keep & (at < alloc)) rowp[idx] = IMAGES ? entry : (entry & kSuperIndexMask);
    }
    total += __popc(mn);
    total_far += __popc(mf);
  };
  const uint32_t* __restrict__ blk = snbr + sb0;  // block-uniform; per-lane offsets stay 32-bit
  struct Set {
    uint32_t e1, e2;  // entries (lane, lane + 32) of the set's trip
    uint32_t dy;      // its descriptor (.y)
    float4 l1, l2;    // their records
  };
  unsigned it = 0u;  // next trip of the table
  auto refill = [&](Set& s) {  // descriptor and entries of the table's next trip, L2 prefetch four trips down
    const uint2 d = s_trip[wid][it];
    const uint2 pd = s_trip[wid][it + 4];
    ++it;
    const unsigned rem = d.y & kRem;
    s.dy = d.y;
    s.e1 = (lane < rem) ? load_u32_idx(blk, d.x + lane) : 0u;  // 0 = sorted atom 0, home image: a valid record
    s.e2 = (lane + 32u < rem) ? load_u32_idx(blk, d.x + 32u + lane) : 0u;
    prefetch_l2_idx(blk, pd.x + 2u * lane);  // 8 bytes per lane cover the 256 bytes of a trip
  };
  auto request = [&](Set& s) {
    s.l1 = load_f4_idx(lpos, s.e1 & kSuperIndexMask);
    s.l2 = load_f4_idx(lpos, s.e2 & kSuperIndexMask);
  };
  // one trip: `cur` holds its entries and (arrived) records, `nxt` the entries of the next trip.  As in k_sweep_img,
  // everything the trip waits for is consumed before it issues new loads (two register sets alternate, no copies).
  auto step = [&](Set& cur, Set& nxt) {
    const uint32_t a1 = cur.e1, a2 = cur.e2, dy = cur.dy;
    if (dy & kNew) {  // a new row starts with this trip
      if (have_row) flush();
      have_row = true;
      const unsigned r = wid + 8u * ((dy >> 16) & 15u);
      k = first + r;
      const float4 li = s_li[r];
      lix = li.x;
      liy = li.y;
      liz = li.z;
      my_abs = __float_as_uint(li.w);
      grp_a = (k < n_a);
      alloc = s_alloc[r];
      if (FILL) rowp = nbr + s_base[r];
      total = total_far = 0u;
    }
    const unsigned rem = dy & kRem;
    float dx1 = cur.l1.x - lix, dy1 = cur.l1.y - liy, dz1 = cur.l1.z - liz;
    float dx2 = cur.l2.x - lix, dy2 = cur.l2.y - liy, dz2 = cur.l2.z - liz;
    if (__any_sync(0xffffffffu, (a1 | a2) > kSuperIndexMask)) {  // some entry is seen through a periodic image
      auto shifted = [&](uint32_t entry, const float4 lj, float& dx, float& dy_, float& dz) {
        const uint32_t code = (entry >> 26) ^ kImageCentre;
        const float wx = (float)((int)(code & 3u) - 1), wy = (float)((int)((code >> 2) & 3u) - 1),
                    wz = (float)((int)((code >> 4) & 3u) - 1);
        dx = lj.x - (lix - (wx * f.box[0] + wy * f.box[3] + wz * f.box[6]));
        dy_ = lj.y - (liy - (wx * f.box[1] + wy * f.box[4] + wz * f.box[7]));
        dz = lj.z - (liz - (wx * f.box[2] + wy * f.box[5] + wz * f.box[8]));
      };
      shifted(a1, cur.l1, dx1, dy1, dz1);
      shifted(a2, cur.l2, dx2, dy2, dz2);
    }
    const bool ne1 = __float_as_uint(cur.l1.w) != my_abs, ne2 = __float_as_uint(cur.l2.w) != my_abs;
    // issue side: records of the next trip, entries of the trip after next
    request(nxt);
    refill(cur);
    const float r1 = fmaf(dz1, dz1, fmaf(dy1, dy1, dx1 * dx1));
    const float r2 = fmaf(dz2, dz2, fmaf(dy2, dy2, dx2 * dx2));
    batch((lane < rem) & ne1 & (r1 < c2_hi), r1, a1);  // j != k already holds in the super-list
    if (rem > 32u) batch((lane + 32u < rem) & ne2 & (r2 < c2_hi), r2, a2);
  };
  Set A, B;
  refill(A);
  refill(B);
  request(A);
  for (;;) {
    if (!(A.dy & kOk)) break;
    step(A, B);
    if (!(B.dy & kOk)) break;
    step(B, A);
  }
  if (have_row) flush();
}

__global__ void k_regular_offsets(unsigned rows, unsigned row_cap, unsigned long long* __restrict__ row_start) {
  const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) row_start[r] = (unsigned long long)r * row_cap;
}

// PAIR style with NLIST: pair k=(k, k+nA) is kept iff within the cutoff at build time (NeighborList.cpp:246-259)
__global__ void k_pair_mask(const double* __restrict__ pos, unsigned n_a, DevPbc pbc, double cutoff2,
                            uint8_t* __restrict__ active) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_a) return;
  const size_t a = 3 * (size_t)k, b = 3 * (size_t)(k + n_a);
  double d[3] = {xsub(pos[b], pos[a]), xsub(pos[b + 1], pos[a + 1]), xsub(pos[b + 2], pos[a + 2])};
  min_image_exact(pbc, d);
  active[k] = norm2_exact(d[0], d[1], d[2]) <= cutoff2 ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// exclusive scan uint32 -> uint64 over n elements, three kernels (block sums, scan of sums, apply)
constexpr int kScanBlock = 1024;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* total_out) {
  __shared__ unsigned long long wsum[32];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned long long x = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (unsigned)o) x += y;
  }
  if (lane == 31) wsum[wid] = x;
  __syncthreads();
  if (wid == 0) {
    unsigned long long t = (lane < nw) ? wsum[lane] : 0ull;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= (unsigned)o) t += y;
    }
    wsum[lane] = t;
  }
  __syncthreads();
  const unsigned long long excl = (wid ? wsum[wid - 1] : 0ull) + (x - v);
  if (total_out) *total_out = wsum[nw - 1];
  __syncthreads();
  return excl;
}

// rows are padded to a multiple of 4 entries: every row then starts 16-byte aligned and the sweep can fetch
// four neighbour indices with one 128-bit load
__device__ __forceinline__ uint32_t padded4(uint32_t c) { return (c + 3u) & ~3u; }

// padq: 0 = plain counts (cells); 3 / 7 = list rows padded to 4 x 32-bit / 8 x 16-bit entries (16-byte rows)
__global__ void k_scan_block_sums(const uint32_t* __restrict__ in, unsigned n, unsigned padq,
                                  unsigned long long* __restrict__ bsum) {
  const unsigned i = blockIdx.x * kScanBlock + threadIdx.x;
  unsigned long long tot;
  const uint32_t v = i < n ? in[i] : 0u;
  block_exclusive_scan((v + padq) & ~padq, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void k_scan_sums(unsigned long long* __restrict__ bsum, unsigned nb, unsigned long long* __restrict__ grand_total) {
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned base = 0; base < nb; base += blockDim.x) {
    const unsigned i = base + threadIdx.x;
    const unsigned long long v = (i < nb) ? bsum[i] : 0ull;
    unsigned long long tot;
    const unsigned long long ex = block_exclusive_scan(v, &tot);
    if (i < nb) bsum[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

// PAD: row offsets (u64, rows padded to 4 entries).  !PAD: cell starts (u32) + cleared placement cursors.
template <bool ROWS>
__global__ void k_scan_apply(const uint32_t* __restrict__ in, unsigned n, unsigned padq,
                             const unsigned long long* __restrict__ bsum, unsigned long long* __restrict__ out64,
                             uint32_t* __restrict__ out32, uint32_t* __restrict__ cursor) {
  const unsigned i = blockIdx.x * kScanBlock + threadIdx.x;
  const uint32_t v = i < n ? in[i] : 0u;
  const unsigned long long ex = block_exclusive_scan((v + padq) & ~padq, nullptr);
  if (i < n) {
    if (ROWS) {
      out64[i] = bsum[blockIdx.x] + ex;
    } else {
      out32[i] = (uint32_t)(bsum[blockIdx.x] + ex);
      cursor[i] = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host-side launch wrappers
void launch_bbox(const double* pos, unsigned n, double* out6, unsigned long long* scratch6, cudaStream_t st) {
  static const unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
  cudaMemcpyAsync(scratch6, init, sizeof(init), cudaMemcpyHostToDevice, st);
  const int blocks = (int)min((n + 255u) / 256u, 148u * 8u);
  k_bbox<<<blocks ? blocks : 1, 256, 0, st>>>(pos, n, out6, scratch6);
  k_bbox_decode<<<1, 32, 0, st>>>(scratch6, out6);
}

void launch_sort(const double* pos, unsigned n, unsigned n_a, int ngroups, const DevGrid& g, uint32_t* cell_of_slot,
                 uint32_t* ccount, uint32_t* cstart, uint32_t* cursor, uint32_t* tmp, uint32_t* perm, uint32_t* scell,
                 unsigned long long* scan_tmp, cudaStream_t st) {
  const unsigned m = (unsigned)ngroups * (unsigned)g.ncell;
  cudaMemsetAsync(ccount, 0, sizeof(uint32_t) * m, st);
  k_bin_atoms<<<(n + 255) / 256, 256, 0, st>>>(pos, n, n_a, g, cell_of_slot, ccount);
  if (m <= 4096u || !scan_tmp) {
    k_scan_cells<<<1, 1024, 0, st>>>(ccount, cstart, cursor, m);
  } else {  // large grids: three-kernel scan instead of one serial block
    const unsigned nb = (m + kScanBlock - 1) / kScanBlock;
    k_scan_block_sums<<<nb, kScanBlock, 0, st>>>(ccount, m, 0u, scan_tmp);
    k_scan_sums<<<1, kScanBlock, 0, st>>>(scan_tmp, nb, nullptr);
    k_scan_apply<false><<<nb, kScanBlock, 0, st>>>(ccount, m, 0u, scan_tmp, nullptr, cstart, cursor);
  }
  k_place_atoms<<<(n + 255) / 256, 256, 0, st>>>(n, n_a, g.ncell, cell_of_slot, cstart, cursor, tmp);
  const unsigned long long threads = (unsigned long long)m * 32ull;
  k_order_cells<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(m, g.ncell, cstart, ccount, tmp, perm, scell);
}

void launch_identity(unsigned n, uint32_t* perm, uint32_t* scell, cudaStream_t st) {
  k_identity_perm<<<(n + 255) / 256, 256, 0, st>>>(n, perm, scell);
}

void launch_gather(const double* pos, const uint32_t* perm, const uint32_t* abs_index, unsigned n, SPos* spos,
                   cudaStream_t st) {
  k_gather_sorted<0, 0><<<(n + 255) / 256, 256, 0, st>>>(pos, perm, abs_index, n, spos, nullptr, DevPbc{}, nullptr);
}
__global__ void k_gather_charges(const double* __restrict__ q, const uint32_t* __restrict__ perm, unsigned n,
                                 double* __restrict__ sq) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) sq[k] = q[perm[k]];
}
void launch_gather_charges(const double* charges, const uint32_t* perm, unsigned n, double* sq, cudaStream_t st) {
  if (n) k_gather_charges<<<(n + 255) / 256, 256, 0, st>>>(charges, perm, n, sq);
}

__global__ void k_gather_types(const uint32_t* __restrict__ t, const uint32_t* __restrict__ perm, unsigned n,
                               uint32_t* __restrict__ st) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) st[k] = t[perm[k]];
}
void launch_gather_types(const uint32_t* types, const uint32_t* perm, unsigned n, uint32_t* stype, cudaStream_t st) {
  if (n) k_gather_types<<<(n + 255) / 256, 256, 0, st>>>(types, perm, n, stype);
}

void launch_gather_track(int track, const double* pos, const uint32_t* perm, const uint32_t* abs_index, unsigned n,
                         SPos* spos, double* bpos, const DevPbc& pbc, unsigned long long* disp2, cudaStream_t st) {
  const unsigned blocks = (n + 255) / 256;
  if (track == 1) {
    k_gather_sorted<1, 0><<<blocks, 256, 0, st>>>(pos, perm, abs_index, n, spos, bpos, pbc, disp2);
  } else {
    switch (pbc.type) {
      case 0: k_gather_sorted<2, 0><<<blocks, 256, 0, st>>>(pos, perm, abs_index, n, spos, bpos, pbc, disp2); break;
      case 1: k_gather_sorted<2, 1><<<blocks, 256, 0, st>>>(pos, perm, abs_index, n, spos, bpos, pbc, disp2); break;
      default: k_gather_sorted<2, 2><<<blocks, 256, 0, st>>>(pos, perm, abs_index, n, spos, bpos, pbc, disp2); break;
    }
  }
}

void launch_nl_rows(bool fill, const SPos* spos, const uint32_t* scell, const uint32_t* cstart, const uint32_t* ccount,
                    const DevGrid& g, const DevPbc& pbc, double cutoff2, unsigned n_a, int two_groups, unsigned row_begin,
                    unsigned row_end, uint32_t* row_count, const unsigned long long* row_start, uint32_t* nbr,
                    cudaStream_t st) {
  const unsigned rows = row_end - row_begin;
  if (!rows) return;
  const unsigned blocks = (unsigned)(((unsigned long long)rows * 32ull + 255ull) / 256ull);
#define B200_NL_LAUNCH(F, P)                                                                                         \
  k_nl_rows<F, P><<<blocks, 256, 0, st>>>(spos, scell, cstart, ccount, g, pbc, cutoff2, n_a, two_groups, row_begin, \
                                          row_end, row_count, row_start, nbr)
  if (fill) {
    if (pbc.type == 0) B200_NL_LAUNCH(true, 0);
    else if (pbc.type == 1) B200_NL_LAUNCH(true, 1);
    else B200_NL_LAUNCH(true, 2);
  } else {
    if (pbc.type == 0) B200_NL_LAUNCH(false, 0);
    else if (pbc.type == 1) B200_NL_LAUNCH(false, 1);
    else B200_NL_LAUNCH(false, 2);
  }
#undef B200_NL_LAUNCH
}

void launch_sort_init(const double* pos, const uint32_t* perm, const uint32_t* abs_index, unsigned n, const DevGrid& g,
                      const DevPbc& box, bool use_wrapped, float4* lpos, double* wpos, double* braw, SPos* spos, double* ubuild,
                      cudaStream_t st) {
  if (n)
    k_sort_init<<<(n + 255) / 256, 256, 0, st>>>(pos, perm, abs_index, n, g, box, use_wrapped ? 1 : 0, lpos, wpos, braw, spos,
                                                  ubuild);
}
void launch_gather_u(bool build, const PosSrc& pos, const uint32_t* perm, const uint32_t* abs_index, const IdxRanges& rng,
                     const double* wpos, const double* braw, const DevPbc& pbc, SPos* spos, double* ubuild, float4* lpos,
                     unsigned long long* disp2, cudaStream_t st) {
  if (!rng.total) return;
  const unsigned blocks = (rng.total + 255) / 256;
#define B200_GU(P, B) k_gather_u<P, B><<<blocks, 256, 0, st>>>(pos, perm, abs_index, rng, wpos, braw, pbc, spos, ubuild, lpos, disp2)
  if (build) {
    switch (pbc.type) {
      case 0: B200_GU(0, true); break;
      case 1: B200_GU(1, true); break;
      default: B200_GU(2, true); break;
    }
  } else {
    switch (pbc.type) {
      case 0: B200_GU(0, false); break;
      case 1: B200_GU(1, false); break;
      default: B200_GU(2, false); break;
    }
  }
#undef B200_GU
}
__global__ void k_invert_perm(const uint32_t* __restrict__ perm, unsigned n, uint32_t* __restrict__ inv) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) inv[perm[k]] = k;
}
void launch_invert_perm(const uint32_t* perm, unsigned n, uint32_t* inv, cudaStream_t st) {
  if (n) k_invert_perm<<<(n + 255) / 256, 256, 0, st>>>(perm, n, inv);
}

// ------------------------------------------------------------------------------------------------
// The list as a tool for other consumers: (i0, i1) pairs of indices into the caller's position array, each pair once
// in the reference's orientation (NeighborList::getIndexPair: GROUPA atom first, resp. lower index first).
// One warp per row; FILL = false counts the pairs a row emits, FILL = true writes them at the row's offset.
template <bool FILL>
__global__ void __launch_bounds__(256)
    k_export_pairs(unsigned rows, unsigned row_begin, unsigned n_a, int two_groups, const uint32_t* __restrict__ perm,
                   const unsigned long long* __restrict__ row_start, const uint32_t* __restrict__ row_count,
                   const uint32_t* __restrict__ far_off, const uint32_t* __restrict__ far_cnt, const uint32_t* __restrict__ nbr,
                   uint32_t idx_mask, uint32_t* __restrict__ emit_count, const unsigned long long* __restrict__ emit_start,
                   unsigned* __restrict__ pairs, unsigned long long capacity) {
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const unsigned k = row_begin + warp;
  const uint32_t si = perm[k];
  const uint32_t* __restrict__ row = nbr + row_start[warp];
  const unsigned near = row_count[warp], far = far_cnt ? far_cnt[warp] : 0u, foff = far_off ? far_off[warp] : 0u;
  unsigned total = 0;
  const unsigned long long base = FILL ? emit_start[warp] : 0ull;
  for (unsigned e0 = 0; e0 < near + far; e0 += 32) {
    const unsigned e = e0 + lane;
    bool keep = false;
    uint32_t sj = 0;
    if (e < near + far) {
      const uint32_t j = (e < near ? row[e] : row[foff + (e - near)]) & idx_mask;
      sj = perm[j];
      keep = two_groups ? (k < n_a) : (si < sj);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (FILL && keep) {
      const unsigned long long at = base + total + __popc(mask & ((1u << lane) - 1u));
      if (at < capacity) {
        pairs[2 * at] = si;
        pairs[2 * at + 1] = sj;
      }
    }
    total += __popc(mask);
  }
  if (!FILL && lane == 0) emit_count[warp] = total;
}

void launch_export_pairs(bool fill, unsigned rows, unsigned row_begin, unsigned n_a, int two_groups, const uint32_t* perm,
                         const unsigned long long* row_start, const uint32_t* row_count, const uint32_t* far_off,
                         const uint32_t* far_cnt, const uint32_t* nbr, uint32_t idx_mask, uint32_t* emit_count,
                         const unsigned long long* emit_start, unsigned* pairs, unsigned long long capacity, cudaStream_t st) {
  if (!rows) return;
  const unsigned blocks = (unsigned)(((unsigned long long)rows * 32ull + 255ull) / 256ull);
  if (fill)
    k_export_pairs<true><<<blocks, 256, 0, st>>>(rows, row_begin, n_a, two_groups, perm, row_start, row_count, far_off, far_cnt,
                                                 nbr, idx_mask, emit_count, emit_start, pairs, capacity);
  else
    k_export_pairs<false><<<blocks, 256, 0, st>>>(rows, row_begin, n_a, two_groups, perm, row_start, row_count, far_off, far_cnt,
                                                  nbr, idx_mask, emit_count, emit_start, pairs, capacity);
}

void launch_pack_meta(unsigned rows, const unsigned long long* row_start, const uint32_t* row_count, const uint32_t* far_off,
                      const uint32_t* far_cnt, uint4* meta, unsigned long long* listed, cudaStream_t st) {
  cudaMemsetAsync(listed, 0, sizeof(unsigned long long), st);
  if (rows) k_pack_meta<<<(rows + 255) / 256, 256, 0, st>>>(rows, row_start, row_count, far_off, far_cnt, meta, listed);
}

void launch_nl_rows_f32(int mode /*0 count, 1 fill, 2 capped single pass*/, bool super, bool images, const double* pos,
                        const uint32_t* perm, const float4* lpos, const uint32_t* scell, const uint32_t* cstart, const uint32_t* ccount,
                        const DevGrid& g, const DevPbc* pbc_g, const DevPbc& box, double cutoff2, double band_rel,
                        unsigned n_a, int two_groups, unsigned row_begin, unsigned row_end, uint32_t* row_count,
                        unsigned long long* row_start, uint32_t* nbr, unsigned row_cap, unsigned* cap_info, float far2,
                        uint32_t* row_far_off, uint32_t* row_far_cnt, cudaStream_t st) {
  const unsigned rows = row_end - row_begin;
  if (!rows) return;
  const unsigned blocks = (unsigned)(((unsigned long long)rows * 32ull + 255ull) / 256ull);
  SearchF32 f;
  for (int i = 0; i < 9; ++i) f.box[i] = (float)box.box[i];
  f.c2_hi = (float)(cutoff2 * (1.0 + band_rel));
  f.c2_lo = (float)(cutoff2 * (1.0 - band_rel));
#define B200_F32_ARGS pos, perm, lpos, scell, cstart, ccount, g, pbc_g, f, cutoff2, n_a, two_groups, row_begin, row_end, \
                      row_count, row_start, nbr, row_cap, cap_info, far2, row_far_off, row_far_cnt
  if (mode == 2) k_regular_offsets<<<(rows + 255) / 256, 256, 0, st>>>(rows, row_cap, row_start);
  // resident blocks per SM the register allocation is sized for: 4 -> 64 registers (spills ~120 bytes), 3 -> 80
  // (60 bytes), 2 -> no spills
  static const int minb = [] {
    const char* e = std::getenv("B200COORD_ROWS_MINB");
    const int v = e ? std::atoi(e) : 3;
    return (v == 2 || v == 4) ? v : 3;  // measured at 1 M atoms, whole rebuild: 7.93 ms (4), 7.46 ms (3), 9.03 ms (2)
  }();
#define B200_F32_GO(MB)                                                                                          \
  do {                                                                                                           \
    if (super) { /* super-list rows: always with images */                                                       \
      if (mode == 0) k_nl_rows_f32<false, false, true, true, MB><<<blocks, 256, 0, st>>>(B200_F32_ARGS);          \
      else if (mode == 1) k_nl_rows_f32<true, false, true, true, MB><<<blocks, 256, 0, st>>>(B200_F32_ARGS);      \
      else k_nl_rows_f32<true, true, true, true, MB><<<blocks, 256, 0, st>>>(B200_F32_ARGS);                      \
    } else if (mode == 0) k_nl_rows_f32<false, false, false, false, MB><<<blocks, 256, 0, st>>>(B200_F32_ARGS);   \
    else if (images) {                                                                                           \
      if (mode == 1) k_nl_rows_f32<true, false, false, true, MB><<<blocks, 256, 0, st>>>(B200_F32_ARGS);          \
      else k_nl_rows_f32<true, true, false, true, MB><<<blocks, 256, 0, st>>>(B200_F32_ARGS);                     \
    } else {                                                                                                     \
      if (mode == 1) k_nl_rows_f32<true, false, false, false, MB><<<blocks, 256, 0, st>>>(B200_F32_ARGS);         \
      else k_nl_rows_f32<true, true, false, false, MB><<<blocks, 256, 0, st>>>(B200_F32_ARGS);                    \
    }                                                                                                            \
  } while (0)
  if (minb == 2) B200_F32_GO(2);
  else if (minb == 4) B200_F32_GO(4);
  else B200_F32_GO(3);
#undef B200_F32_GO
#undef B200_F32_ARGS
}

void launch_nl_filter(int mode /*0 count, 1 fill, 2 capped single pass*/, bool images, const double* pos, const uint32_t* perm,
                      const float4* lpos,
                      const unsigned long long* srow_start, const uint32_t* srow_count, const uint32_t* snbr, const DevPbc* pbc_g,
                      const DevPbc& box, double cutoff2, double band_rel, unsigned n_a, int two_groups, unsigned row_begin,
                      unsigned row_end, uint32_t* row_count, unsigned long long* row_start, uint32_t* nbr, unsigned row_cap,
                      unsigned* cap_info, float far2, uint32_t* row_far_off, uint32_t* row_far_cnt, unsigned max_srow,
                      int flat_minb, cudaStream_t st) {
  const unsigned rows = row_end - row_begin;
  if (!rows) return;
  const unsigned blocks = (unsigned)(((unsigned long long)rows * 32ull + 255ull) / 256ull);
  SearchF32 f;
  for (int i = 0; i < 9; ++i) f.box[i] = (float)box.box[i];
  f.c2_hi = (float)(cutoff2 * (1.0 + band_rel));
  f.c2_lo = (float)(cutoff2 * (1.0 - band_rel));
  // one pipeline per warp across rows when the rows are short enough for its trip table (max_srow = longest
  // super-list row, 0 = unknown / switched off)
  unsigned per_warp = 0u;
  if (max_srow > 0u && max_srow <= 0xffffu) per_warp = std::min<unsigned>(kFltRowsPerWarp, kFltTripCap / ((max_srow + 63u) / 64u));
  if (per_warp) {
    const unsigned rpb = 8u * per_warp;
    const unsigned fb = (rows + rpb - 1u) / rpb;
#define B200_FLAT_ARGS pos, perm, lpos, srow_start, srow_count, snbr, pbc_g, f, cutoff2, n_a, two_groups, row_begin, row_end, \
                       row_count, row_start, nbr, row_cap, cap_info, far2, row_far_off, row_far_cnt, rpb
    if (mode == 2) k_regular_offsets<<<(rows + 255) / 256, 256, 0, st>>>(rows, row_cap, row_start);
#define B200_FLAT_GO(MB)                                                                                     \
    do {                                                                                                       \
      if (mode == 0) k_nl_filter_flat<false, false, false, MB><<<fb, 256, 0, st>>>(B200_FLAT_ARGS);            \
      else if (images) {                                                                                       \
        if (mode == 1) k_nl_filter_flat<true, false, true, MB><<<fb, 256, 0, st>>>(B200_FLAT_ARGS);            \
        else k_nl_filter_flat<true, true, true, MB><<<fb, 256, 0, st>>>(B200_FLAT_ARGS);                       \
      } else {                                                                                                 \
        if (mode == 1) k_nl_filter_flat<true, false, false, MB><<<fb, 256, 0, st>>>(B200_FLAT_ARGS);           \
        else k_nl_filter_flat<true, true, false, MB><<<fb, 256, 0, st>>>(B200_FLAT_ARGS);                      \
      }                                                                                                        \
    } while (0)
    // 2 resident blocks per SM: 122 registers, nothing spilled; 3: 80 registers, ~12 spill accesses per trip
    if (flat_minb == 3) B200_FLAT_GO(3);
    else B200_FLAT_GO(2);
#undef B200_FLAT_GO
#undef B200_FLAT_ARGS
    return;
  }
#define B200_FLT_ARGS pos, perm, lpos, srow_start, srow_count, snbr, pbc_g, f, cutoff2, n_a, two_groups, row_begin, row_end, \
                      row_count, row_start, nbr, row_cap, cap_info, far2, row_far_off, row_far_cnt
  if (mode == 2) k_regular_offsets<<<(rows + 255) / 256, 256, 0, st>>>(rows, row_cap, row_start);
  if (mode == 0) k_nl_filter<false, false, false><<<blocks, 256, 0, st>>>(B200_FLT_ARGS);
  else if (images) {
    if (mode == 1) k_nl_filter<true, false, true><<<blocks, 256, 0, st>>>(B200_FLT_ARGS);
    else k_nl_filter<true, true, true><<<blocks, 256, 0, st>>>(B200_FLT_ARGS);
  } else {
    if (mode == 1) k_nl_filter<true, false, false><<<blocks, 256, 0, st>>>(B200_FLT_ARGS);
    else k_nl_filter<true, true, false><<<blocks, 256, 0, st>>>(B200_FLT_ARGS);
  }
#undef B200_FLT_ARGS
}

void launch_pair_mask(const double* pos, unsigned n_a, const DevPbc& pbc, double cutoff2, uint8_t* active, cudaStream_t st) {
  if (n_a) k_pair_mask<<<(n_a + 255) / 256, 256, 0, st>>>(pos, n_a, pbc, cutoff2, active);
}

void launch_scan_rows(const uint32_t* row_count, unsigned rows, unsigned padq, unsigned long long* bsum,
                      unsigned long long* row_start, unsigned long long* grand_total, cudaStream_t st) {
  if (!rows) {
    cudaMemsetAsync(grand_total, 0, sizeof(unsigned long long), st);
    return;
  }
  const unsigned nb = (rows + kScanBlock - 1) / kScanBlock;
  k_scan_block_sums<<<nb, kScanBlock, 0, st>>>(row_count, rows, padq, bsum);
  k_scan_sums<<<1, kScanBlock, 0, st>>>(bsum, nb, grand_total);
  k_scan_apply<true><<<nb, kScanBlock, 0, st>>>(row_count, rows, padq, bsum, row_start, nullptr, nullptr);
}

}  // namespace b200
