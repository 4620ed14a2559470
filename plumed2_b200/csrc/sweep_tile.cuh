// Cell-tile sweep for NLISTCELLS and "no list" (the reference's superset semantics: every pair of the <= 27 stencil
// cells of the frozen binning, resp. every pair; NeighborList.cpp:177-236, :133-140; CoordinationBase.cpp:177-208).
//
// One block = one chunk of <= 128 rows (i-atoms) of ONE cell.  All rows of a cell see the same partners -- the <= 9
// contiguous sorted ranges of the stencil cells (kernels.cuh: for_each_stencil_range) -- so the block stages them ONCE
// into shared memory with 1-D bulk copies (cp.async.bulk, completion on an mbarrier: the TMA engine moves the bytes, no
// thread touches them) and every row reuses them from there.  The warp-per-row kernel this replaces re-gathered the
// same records from L1/L2 for every row and spent a full warp pass on each short range.
//
// Work inside the block is cut into units (row, 128 staged partners) that the 8 warps take in turn, so a cell with
// one i-atom and 2700 partners (a solute atom in water) keeps all warps busy, and so does a cell with 100 i-atoms
// and 27 partners (the water around it).  Unit sums are combined in a fixed order: results do not depend on
// scheduling.  Pair arithmetic = pair_term (sweep_math.cuh), the same function the row kernels use.
#pragma once
#include "sweep_math.cuh"

namespace b200 {

constexpr int kTileRows = 128;       // rows per block at most (a whole cell of water)
constexpr int kTileSub = 32;         // rows whose units are in flight together
constexpr int kTileCap = 2048;       // staged partner records per pass (64 KB)
constexpr int kTileUnit = 128;       // partners per unit
constexpr int kTileUnitsMax = kTileSub * (kTileCap / kTileUnit);

struct TileWork {
  uint32_t seg;    // group * ncell + cell
  uint32_t row0;   // first row (sorted index) of the chunk
  uint32_t nrows;  // rows of the chunk
  uint32_t pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Waits for the phase; a transfer that never completes (a bug, not a load) traps instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
  for (unsigned spins = 0;; ++spins) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    if (done) return;
    if (spins > (1u << 24)) __trap();
  }
}
// 1-D bulk copy global -> shared, bytes a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int K, int PBC, bool ACC>
__global__ void __launch_bounds__(kSweepThreads, 2)
    k_sweep_tile(SweepArgs a, DevPbc pbc, DevSwitch sw, const TileWork* __restrict__ work,
                 const unsigned long long* __restrict__ first_dev, const unsigned long long* __restrict__ end_dev) {
  extern __shared__ __align__(128) unsigned char tile_raw[];
  SPos* tile = reinterpret_cast<SPos*>(tile_raw);
  __shared__ uint64_t bar;
  // stencil ranges (first sorted index, count): 9 (y,z) columns, each one run or two where it wraps around the box
  constexpr int kMaxRanges = 20;
  __shared__ uint32_t s_rs[kMaxRanges], s_rm[kMaxRanges];
  __shared__ int s_nr;
  __shared__ double s_part[kTileUnitsMax][3];
  __shared__ double s_row[kTileRows][3];
  // the launch covers an upper bound of the item count (known on the device only): surplus blocks leave at once
  const unsigned long long item = *first_dev + blockIdx.x;
  if (item >= *end_dev) return;
  const TileWork w = work[item];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const DevGrid& g = a.grid;
  const unsigned my_grp = w.seg / (unsigned)g.ncell;
  const unsigned other = a.two_groups ? (1u - my_grp) : 0u;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  if (wid == 0) {
    // range table: lane = one (y, z) column of the stencil; its x-run is one sorted range, or two where it wraps around
    // the box.  All the dependent cstart / ccount loads of the block are in flight at once (a serial walk costs ~27
    // L2 round trips per block).
    int c[3], lo[3], hi[3];
    cell_coords(g, (int)(w.seg % (unsigned)g.ncell), c);
    stencil_bounds(g, c, lo, hi);
    const int nyn = hi[1] - lo[1], nzn = hi[2] - lo[2];
    uint32_t sA = 0, mA = 0, sB = 0, mB = 0;
    if ((int)lane < nyn * nzn) {
      const int ny = lo[1] + (int)lane / nzn, nz = lo[2] + (int)lane % nzn;
      const unsigned cbase = other * (unsigned)g.ncell + (unsigned)(wrap_cell(ny, g.n[1]) * g.n[0] + wrap_cell(nz, g.n[2]) * g.n[0] * g.n[1]);
      int x = lo[0];
      {
        const int xw = wrap_cell(x, g.n[0]);
        const int run = min(hi[0] - x, g.n[0] - xw);
        const unsigned f0 = cbase + (unsigned)xw, l0 = f0 + (unsigned)run - 1u;
        sA = a.cstart[f0];
        mA = a.cstart[l0] + a.ccount[l0] - sA;
        x += run;
      }
      if (x < hi[0]) {
        const int xw = wrap_cell(x, g.n[0]);
        const int run = min(hi[0] - x, g.n[0] - xw);
        const unsigned f0 = cbase + (unsigned)xw, l0 = f0 + (unsigned)run - 1u;
        sB = a.cstart[f0];
        mB = a.cstart[l0] + a.ccount[l0] - sB;
      }
    }
    // compact the non-empty ranges in lane order (A before B): same order as for_each_stencil_range
    const unsigned hasA = __ballot_sync(0xffffffffu, mA != 0u), hasB = __ballot_sync(0xffffffffu, mB != 0u);
    const unsigned below = (1u << lane) - 1u;
    const unsigned posA = __popc(hasA & below) + __popc(hasB & below);
    if (mA && posA < (unsigned)kMaxRanges) {
      s_rs[posA] = sA;
      s_rm[posA] = mA;
    }
    if (mB && posA + (mA ? 1u : 0u) < (unsigned)kMaxRanges) {
      s_rs[posA + (mA ? 1u : 0u)] = sB;
      s_rm[posA + (mA ? 1u : 0u)] = mB;
    }
    if (lane == 0) s_nr = min(__popc(hasA) + __popc(hasB), kMaxRanges);
  }
  for (unsigned t = threadIdx.x; t < (unsigned)kTileRows * 3u; t += kSweepThreads) (&s_row[0][0])[t] = 0.0;
  __syncthreads();
  const int nr = s_nr;
  unsigned P = 0;
  for (int q = 0; q < nr; ++q) P += s_rm[q];

  LaneAcc acc = {0, 0, 0, 0, 0, 0, 0};
  unsigned phase = 0;
  for (unsigned base = 0; base < P; base += kTileCap) {
    const unsigned np = min((unsigned)kTileCap, P - base);  // partners staged in this pass
    if (threadIdx.x == 0) {
      // generic-proxy reads of the previous pass are ordered before the async-proxy writes of this one
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar, np * (unsigned)sizeof(SPos));
      unsigned off = 0;  // flattened index of the range's first partner
      for (int q = 0; q < nr; ++q) {
        const unsigned lo = max(off, base), hi = min(off + s_rm[q], base + np);
        if (lo < hi) bulk_g2s(tile + (lo - base), a.spos + s_rs[q] + (lo - off), (hi - lo) * (unsigned)sizeof(SPos), &bar);
        off += s_rm[q];
      }
    }
    mbar_wait(&bar, phase);
    phase ^= 1u;
    const unsigned nq = (np + kTileUnit - 1) / kTileUnit;
    for (unsigned r0 = 0; r0 < w.nrows; r0 += kTileSub) {  // the staged partners serve every row of the cell
    const unsigned nsub = min((unsigned)kTileSub, w.nrows - r0);
    const unsigned units = nsub * nq;
    for (unsigned u = wid; u < units; u += kSweepWarps) {
      const unsigned r = u / nq, q = u - r * nq;
      const SPos pi = load_spos(a.spos + w.row0 + r0 + r);
      const unsigned long long wi = ((unsigned long long)pi.slot << 32) | pi.abs_index;
      double fx = 0.0, fy = 0.0, fz = 0.0;
      bool unused_near = false;
      const unsigned p_end = min(np, (q + 1) * (unsigned)kTileUnit);
      for (unsigned p = q * kTileUnit + lane; p < p_end; p += 32) {
        const SPos pj = tile[p];
        const unsigned long long wj = ((unsigned long long)pj.slot << 32) | pj.abs_index;
        const bool valid = (wj != wi) && (!a.check_abs || pj.abs_index != pi.abs_index);
        const bool flip = a.two_groups ? (my_grp == 1u) : (pi.slot > pj.slot);
        if (valid) {
          if (ACC && a.scatter_b) {
            // few GROUPA atoms among many GROUPB atoms (kernels.cuh: scatter_b): this GROUPA row hands +dd to its
            // partner (deriv[i1] += dd, CoordinationBase.cpp:201) and the GROUPB rows are not swept
            double tx = 0.0, ty = 0.0, tz = 0.0;
            pair_term<K, PBC, ACC, true>(pbc, sw, unused_near, pi.x, pi.y, pi.z, pj, flip, tx, ty, tz, acc);
            fx += tx;
            fy += ty;
            fz += tz;
            if (tx != 0.0 || ty != 0.0 || tz != 0.0) {
              double* dj = a.sderiv + 3 * (size_t)a.inv[pj.slot];
              atomicAdd(dj, -tx);
              atomicAdd(dj + 1, -ty);
              atomicAdd(dj + 2, -tz);
            }
          } else {
            pair_term<K, PBC, ACC, true>(pbc, sw, unused_near, pi.x, pi.y, pi.z, pj, flip, fx, fy, fz, acc);
          }
        }
      }
      fx = warp_sum(fx);
      fy = warp_sum(fy);
      fz = warp_sum(fz);
      if (lane == 0) {
        s_part[u][0] = fx;
        s_part[u][1] = fy;
        s_part[u][2] = fz;
      }
    }
    __syncthreads();
    // fixed-order combine: thread (r, c) adds the units of row r
    if (threadIdx.x < nsub * 3u) {
      const unsigned r = threadIdx.x / 3u, c = threadIdx.x - 3u * r;
      double t = s_row[r0 + r][c];
      for (unsigned q = 0; q < nq; ++q) t += s_part[r * nq + q][c];
      s_row[r0 + r][c] = t;
    }
    __syncthreads();
    }
  }
  for (unsigned t = threadIdx.x; t < w.nrows * 3u; t += kSweepThreads) a.sderiv[3 * (size_t)w.row0 + t] = (&s_row[0][0])[t];
  const unsigned long long evals = (threadIdx.x == 0) ? (unsigned long long)w.nrows * P : 0ull;
  if (threadIdx.x == 0 && evals) atomicAdd(a.executed, evals);
  if (ACC) {
    block_store_partials(acc, evals, a.partials, a.evals);  // record blockIdx.x
  } else if (threadIdx.x == 0 && evals) {
    atomicAdd(a.evals, evals);
  }
}

// ---- work list: one item per chunk of <= rows_per_item rows of a (group, cell) segment, rows restricted to the rank
__global__ void k_tile_count(unsigned nseg, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ ccount,
                             unsigned row_begin, unsigned row_end, unsigned rows_per_item, uint32_t* __restrict__ nitems) {
  const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const unsigned lo = max(cstart[s], row_begin), hi = min(cstart[s] + ccount[s], row_end);
  nitems[s] = (hi > lo) ? (hi - lo + rows_per_item - 1) / rows_per_item : 0u;
}
__global__ void k_tile_fill(unsigned nseg, const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ ccount,
                            unsigned row_begin, unsigned row_end, unsigned rows_per_item,
                            const unsigned long long* __restrict__ item_start, TileWork* __restrict__ work) {
  const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const unsigned lo = max(cstart[s], row_begin), hi = min(cstart[s] + ccount[s], row_end);
  unsigned long long at = item_start[s];
  for (unsigned r = lo; r < hi; r += rows_per_item) {
    TileWork w;
    w.seg = s;
    w.row0 = r;
    w.nrows = min(rows_per_item, hi - r);
    w.pad = 0;
    work[at++] = w;
  }
}

}  // namespace b200
