// Tile sweep: the pair sweep of kernels_sweep.cu with the partner records staged in shared memory.
//
// Why: the list sweep is bound by the L1 gather path -- a warp-wide gather of 32-byte records touches ~20
// distinct 128-byte lines and L1tex replays one line every ~2 cycles (profiles/: FP64 pipe only ~50 % busy, more
// warps or deeper prefetch do not help).  Shared memory serves the same random 8-byte reads at ~3x the rate
// and at a fixed ~30-cycle latency, so the FP64 pipe becomes the bound it is supposed to be.
//
// How: a block owns a PENCIL of consecutive cells along x (6 half-cutoff cells, or one cutoff-sized cell).  All
// partner atoms any of its rows can see lie in the pencil's stencil: <= 25 columns (dy,dz), each ONE contiguous
// range of the sorted array (two when the run wraps around the box).  The block copies those ranges into shared
// memory with coalesced 256-bit loads -- every record is read from L2 once per pencil -- and the neighbour list
// holds 16-bit indices into that tile (built by k_nl_rows_f32<.., TILE=true> with the same layout code), which
// also halves the list traffic and footprint.  The arithmetic, the canonical pair orientation, the reductions
// and the optional peer-store exchange are exactly those of k_sweep_list.
#include "sweep_math.cuh"

namespace b200 {

constexpr int kTileThreads = 512;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileMaxCols = 32;

template <int K, int PBC, bool ACC>
__global__ void __launch_bounds__(kTileThreads, 2)
    k_sweep_tile(SweepArgs a, DevPbc pbc, DevSwitch sw, unsigned tile_cap, unsigned row_group /*0: A rows, 1: B rows*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* __restrict__ tx = reinterpret_cast<double*>(smem_raw);
  double* __restrict__ ty = tx + tile_cap;
  double* __restrict__ tz = ty + tile_cap;
  uint32_t* __restrict__ tslot = reinterpret_cast<uint32_t*>(tz + tile_cap);
  __shared__ TileCol s_col[kTileMaxCols];
  __shared__ uint32_t s_off[kTileMaxCols];
  __shared__ unsigned s_rows[2];  // first, last (exclusive) row of this block

  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const DevGrid& g = a.grid;
  const int P = pencil_cells(g);
  const int npx = (g.n[0] + P - 1) / P;
  // pencil of this block: blockIdx -> (xp, y, z)
  const int pid = (int)blockIdx.x;
  const int xp = pid % npx, yz = pid / npx;
  const int cy = yz % g.n[1], cz = yz / g.n[1];
  const int x0 = xp * P, x1 = min(x0 + P, g.n[0]) - 1;
  const unsigned row_off = row_group * (unsigned)g.ncell;                               // my rows' cell table
  const unsigned part_off = (a.two_groups ? (1u - row_group) : 0u) * (unsigned)g.ncell;  // partners' cell table

  if (wid == 0) {
    int c[3] = {x0, cy, cz}, lo[3], hi[3];
    stencil_bounds(g, c, lo, hi);
    const int ny_n = hi[1] - lo[1], nz_n = hi[2] - lo[2];
    const int ncol = ny_n * nz_n;
    TileCol pc;
    pc.gA = pc.lA = pc.gB = pc.lB = 0u;
    pc.wA = pc.wB = 0;
    if ((int)lane < ncol) {
      const int ny = lo[1] + (int)lane / nz_n, nz = lo[2] + (int)lane % nz_n;
      const unsigned cbase = part_off + (unsigned)(wrap_cell(ny, g.n[1]) * g.n[0] + wrap_cell(nz, g.n[2]) * g.n[0] * g.n[1]);
      int xa, xb;
      xrun_bounds(g, x0, x1, xa, xb);
      column_parts(g, cbase, xa, xb, a.cstart, a.ccount, pc);
    }
    uint32_t tile_total;
    const uint32_t off = warp_exclusive_scan(pc.lA + pc.lB, lane, tile_total);
    s_col[lane] = pc;
    s_off[lane] = off;
    if (lane == 0) {
      const unsigned cf = row_off + (unsigned)(x0 + cy * g.n[0] + cz * g.n[0] * g.n[1]);
      const unsigned cl = cf + (unsigned)(x1 - x0);
      unsigned first = a.cstart[cf], last = a.cstart[cl] + a.ccount[cl];
      first = max(first, a.row_begin);
      last = min(last, a.row_end);
      s_rows[0] = first;
      s_rows[1] = max(first, last);
    }
  }
  __syncthreads();
  const unsigned first = s_rows[0], last = s_rows[1];
  if (first >= last) return;  // empty pencil (or not this rank's): uniform exit, nothing staged

  // ---- stage the tile: warp w copies columns w, w+8, ... (coalesced 256-bit loads, many in flight)
  for (int col = (int)wid; col < kTileMaxCols; col += kTileWarps) {
    const TileCol pc = s_col[col];
    const uint32_t off = s_off[col];
    for (uint32_t t = lane; t < pc.lA; t += 32) {
      const SPos r = load_spos(a.spos + pc.gA + t);
      tx[off + t] = r.x;
      ty[off + t] = r.y;
      tz[off + t] = r.z;
      tslot[off + t] = r.slot;
    }
    for (uint32_t t = lane; t < pc.lB; t += 32) {
      const SPos r = load_spos(a.spos + pc.gB + t);
      tx[off + pc.lA + t] = r.x;
      ty[off + pc.lA + t] = r.y;
      tz[off + pc.lA + t] = r.z;
      tslot[off + pc.lA + t] = r.slot;
    }
  }
  __syncthreads();

  LaneAcc acc = {0, 0, 0, 0, 0, 0, 0};
  unsigned long long evals = 0;
  const bool row_is_b = (row_group == 1u);
  for (unsigned k = first + wid; k < last; k += kTileWarps) {
    const SPos pi = load_spos(a.spos + k);
    const unsigned long long base = a.row_start[k - a.row_begin];
    const unsigned cnt = a.row_count[k - a.row_begin];
    const uint16_t* __restrict__ row = a.nbr16 + base;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    bool near = false;
    // two pairs per lane and trip as independent FP64 chains (see k_sweep_list); indices one trip ahead
    unsigned e = lane;
    unsigned ja = (e < cnt) ? row[e] : 0u;
    unsigned jb = (e + 32 < cnt) ? row[e + 32] : 0u;
    for (; e < cnt; e += 64) {
      SPos pa, pb;
      pa.x = tx[ja];
      pa.y = ty[ja];
      pa.z = tz[ja];
      pa.slot = tslot[ja];
      pb.x = tx[jb];
      pb.y = ty[jb];
      pb.z = tz[jb];
      pb.slot = tslot[jb];
      const bool vb = (e + 32 < cnt);
      ja = (e + 64 < cnt) ? row[e + 64] : 0u;
      jb = (e + 96 < cnt) ? row[e + 96] : 0u;
      const bool flipa = a.two_groups ? row_is_b : (pi.slot > pa.slot);
      const bool flipb = a.two_groups ? row_is_b : (pi.slot > pb.slot);
      pair_term2<K, PBC, ACC>(pbc, sw, near, pi.x, pi.y, pi.z, pa, flipa, pb, flipb, vb, fx, fy, fz, acc);
    }
    if (__any_sync(0xffffffffu, near)) {
      const RowFix f = row_fixup_tile<K, PBC>(a.pbc_g, a.sw_g, a.spos, tx, ty, tz, tslot, row, cnt, k, lane, a.two_groups, row_is_b);
      apply_fix(f, ACC, fx, fy, fz, acc);
    }
    fx = warp_sum(fx);
    fy = warp_sum(fy);
    fz = warp_sum(fz);
    if (lane == 0) {
      a.sderiv[3 * (size_t)k] = fx;
      a.sderiv[3 * (size_t)k + 1] = fy;
      a.sderiv[3 * (size_t)k + 2] = fz;
      evals += cnt;
    }
  }
  if (a.npeers) {  // fused exchange, as in k_sweep_list
    __syncthreads();
    if ((int)wid < a.npeers) {
      double* __restrict__ q = a.peers[wid] + 3 * (size_t)first;
      const double* __restrict__ mine = a.sderiv + 3 * (size_t)first;
      const unsigned m = 3u * (last - first);
      for (unsigned t = lane; t < m; t += 32) q[t] = mine[t];
    }
  }
  if (ACC)
    block_store_partials(acc, evals, a.partials, a.evals);
  else if (lane == 0 && evals)
    atomicAdd(a.evals, evals);
}

// empty pencils return before block_store_partials: their partial records must read as zero
__global__ void k_zero_partials(double* __restrict__ partials, unsigned n) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) partials[i] = 0.0;
}

template <int K, int PBC>
static int run_tile(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, unsigned tile_cap, cudaStream_t st) {
  const DevGrid& g = a.grid;
  const int P = g.pencil > 0 ? g.pencil : 1;
  const int npencil = ((g.n[0] + P - 1) / P) * g.n[1] * g.n[2];
  const size_t smem = (size_t)tile_cap * 28u + 16u;
  static bool configured_acc = false, configured_noacc = false;
  if (!configured_acc) {
    if (cudaFuncSetAttribute(k_sweep_tile<K, PBC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return -2;
    configured_acc = true;
  }
  k_zero_partials<<<(npencil * kPartialStride + 255) / 256, 256, 0, st>>>(a.partials, (unsigned)(npencil * kPartialStride));
  k_sweep_tile<K, PBC, true><<<npencil, kTileThreads, smem, st>>>(a, pbc, sw, tile_cap, 0u);
  if (a.two_groups) {
    if (!configured_noacc) {
      if (cudaFuncSetAttribute(k_sweep_tile<K, PBC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return -2;
      configured_noacc = true;
    }
    k_sweep_tile<K, PBC, false><<<npencil, kTileThreads, smem, st>>>(a, pbc, sw, tile_cap, 1u);
  }
  return npencil;
}

template <int K>
static int run_tile_pbc(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, unsigned tile_cap, cudaStream_t st) {
  switch (pbc.type) {
    case 0: return run_tile<K, 0>(a, pbc, sw, tile_cap, st);
    case 1: return run_tile<K, 1>(a, pbc, sw, tile_cap, st);
    default: return run_tile<K, 2>(a, pbc, sw, tile_cap, st);
  }
}

int launch_sweep_tile(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, unsigned tile_cap, cudaStream_t st) {
  if ((size_t)tile_cap * 28u + 16u > 200u * 1024u) return -2;
  switch (kind_of(sw.type)) {
    case K_FIX6: return run_tile_pbc<K_FIX6>(a, pbc, sw, tile_cap, st);
    case K_FIXN: return run_tile_pbc<K_FIXN>(a, pbc, sw, tile_cap, st);
    case K_RAT_R2: return run_tile_pbc<K_RAT_R2>(a, pbc, sw, tile_cap, st);
    case K_RAT_R: return run_tile_pbc<K_RAT_R>(a, pbc, sw, tile_cap, st);
    case K_EXP: return run_tile_pbc<K_EXP>(a, pbc, sw, tile_cap, st);
    case K_GAUSS: return run_tile_pbc<K_GAUSS>(a, pbc, sw, tile_cap, st);
    case K_FASTGAUSS: return run_tile_pbc<K_FASTGAUSS>(a, pbc, sw, tile_cap, st);
    case K_SMAP: return run_tile_pbc<K_SMAP>(a, pbc, sw, tile_cap, st);
    case K_CUBIC: return run_tile_pbc<K_CUBIC>(a, pbc, sw, tile_cap, st);
    case K_TANH: return run_tile_pbc<K_TANH>(a, pbc, sw, tile_cap, st);
    case K_COS: return run_tile_pbc<K_COS>(a, pbc, sw, tile_cap, st);
    case K_NATIVEQ: return run_tile_pbc<K_NATIVEQ>(a, pbc, sw, tile_cap, st);
    default: return -1;
  }
}

}  // namespace b200
