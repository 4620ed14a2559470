// Launch wrappers of the hand-written sm_100a kernels (kernels_build.cu, kernels_sweep.cu) and the device
// helpers both translation units share.
#pragma once
#include "device_types.cuh"

namespace b200 {

// ---- 27-cell stencil, LinkCells::addRequiredCells / min_cell / max_cell (LinkCells.cpp:183-239)
// Per axis the unwrapped range is [lo,hi); ranges shrink for grids thinner than 3 cells so that a cell is
// never visited twice, and clamp to [0,n) without PBC.
__host__ __device__ __forceinline__ void stencil_bounds(const DevGrid& g, const int c[3], int lo[3], int hi[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int n = g.n[k];
    int mn, mx;
    if (g.radius == 1) {
      mn = c[k] + ((n < 2) ? 0 : -1);
      mx = c[k] + ((n < 3 && g.stencil_pbc) ? 1 : 2);
    } else {  // only built for grids with n >= 2*radius+1 cells per periodic direction
      mn = c[k] - g.radius;
      mx = c[k] + g.radius + 1;
    }
    if (!g.stencil_pbc) {
      mn = max(mn, 0);
      mx = min(mx, n);
    }
    lo[k] = mn;
    hi[k] = mx;
  }
}
// LINKC_PBC (n<0 ? num-1 : n%num).  Stencil offsets never exceed one box length (|offset| <= radius <= n),
// so the wrap is one conditional add/subtract -- an integer modulo here costs more than the distance test.
__host__ __device__ __forceinline__ int wrap_cell(int m, int n) { return (m < 0) ? m + n : ((m >= n) ? m - n : m); }
// how many box lengths the unwrapped cell index m lies outside [0,n): -1, 0 or +1 for our stencils
__host__ __device__ __forceinline__ int wrap_count(int m, int n) { return (m < 0) ? -1 : ((m >= n) ? 1 : 0); }

// Visit the stencil of cell c as CONTIGUOUS sorted ranges of the partner group: cells that are neighbours
// along x are neighbours in memory (cell id = x + y*n0 + z*n0*n1, atoms sorted by cell), so a run of x cells
// is one range [start, start+count).  A run is split only where it wraps around the box.  f(start, count).
template <typename F>
__device__ __forceinline__ void for_each_stencil_range(const DevGrid& g, const int c[3], unsigned group_off,
                                                       const uint32_t* __restrict__ cstart,
                                                       const uint32_t* __restrict__ ccount, F&& f) {
  int lo[3], hi[3];
  stencil_bounds(g, c, lo, hi);
  for (int ny = lo[1]; ny < hi[1]; ++ny) {
    const int yv = wrap_cell(ny, g.n[1]) * g.n[0];
    for (int nz = lo[2]; nz < hi[2]; ++nz) {
      const unsigned base = group_off + (unsigned)(yv + wrap_cell(nz, g.n[2]) * g.n[0] * g.n[1]);
      int x = lo[0];
      while (x < hi[0]) {
        const int xw = wrap_cell(x, g.n[0]);
        const int run = min(hi[0] - x, g.n[0] - xw);
        const unsigned first = base + (unsigned)xw, last = first + (unsigned)run - 1u;
        const uint32_t s = cstart[first];
        // (wx,wy,wz): the periodic image these cells are seen through (0 without wrapping)
        f(s, cstart[last] + ccount[last] - s, wrap_count(x, g.n[0]), wrap_count(ny, g.n[1]), wrap_count(nz, g.n[2]));
        x += run;
      }
    }
  }
}

__host__ __device__ __forceinline__ void cell_coords(const DevGrid& g, int cell, int c[3]) {  // LinkCells::findMyCell(idx) :294-304
  c[2] = cell / (g.n[0] * g.n[1]);
  const int rem = cell - c[2] * g.n[0] * g.n[1];
  c[1] = rem / g.n[0];
  c[0] = rem - c[1] * g.n[0];
}

__device__ __forceinline__ uint32_t warp_exclusive_scan(uint32_t v, unsigned lane, uint32_t& total) {
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (unsigned)o) x += y;
  }
  total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}

// ---- where things live when the atoms are spread over several GPUs of one node (NVLink peer memory)
// The caller's positions: slot s is at base[s / chunk] + 3 * (s % chunk).  One rank: base[0] = the whole array and
// chunk = 0xffffffff.  Several ranks (b200coord_calculate_distributed): base[r] = rank r's slice buffer, mapped into
// this process, and a gather PULLS what it needs from the owners.
struct PosSrc {
  const double* base[8];
  unsigned chunk;
};
__device__ __forceinline__ const double* pos_at(const PosSrc& p, uint32_t slot) {
  const unsigned o = slot / p.chunk;
  return p.base[o] + 3 * (size_t)(slot - o * p.chunk);
}
static inline PosSrc pos_src_local(const double* pos) {
  PosSrc p;
  for (int i = 0; i < 8; ++i) p.base[i] = pos;
  p.chunk = 0xffffffffu;
  return p;
}
// Derivative rows (3 doubles per sorted atom, global sorted index): row k is in base[k / chunk].
struct RowSrc {
  const double* base[8];
  unsigned chunk;
};
// up to six intervals of sorted indices: the atoms a rank needs (its own rows and every possible partner of them)
struct IdxRanges {
  unsigned n;
  unsigned lo[6], len[6];
  unsigned total;
};

// ---- build
void launch_bbox(const double* pos, unsigned n, double* out6, unsigned long long* scratch6, cudaStream_t st);
void launch_sort(const double* pos, unsigned n, unsigned n_a, int ngroups, const DevGrid& g, uint32_t* cell_of_slot,
                 uint32_t* ccount, uint32_t* cstart, uint32_t* cursor, uint32_t* tmp, uint32_t* perm, uint32_t* scell,
                 unsigned long long* scan_tmp /* >= ngroups*ncell/1024+2 entries, or null */, cudaStream_t st);
void launch_identity(unsigned n, uint32_t* perm, uint32_t* scell, cudaStream_t st);
void launch_gather(const double* pos, const uint32_t* perm, const uint32_t* abs_index, unsigned n, SPos* spos,
                   cudaStream_t st);
// sq[k] = charges[perm[k]] (DHENERGY; after every re-sort)
void launch_gather_charges(const double* charges, const uint32_t* perm, unsigned n, double* sq, cudaStream_t st);
void launch_gather_types(const uint32_t* types, const uint32_t* perm, unsigned n, uint32_t* stype, cudaStream_t st);
// the same, plus displacement tracking: track 1 = store the build-time positions in bpos (sorted order),
// track 2 = atomicMax the largest squared displacement since then into *disp2 (bits of a double)
void launch_gather_track(int track, const double* pos, const uint32_t* perm, const uint32_t* abs_index, unsigned n,
                         SPos* spos, double* bpos, const DevPbc& pbc, unsigned long long* disp2, cudaStream_t st);
void launch_nl_rows(bool fill, const SPos* spos, const uint32_t* scell, const uint32_t* cstart, const uint32_t* ccount,
                    const DevGrid& g, const DevPbc& pbc, double cutoff2, unsigned n_a, int two_groups, unsigned row_begin,
                    unsigned row_end, uint32_t* row_count, const unsigned long long* row_start, uint32_t* nbr,
                    cudaStream_t st);
// image mode (sweep_img.cuh): see k_sort_init / k_gather_u / k_pack_meta in kernels_build.cu
void launch_sort_init(const double* pos, const uint32_t* perm, const uint32_t* abs_index, unsigned n, const DevGrid& g,
                      const DevPbc& box, bool use_wrapped, float4* lpos, double* wpos, double* braw, SPos* spos, double* ubuild,
                      cudaStream_t st);
void launch_gather_u(bool build, const PosSrc& pos, const uint32_t* perm, const uint32_t* abs_index, const IdxRanges& rng,
                     const double* wpos, const double* braw, const DevPbc& pbc, SPos* spos, double* ubuild, float4* lpos,
                     unsigned long long* disp2, cudaStream_t st);
// inv[perm[k]] = k
void launch_invert_perm(const uint32_t* perm, unsigned n, uint32_t* inv, cudaStream_t st);
// the list as (i0, i1) index pairs on the device (b200coord_nl_pairs_device): count pass, scan, fill pass
void launch_export_pairs(bool fill, unsigned rows, unsigned row_begin, unsigned n_a, int two_groups, const uint32_t* perm,
                         const unsigned long long* row_start, const uint32_t* row_count, const uint32_t* far_off,
                         const uint32_t* far_cnt, const uint32_t* nbr, uint32_t idx_mask, uint32_t* emit_count,
                         const unsigned long long* emit_start, unsigned* pairs, unsigned long long capacity, cudaStream_t st);
void launch_pack_meta(unsigned rows, const unsigned long long* row_start, const uint32_t* row_count, const uint32_t* far_off,
                      const uint32_t* far_cnt, uint4* meta, unsigned long long* listed, cudaStream_t st);
// ---- super-list: a list with cutoff NL_CUTOFF + delta whose rows are the candidate sets of the following rebuilds
// (k_nl_filter) while 2 * max displacement since its build stays below delta.  An entry = sorted index | image << 26.
constexpr uint32_t kSuperIndexMask = 0x03ffffffu;
// image code = (wx+1) | (wy+1)<<2 | (wz+1)<<4, stored XOR the code of the home cell (21): an entry that needs no
// shift has zero top bits, so "does any entry of this trip need a shift" is one unsigned compare
constexpr uint32_t kImageCentre = 21u;
__host__ __device__ __forceinline__ uint32_t super_image(int wx, int wy, int wz) {
  return ((((uint32_t)(wx + 1) | ((uint32_t)(wy + 1) << 2) | ((uint32_t)(wz + 1) << 4))) ^ kImageCentre) << 26;
}
void launch_nl_filter(int mode /*0 count, 1 fill, 2 capped single pass*/, bool images /*entries keep their image bits*/,
                      const double* pos /*caller's positions (exact band test)*/, const uint32_t* perm, const float4* lpos,
                      const unsigned long long* srow_start, const uint32_t* srow_count, const uint32_t* snbr,
                      const DevPbc* pbc_g /*box parameters in global memory*/, const DevPbc& box, double cutoff2, double band_rel, unsigned n_a, int two_groups, unsigned row_begin,
                      unsigned row_end, uint32_t* row_count, unsigned long long* row_start, uint32_t* nbr, unsigned row_cap,
                      unsigned* cap_info, float far2, uint32_t* row_far_off, uint32_t* row_far_cnt, unsigned max_srow /*longest super-list row; 0: per-row kernel*/, int flat_minb /*2 or 3 blocks per SM*/,
                      cudaStream_t st);
// FP32 candidate search on the local copy; the thin band around the cutoff falls back to the exact FP64 test
void launch_nl_rows_f32(int mode /*0 count, 1 fill, 2 capped single pass*/, bool super /*super-list rows*/,
                        bool images /*entries carry the periodic image in their top 6 bits*/, const double* pos,
                        const uint32_t* perm, const float4* lpos, const uint32_t* scell, const uint32_t* cstart, const uint32_t* ccount,
                        const DevGrid& g, const DevPbc* pbc_g /*global memory*/, const DevPbc& box, double cutoff2, double band_rel,
                        unsigned n_a, int two_groups, unsigned row_begin, unsigned row_end, uint32_t* row_count,
                        unsigned long long* row_start, uint32_t* nbr, unsigned row_cap,
                        unsigned* cap_info /*[0] max row, [1] overflow*/, float far2 /*near/far split, r^2*/,
                        uint32_t* row_far_off, uint32_t* row_far_cnt, cudaStream_t st);
void launch_pair_mask(const double* pos, unsigned n_a, const DevPbc& pbc, double cutoff2, uint8_t* active, cudaStream_t st);
void launch_scan_rows(const uint32_t* row_count, unsigned rows, unsigned padq /*3: list rows start on 16-byte boundaries*/,
                      unsigned long long* bsum, unsigned long long* row_start, unsigned long long* grand_total,
                      cudaStream_t st);

// ---- sweep
constexpr int kPartialStride = 12;  // value, then the 3x3 sum c[a][b] (virial = -weight * c), (pad)

struct SweepArgs {
  const SPos* spos;
  const double* sq;        // charges in sorted order (DHENERGY), else null
  const uint32_t* stype;   // interaction types in sorted order, the ntypes x ntypes table (GHBFIX), else null
  const double* etas;
  unsigned ntypes;
  unsigned n_a;            // atoms of group A (sorted rows [0,n_a) are A rows)
  int two_groups;          // TwoList
  int check_abs;           // skip pairs whose absolute indices coincide (only when the groups overlap)
  unsigned row_begin, row_end;  // this rank's rows (sorted indices)
  // CSR list (classic NL)
  const unsigned long long* row_start;
  const uint32_t* row_count;
  const uint32_t* nbr;
  uint32_t idx_mask;       // sorted index of an entry = entry & idx_mask (image-mode lists keep the image above bit 26)
  const uint4* row_meta;   // image mode: {row start / 4, near count, far offset, far count} per row (k_pack_meta)
  PosSrc pos;              // the caller's positions (slot order): exact boundary patch
  double img_disp2_max;    // image mode is valid while the squared displacement since the rebuild is below this
  unsigned rows_per_block; // image mode: both kernels use this block shape (0: the kernel picks)
  const uint32_t* inv;     // slot -> sorted index (scatter_b in the cell-tile sweep)
  int scatter_b;           // image mode, two groups with few GROUPA atoms on one rank: the GROUPA rows add +dd to their
                           // partners' derivative rows (RED.ADD.F64) and the GROUPB rows are not swept at all
  // every row is stored in two parts: [row_start, +row_count) holds the partners that were inside D_MAX (+ a skin)
  // when the list was built, [row_start + row_far_off, +row_far_cnt) the rest (filled from the end of the row's
  // allocation).  A trip of the far part whose 32 pairs are all beyond D_MAX contributes exactly zero and stops
  // after the distance test.
  const uint32_t* row_far_off;
  const uint32_t* row_far_cnt;
  double far_skip2;        // r^2 above which a pair is certainly beyond D_MAX and its boundary band (inf: never)
  // Verlet skin: the far parts are not visited at all while the largest displacement since the rebuild is below half
  // the skin the far partners had beyond D_MAX (k_gather_sorted tracks it)
  const unsigned long long* disp2_bits;  // largest squared displacement since the rebuild (bits of a double)
  double far_disp2_max;                  // visit the far parts unless disp2 < this
  int force_far;                         // box changed since the rebuild: always visit them
  int f32;                               // B200COORD_FP32: FP32 pair arithmetic (kernels_sweep_f32.cu)
  // implicit ranges (no NL / NLISTCELLS)
  const uint32_t* scell;
  const uint32_t* cstart;
  const uint32_t* ccount;
  DevGrid grid;
  // outputs
  double* sderiv;          // 3 doubles per sorted row (only rows of this rank are written)
  double* partials;        // kPartialStride doubles per block
  unsigned long long* evals;  // list entries of the rows swept (both directions)
  unsigned long long* executed;  // entries actually evaluated (far parts that were skipped do not count)
  // box and switch parameters in global memory, for the out-of-line row patch (sweep_math.cuh: row_fixup_*)
  const DevPbc* pbc_g;
  const DevSwitch* sw_g;
};

// returns the number of blocks launched (= number of partial records), or -1 for an unsupported switch
int launch_sweep_list(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st);
// image mode (sweep_img.cuh): continuous coordinates + image in the list entry; box = the lattice vectors the
// images refer to.  Runs only while the displacement bound holds (device-side gate), k_sweep_list otherwise:
// launch both, exactly one of them does the step.
int launch_sweep_img(const SweepArgs& a, const DevPbc& box, const DevSwitch& sw, int variant, cudaStream_t st);
// block shape for both kernels of an image-mode step (0: rows too long for the image sweep)
unsigned sweep_img_rows_per_block(unsigned rows_a, unsigned rows_b, unsigned max_row);
int launch_sweep_cells(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st);
// PAIR style: one thread per pair (k, k+n_a); writes derivatives straight into out (slot order)
int launch_sweep_pairs(const double* pos, const double* charges /*slot order, DHENERGY*/,
                       const uint32_t* types /*slot order*/, const double* etas, unsigned ntypes /*GHBFIX*/, const uint32_t* abs_index, const uint8_t* active, unsigned n_a,
                       unsigned pair_begin, unsigned pair_end, const DevPbc& pbc, const DevSwitch& sw, double* out,
                       double* partials, unsigned long long* evals, cudaStream_t st);

// cell-tile sweep (sweep_tile.cuh): NLISTCELLS / no list, FP64, all kinds but DHENERGY / GHBFIX
struct TileWork;
bool sweep_tile_supports(int sw_type);
unsigned tile_rows_per_item(unsigned rows);
unsigned tile_items_bound(unsigned rows, unsigned nseg, unsigned rows_per_item);
void launch_tile_work(unsigned nseg, const uint32_t* cstart, const uint32_t* ccount, unsigned row_begin, unsigned row_end,
                      unsigned rows_per_item, uint32_t* nitems /*nseg*/, unsigned long long* item_start /*nseg + 1*/,
                      unsigned long long* bsum, TileWork* work, cudaStream_t st);
// zero_dev / split_dev / total_dev: device words holding 0, the first GROUPB item, the item count;
// bound_a / bound_b: blocks to launch (upper bounds).  Returns the number of partial records (= bound_a), -1: unsupported
int launch_sweep_tile(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, const TileWork* work,
                      const unsigned long long* zero_dev, const unsigned long long* split_dev, const unsigned long long* total_dev,
                      unsigned bound_a, unsigned bound_b, cudaStream_t st);

// sum the per-block partials in a fixed order and write virial (9) + value behind the 3n derivatives;
// weight = 0.5 when every pair was visited from both sides (SingleList), 1 otherwise
// scratch: 16 doubles per 256 partial records
void launch_finalize(const double* partials, int nblocks, double weight, double* out_tail /*[10]*/, double* scratch,
                     cudaStream_t st);
// out[3*slot+c] = row inv[slot], component c, for the slots [slot_lo, slot_lo+slot_cnt): every rank PULLS the rows of its
// own atoms from whoever swept them (NVLink peer loads; one rank: a local gather)
void launch_coupled_gather(const double* pos_all, const uint32_t* index, unsigned n, double* pos, cudaStream_t st);
void launch_coupled_apply(const double* deriv, const uint32_t* index /*nullptr: 0..n-1*/, unsigned n, double factor,
                          double* force_all, cudaStream_t st);
void launch_unsort_pull(const RowSrc& rows, const uint32_t* inv /*slot -> sorted*/, double* out, unsigned slot_lo,
                        unsigned slot_cnt, cudaStream_t st);

// ---- util
double measure_dfma_tflops(cudaStream_t st, int sm_count, int reps);

}  // namespace b200
