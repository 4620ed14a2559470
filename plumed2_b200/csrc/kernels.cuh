// Launch wrappers of the hand-written sm_100a kernels (kernels_build.cu, kernels_sweep.cu) and the device
// helpers both translation units share.
#pragma once
#include "device_types.cuh"

namespace b200 {

// ---- 27-cell stencil, LinkCells::addRequiredCells / min_cell / max_cell (LinkCells.cpp:183-239)
// Per axis the unwrapped range is [lo,hi); ranges shrink for grids thinner than 3 cells so that a cell is
// never visited twice, and clamp to [0,n) without PBC.
__device__ __forceinline__ void stencil_bounds(const DevGrid& g, const int c[3], int lo[3], int hi[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int n = g.n[k];
    int mn = c[k] + ((n < 2) ? 0 : -1);
    int mx = c[k] + ((n < 3 && g.stencil_pbc) ? 1 : 2);
    if (!g.stencil_pbc) {
      mn = max(mn, 0);
      mx = min(mx, n);
    }
    lo[k] = mn;
    hi[k] = mx;
  }
}
__device__ __forceinline__ int wrap_cell(int m, int n) { return (m < 0) ? n - 1 : m % n; }  // LINKC_PBC

// ---- build
void launch_bbox(const double* pos, unsigned n, double* out6, unsigned long long* scratch6, cudaStream_t st);
void launch_sort(const double* pos, unsigned n, unsigned n_a, int ngroups, const DevGrid& g, uint32_t* cell_of_slot,
                 uint32_t* ccount, uint32_t* cstart, uint32_t* cursor, uint32_t* tmp, uint32_t* perm, uint32_t* scell,
                 cudaStream_t st);
void launch_identity(unsigned n, uint32_t* perm, uint32_t* scell, cudaStream_t st);
void launch_gather(const double* pos, const uint32_t* perm, const uint32_t* abs_index, unsigned n, SPos* spos,
                   cudaStream_t st);
void launch_nl_rows(bool fill, const SPos* spos, const uint32_t* scell, const uint32_t* cstart, const uint32_t* ccount,
                    const DevGrid& g, const DevPbc& pbc, double cutoff2, unsigned n_a, int two_groups, unsigned row_begin,
                    unsigned row_end, uint32_t* row_count, const unsigned long long* row_start, uint32_t* nbr,
                    cudaStream_t st);
void launch_pair_mask(const double* pos, unsigned n_a, const DevPbc& pbc, double cutoff2, uint8_t* active, cudaStream_t st);
void launch_scan_rows(const uint32_t* row_count, unsigned rows, unsigned long long* bsum, unsigned long long* row_start,
                      unsigned long long* grand_total, cudaStream_t st);

// ---- sweep
constexpr int kPartialStride = 8;  // value, vxx, vxy, vxz, vyy, vyz, vzz, (pad)

struct SweepArgs {
  const SPos* spos;
  unsigned n_a;            // atoms of group A (sorted rows [0,n_a) are A rows)
  int two_groups;          // TwoList
  int check_abs;           // skip pairs whose absolute indices coincide (only when the groups overlap)
  unsigned row_begin, row_end;  // this rank's rows (sorted indices)
  // CSR list (classic NL)
  const unsigned long long* row_start;
  const uint32_t* row_count;
  const uint32_t* nbr;
  // implicit ranges (no NL / NLISTCELLS)
  const uint32_t* scell;
  const uint32_t* cstart;
  const uint32_t* ccount;
  DevGrid grid;
  // outputs
  double* sderiv;          // 3 doubles per sorted row (only rows of this rank are written)
  double* partials;        // kPartialStride doubles per block
  unsigned long long* evals;  // pair evaluations executed (both directions)
};

// returns the number of blocks launched (= number of partial records), or -1 for an unsupported switch
int launch_sweep_list(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st);
int launch_sweep_cells(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st);
// PAIR style: one thread per pair (k, k+n_a); writes derivatives straight into out (slot order)
int launch_sweep_pairs(const double* pos, const uint32_t* abs_index, const uint8_t* active, unsigned n_a,
                       unsigned pair_begin, unsigned pair_end, const DevPbc& pbc, const DevSwitch& sw, double* out,
                       double* partials, unsigned long long* evals, cudaStream_t st);

// sum the per-block partials in a fixed order and write virial (9) + value behind the 3n derivatives;
// weight = 0.5 when every pair was visited from both sides (SingleList), 1 otherwise
void launch_finalize(const double* partials, int nblocks, double weight, double* out_tail /*[10]*/, cudaStream_t st);
// out[3*slot+c] = sderiv[3*k+c] for all sorted rows k in [0,n)
void launch_unsort_derivs(const double* sderiv, const SPos* spos, unsigned n, double* out, cudaStream_t st);

// ---- util
double measure_dfma_tflops(cudaStream_t st, int sm_count, int reps);

}  // namespace b200
