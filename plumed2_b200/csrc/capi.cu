// C ABI of libb200coord.so (include/b200coord.h): context, neighbour-list schedule, rebuild and sweep
// orchestration on one CUDA stream, host<->device staging, optional NCCL combine across ranks.
// There is no CPU fallback anywhere in this file: every numeric result comes from the kernels in
// kernels_build.cu / kernels_sweep.cu, and a missing device or kernel image is an error.
#include "../../include/b200coord.h"
#include "host_setup.hpp"
#include "kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace b200;

namespace {

thread_local std::string g_last_error;

// ---- NCCL resolved lazily with dlopen so that single-GPU users need no libnccl at all
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;  // optional
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl_api() {
  static NcclApi api;
  if (api.handle || api.ok) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
#define B200_SYM(field, sym) *(void**)(&api.field) = dlsym(api.handle, sym)
  B200_SYM(GetUniqueId, "ncclGetUniqueId");
  B200_SYM(CommInitRank, "ncclCommInitRank");
  B200_SYM(CommDestroy, "ncclCommDestroy");
  B200_SYM(CommAbort, "ncclCommAbort");
  B200_SYM(AllGather, "ncclAllGather");
  B200_SYM(AllReduce, "ncclAllReduce");
  B200_SYM(GroupStart, "ncclGroupStart");
  B200_SYM(GroupEnd, "ncclGroupEnd");
  B200_SYM(GetErrorString, "ncclGetErrorString");
#undef B200_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce && api.GroupStart &&
           api.GroupEnd && api.GetErrorString;
  return api;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = n + n / 8 + 16;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct b200coord_ctx {
  b200coord_config cfg;
  b200coord_switch sw;
  DevSwitch dsw;
  unsigned n = 0, n_a = 0, n_b = 0;
  int two_groups = 0, check_abs = 0;
  unsigned long long n_self_pairs = 0;  // listed by the reference, skipped by the sweep
  std::vector<unsigned> abs_host;
  int device = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[10] = {nullptr};  // 0/1 h2d, 2/3 sweep, 4/5 build, 6/7 d2h, 8/9 user stopwatch
  bool ev_valid[4] = {false, false, false, false};
  static constexpr int kRing = 64;
  cudaEvent_t sweep_ev[2 * kRing] = {nullptr};  // per-step sweep stopwatch pairs since the last stream_mark(0)
  unsigned sweep_n = 0;
  float build_ms_acc = 0.f, build_ms_max = 0.f;
  unsigned build_n = 0;

  HostPbc hpbc;
  DevPbc dpbc;
  bool box_set = false;
  double box_cached[9] = {0};
  DevGrid grid;
  bool sorted_valid = false;  // perm/scell/cstart/ccount describe the current grouping
  bool list_valid = false;

  // schedule (CoordinationBase members firsttime / invalidateList, CoordinationBase.cpp:46-47)
  bool firsttime = true, invalidate = true;

  // rows of this rank
  unsigned row_begin = 0, row_end = 0, row_chunk = 0;
  unsigned slot_begin = 0, slot_count = 0;  // this rank's slice of the position array (distributed step)

  DevBuf<double> d_pos, d_out, d_sderiv, d_partials, d_small, d_fin;
  DevBuf<uint32_t> d_abs, d_perm, d_scell, d_cell_of_slot, d_tmp, d_ccount, d_cstart, d_cursor, d_rowcount, d_nbr;
  DevBuf<unsigned long long> d_rowstart, d_bsum, d_u64;  // d_u64: [0] grand total, [1] evals, [2..7] bbox scratch, [10] max displacement^2 since the list build (bits),
                                                          // [11] the same since the super-list build
  DevBuf<SPos> d_spos;
  DevBuf<uint8_t> d_active;
  DevBuf<unsigned char> d_params;  // [DevPbc | DevSwitch] in global memory for the out-of-line row patch
  bool params_dirty = true;
  DevBuf<float4> d_lpos;
  DevPbc dbox;           // the box itself (type/ortho flags irrelevant): lattice vectors for image shifts
  bool f32_search = false;
  double band_rel = 0.0;
  unsigned row_cap = 0;        // per-row capacity learnt from the last rebuild (0: unknown -> two-pass build)
  unsigned max_row = 0;        // longest row (near + far entries) of the current list
  unsigned super_cap = 0;      // the same capacity for the rows of the super-list
  DevBuf<double> d_q, d_sq;    // charges in slot order / sorted order (DHENERGY)
  bool have_charges = false, sq_valid = false;
  DevBuf<uint32_t> d_types, d_stype;  // interaction types in slot order / sorted order (GHBFIX)
  DevBuf<double> d_etas;
  unsigned ntypes = 0;
  bool have_types = false, stype_valid = false;
  // super-list (kernels.cuh): rows with cutoff NL_CUTOFF + super_delta, the candidate sets of later rebuilds
  DevBuf<unsigned long long> d_srowstart;
  DevBuf<uint32_t> d_srowcount, d_snbr;
  DevBuf<double> d_wpos, d_braw;  // wrapped / raw positions at the super-list build (sorted order)
  bool super_on = true;          // B200COORD_NO_SUPERLIST=1 turns it off
  bool super_off = false;        // given up at run time (box or displacements change too fast for it to pay)
  bool super_valid = false;
  int super_streak = 0;          // consecutive rebuilds that had to rebuild the super-list as well
  double super_delta = 0.0, search_cutoff = 0.0;
  unsigned long super_box_epoch = 0;
  unsigned long long super_builds = 0, filter_rebuilds = 0;
  DevBuf<double> d_bpos;       // positions the list was built from (sorted order), for the displacement bound
                               // (continuous coordinates u in u_mode)
  // u_mode (FP32-search lists): the sorted records hold continuous coordinates u = wrapped position at the sort +
  // minimum-image displacement since (kernels_build.cu: k_sort_init / k_gather_u); img_list: the list entries carry
  // the periodic image their partner was found through, and k_sweep_img (sweep_img.cuh) may take the step
  bool u_mode = false, img_list = false;
  unsigned long sort_box_epoch = 0;   // box at the last re-sort (wpos / braw / images refer to it)
  DevBuf<uint4> d_meta;               // per row {start / 4, near count, far offset, far count}
  int img_variant = 2;                // resident blocks per SM the image sweep is compiled for (B200COORD_IMG_VARIANT)
  bool img_on = true;                 // B200COORD_NO_IMG_SWEEP=1: always the general kernel
  bool scatter_on = true;             // B200COORD_NO_SCATTER=1: sweep GROUPB rows even when they are short
  DevBuf<uint32_t> d_rowfar;   // [0,rows) offset of the far part inside the row's allocation, [rows, 2 rows) its length
  DevBuf<unsigned> d_capinfo;  // [0] max row count seen, [1] overflow flag
  unsigned* h_capinfo = nullptr;
  unsigned long long nbr_total = 0;
  unsigned far_rows = 0;     // rows of d_rowfar (offsets, then lengths)
  bool far_split = true;     // B200COORD_NO_FAR_SPLIT=1 keeps rows in one part
  double far_skin = 0.0;     // far partners were beyond D_MAX + far_skin when the list was built (0: no far parts)
  unsigned long box_epoch = 0, build_box_epoch = 0;  // set_box calls that changed the box / value at the last rebuild
  int sweep_blocks = 0;

  double* h_small = nullptr;  // pinned: [0..9] tail, [10..15] bbox
  unsigned long long* h_u64 = nullptr;  // pinned: [0] grand total, [1] evals

  // opt-in (B200COORD_PIN_HOST=1): page-lock the caller's position/derivative arrays so that the per-step
  // copies are direct DMA instead of staged pageable copies
  bool pin_host = false;
  struct Pinned { const void* p = nullptr; size_t bytes = 0; } pinned[2];

  ncclComm_t comm = nullptr;
  // fused exchange over peer memory: two row buffers (step parity) per rank, mapped into every process
  bool peer_mode = false;
  bool peer_ipc = false;              // peer pointers came from cudaIpcOpenMemHandle (other processes), not from this process
  DevBuf<double> d_sderiv_b;          // second parity buffer (d_sderiv is the first)
  double* peer_rows[2][8] = {{nullptr}};  // [parity][rank], own entries point at the local buffers
  unsigned parity = 0;
  // positions over peer memory: every rank uploads only its slot slice into one of two slice buffers, and the gather
  // of a step that keeps its list pulls the positions it needs from the owners (kernels.cuh: PosSrc)
  DevBuf<double> d_pslice[2];
  double* peer_pos[2][8] = {{nullptr}};
  unsigned pos_parity = 0;
  DevBuf<uint32_t> d_inv;             // slot -> sorted index (inverse of d_perm)
  // cell-tile sweep (NLISTCELLS / no list): work list of (cell, row chunk) items, rebuilt with the sort
  DevBuf<unsigned char> d_tilework;
  DevBuf<uint32_t> d_tilecnt;
  DevBuf<unsigned long long> d_tilestart;
  unsigned tile_bound_a = 0, tile_bound_b = 0, tile_nseg = 0;
  bool tile_ok = false;               // the current sort has a work list
  unsigned super_max_row = 0;         // longest row of the current super-list
  bool filter_flat = true;            // B200COORD_FILTER_FLAT=0: the per-row filter kernel
  int filter_minb = 2;                // B200COORD_FILTER_MINB=3: the 80-register build of the flat filter (3 blocks per SM)
  bool tile_on = true;                // B200COORD_NO_TILE_SWEEP=1: the warp-per-row kernel instead
  bool in_process = false;            // one of several contexts of ONE process (b200coord_group_*), see collective_done
  bool comm_aborted = false;          // the communicator was aborted after a failure of a group member: nothing to destroy
  int test_fail_at = -1;              // B200COORD_TEST_FAIL=<rank>:<call>: the distributed call that fails on purpose (tests)
  int dist_calls = 0;
  // frames known in advance (b200coord_submit / _collect): two steps in flight, copies on their own streams
  struct Pending {
    bool busy = false;
    double* value = nullptr;
    double* virial = nullptr;
  } pend[2];
  cudaStream_t st_up = nullptr, st_down = nullptr;
  cudaEvent_t q_up[2] = {nullptr, nullptr}, q_done[2] = {nullptr, nullptr}, q_down[2] = {nullptr, nullptr};
  DevBuf<double> d_posq[2], d_outq[2];
  double* h_tailq = nullptr;          // pinned: 2 x [virial 9 | value]
  unsigned q_next = 0;
  // coupling with an engine whose arrays live on the device: where the action's atoms sit in the engine's arrays
  DevBuf<uint32_t> d_cidx;
  bool cidx_set = false;              // false: atom i of the action is atom i of the engine
  bool coupled_fresh = false;         // d_out holds the derivatives of the last b200coord_calculate_coupled
  IdxRanges needed;                   // sorted indices this rank's rows can touch (its rows + every possible partner)
  unsigned* h_idx = nullptr;          // pinned scratch for compute_needed
  b200coord_stats stats;
  std::string err;
};

namespace {

// B200COORD_TRACE=1: one line on stderr at the main stations of a step (which rank is where when something waits)
bool trace_on() {
  static const bool on = [] { const char* e = std::getenv("B200COORD_TRACE"); return e && std::atoi(e) != 0; }();
  return on;
}
#define B200_TRACE(c, ...)                                                              \
  do {                                                                                  \
    if (trace_on()) {                                                                   \
      std::fprintf(stderr, "[b200coord r%d/%d dev%d] ", (c)->cfg.rank, (c)->cfg.nranks, (c)->device); \
      std::fprintf(stderr, __VA_ARGS__);                                                \
      std::fprintf(stderr, "\n");                                                       \
      std::fflush(stderr);                                                              \
    }                                                                                   \
  } while (0)

int fail(b200coord_ctx* c, int code, const std::string& msg) {
  if (c) B200_TRACE(c, "FAIL %d: %s", code, msg.c_str());
  // a context of an in-process group that fails leaves its peers waiting in their next collective: say why
  if (c && c->in_process && !trace_on())
    std::fprintf(stderr, "b200coord: device %d (rank %d of %d in this process) failed: %s\n", c->device, c->cfg.rank,
                 c->cfg.nranks, msg.c_str());
  if (c) c->err = msg;
  g_last_error = msg;
  return code;
}

#define CU(c, expr)                                                                                     \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return fail(c, B200COORD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));          \
  } while (0)

#define CU_LAST(c, what)                                                                                \
  do {                                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                                               \
    if (e__ != cudaSuccess) return fail(c, B200COORD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e__)); \
  } while (0)

void to_dev_switch(const b200coord_switch& s, DevSwitch& d) {
  d.type = s.type;
  d.d0 = s.d0; d.dmax = s.dmax; d.dmax_2 = s.dmax_2; d.invr0 = s.invr0; d.invr0_2 = s.invr0_2;
  d.stretch = s.stretch; d.shift = s.shift;
  d.nn = s.nn; d.mm = s.mm; d.preRes = s.preRes; d.preDfunc = s.preDfunc; d.preSecDev = s.preSecDev;
  d.nnf = s.nnf; d.mmf = s.mmf; d.preDfuncF = s.preDfuncF; d.preSecDevF = s.preSecDevF;
  d.a = s.a; d.b = s.b; d.c = s.c; d.d = s.d; d.beta = s.beta; d.lambda = s.lambda; d.ref = s.ref;
  d.pre_df = 2.0 * s.invr0_2 * s.stretch;
  d.d0_2 = s.d0 * s.d0;
  d.band_dmax = (s.dmax < 1.0e150) ? 1e-10 * s.dmax_2 : -1.0;
  d.band_d0 = (s.d0 > 0.0) ? 1e-10 * d.d0_2 : -1.0;
  if (s.type == B200COORD_PAIR_GHBFIX || s.type == B200COORD_PAIR_DHENERGY) {
    d.band_dmax = -1.0;  // GHBFIX is C1 at D_0, at the joint and at D_MAX; DHENERGY has no cutoff: nothing to patch
    d.band_d0 = -1.0;
  }
  // rationalfixN evaluates y^(N/2-1): keep N/2 in nnf for the device (the host struct leaves the default there)
  switch (s.type) {
    case B200COORD_SW_RATIONALFIX12: d.nnf = 6; break;
    case B200COORD_SW_RATIONALFIX10: d.nnf = 5; break;
    case B200COORD_SW_RATIONALFIX8: d.nnf = 4; break;
    case B200COORD_SW_RATIONALFIX6: d.nnf = 3; break;
    case B200COORD_SW_RATIONALFIX4: d.nnf = 2; break;
    case B200COORD_SW_RATIONALFIX2: d.nnf = 1; break;
    default: break;
  }
  d.fix_df = -(double)d.nnf * d.pre_df;
  d.dmax_2_f64 = d.dmax_2;
  d.d0_2_f64 = d.d0_2;
}

void to_dev_pbc(const HostPbc& h, bool use_pbc, DevPbc& d) {
  std::memset(&d, 0, sizeof(d));
  d.type = use_pbc ? h.type : 0;
  std::memcpy(d.box, h.box, sizeof(d.box));
  std::memcpy(d.inv_box, h.inv_box, sizeof(d.inv_box));
  std::memcpy(d.reduced, h.reduced, sizeof(d.reduced));
  std::memcpy(d.inv_reduced, h.inv_reduced, sizeof(d.inv_reduced));
  std::memcpy(d.nshift, h.nshift, sizeof(d.nshift));
  std::memcpy(d.shifts, h.shifts, sizeof(d.shifts));
}

bool box_is_zero(const double b[9]) {
  for (int i = 0; i < 9; ++i)
    if (b[i] != 0.0) return false;
  return true;
}

void set_grid_from_box(DevGrid& g, const double inv_box[9], const unsigned nc[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) g.inv_box_t[3 * i + j] = inv_box[3 * j + i];
  for (int k = 0; k < 3; ++k) g.n[k] = (int)nc[k];
  g.ncell = g.n[0] * g.n[1] * g.n[2];
}

// cell grid for the coming rebuild.  NLISTCELLS: exactly LinkCells::setupCells (LinkCells.cpp:49-122).
// NLIST: our own grid, cell width >= cutoff*(1+1e-6), only used to FIND candidates (the kept set is decided
// by the exact distance test, so the grid cannot change it).
int setup_grid(b200coord_ctx* c, const double* d_pos) {
  DevGrid& g = c->grid;
  std::memset(&g, 0, sizeof(g));
  const bool cells_mode = (c->cfg.nl_mode == B200COORD_NL_CELLS);
  const double cut = cells_mode ? c->cfg.nl_cutoff : c->search_cutoff * (1.0 + 1e-6);
  bool use_bbox;
  if (cells_mode) {
    use_bbox = box_is_zero(c->hpbc.box);
    if (!use_bbox && c->hpbc.type == 0)
      return fail(c, B200COORD_ERR_INVALID, "Cell lists cannot be built when passing a box with null volume");
    g.stencil_pbc = c->cfg.pbc ? 1 : 0;
  } else {
    use_bbox = !(c->cfg.pbc && c->hpbc.type != 0);
    g.stencil_pbc = use_bbox ? 0 : 1;
  }
  g.bbox = use_bbox ? 1 : 0;
  g.radius = 1;
  c->f32_search = false;
  unsigned nc[3];
  double extent_max = 0.0;
  HostPbc bb;
  if (use_bbox) {
    launch_bbox(d_pos, c->n, c->d_small.p + 16, c->d_u64.p + 2, c->st);
    c->stats.kernel_launches += 2;
    CU(c, cudaMemcpyAsync(c->h_small + 10, c->d_small.p + 16, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));
  }
  // NLIST only: our own search grid may use half-width cells with a 5x5x5 stencil (1.7x fewer candidates)
  // when every periodic direction still has >= 5 of them; NLISTCELLS must keep the reference's grid.
  for (int refine = (cells_mode ? 1 : 2); refine >= 1; --refine) {
    const double w = cut / refine;
    if (use_bbox) {
      double box[9] = {0};
      for (int k = 0; k < 3; ++k) {
        const double mn = c->h_small[10 + k], mx = c->h_small[13 + k];
        box[4 * k] = (w < std::sqrt(1.79769313486231570e308)) ? w * (1 + std::ceil((mx - mn) / w)) : (mx - mn + 1);
        g.origin[k] = (mn + mx) / 2;
        extent_max = std::max(extent_max, box[4 * k]);
      }
      setup_pbc(box, bb);
      cell_grid(bb.inv_box, w, nc);
      set_grid_from_box(g, bb.inv_box, nc);
    } else {
      cell_grid(c->hpbc.inv_box, w, nc);
      set_grid_from_box(g, c->hpbc.inv_box, nc);
      for (int k = 0; k < 9; ++k) extent_max = std::max(extent_max, std::fabs(c->hpbc.box[k]));
    }
    g.radius = refine;
    const unsigned need = 2u * refine + 1u;
    const bool wide = use_bbox || (nc[0] >= need && nc[1] >= need && nc[2] >= need);
    if (refine == 1 || (wide && (unsigned long long)nc[0] * nc[1] * nc[2] <= 64ull * c->n + 4096ull)) {
      // FP32 search needs the stencil image to be the minimum image of every pair within the cutoff
      c->f32_search = !cells_mode && wide;
      break;
    }
  }
  // rounding band of the FP32 test, relative to cutoff^2: coordinates up to ~extent/2 carry 2^-24 relative error
  c->band_rel = 64.0 * 5.96e-8 * (extent_max / c->cfg.nl_cutoff + 4.0);
  if (c->band_rel > 0.05) c->f32_search = false;
  if (use_bbox) {
    std::memset(&c->dbox, 0, sizeof(c->dbox));
  } else {
    std::memset(&c->dbox, 0, sizeof(c->dbox));
    std::memcpy(c->dbox.box, c->hpbc.box, sizeof(c->dbox.box));
  }
  if ((unsigned long long)nc[0] * nc[1] * nc[2] > 400000000ull)
    return fail(c, B200COORD_ERR_INVALID, "cell grid too large for the given NL_CUTOFF and box");
  for (int k = 0; k < 3; ++k) c->stats.ncells[k] = nc[k];
  return B200COORD_OK;
}

int ensure_cell_arrays(b200coord_ctx* c) {
  const size_t m = (size_t)(c->two_groups ? 2 : 1) * (size_t)c->grid.ncell;
  CU(c, c->d_ccount.reserve(m));
  CU(c, c->d_cstart.reserve(m));
  CU(c, c->d_cursor.reserve(m));
  CU(c, c->d_bsum.reserve(std::max<size_t>(m, (size_t)c->n) / 1024 + 4));
  return B200COORD_OK;
}

void needed_all(b200coord_ctx* c);
int collective_done(b200coord_ctx* c);
int build_tile_work(b200coord_ctx* c);

// no neighbour list: one "cell" per group holding every atom in slot order (NeighborList.cpp:133-140 order)
int setup_all_pairs(b200coord_ctx* c) {
  DevGrid& g = c->grid;
  std::memset(&g, 0, sizeof(g));
  g.n[0] = g.n[1] = g.n[2] = 1;
  g.ncell = 1;
  g.stencil_pbc = 0;
  int rc = ensure_cell_arrays(c);
  if (rc) return rc;
  const uint32_t starts[2] = {0u, c->n_a}, counts[2] = {c->n_a, c->n_b};
  const size_t m = c->two_groups ? 2 : 1;
  CU(c, cudaMemcpyAsync(c->d_cstart.p, starts, m * sizeof(uint32_t), cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->d_ccount.p, counts, m * sizeof(uint32_t), cudaMemcpyHostToDevice, c->st));
  launch_identity(c->n, c->d_perm.p, c->d_scell.p, c->st);
  launch_identity(c->n, c->d_inv.p, c->d_scell.p, c->st);
  needed_all(c);
  c->stats.kernel_launches += 2;
  CU(c, cudaStreamSynchronize(c->st));  // starts/counts live on this stack frame
  CU_LAST(c, "identity perm");
  c->sorted_valid = true;
  rc = build_tile_work(c);
  if (rc) return rc;
  for (int k = 0; k < 3; ++k) c->stats.ncells[k] = 1;
  return B200COORD_OK;
}

void needed_all(b200coord_ctx* c) {
  std::memset(&c->needed, 0, sizeof(c->needed));
  c->needed.n = 1;
  c->needed.lo[0] = 0;
  c->needed.len[0] = c->n;
  c->needed.total = c->n;
}

// After a re-sort on a sharded context: the sorted atoms this rank can touch until the next re-sort, as intervals of
// sorted indices.  Atoms are sorted by (group, cell) with cell = x + y n0 + z n0 n1, so the rank's rows cover a range
// of z layers; every partner candidate lies within `radius` layers of its row's layer (the stencil), in either group.
// Conservative (whole layers) and cheap: a handful of 4-byte reads per re-sort.
int compute_needed(b200coord_ctx* c) {
  needed_all(c);
  if (c->cfg.nranks <= 1 || c->row_end <= c->row_begin) return B200COORD_OK;
  const DevGrid& g = c->grid;
  const unsigned layer = (unsigned)g.n[0] * (unsigned)g.n[1];
  const int n2 = g.n[2];
  if (n2 < 4 || layer == 0) return B200COORD_OK;
  // cells of the first and last row of each part (GROUPA rows / GROUPB rows) of my range
  unsigned parts[2][2];
  int nparts = 0;
  const unsigned split = c->two_groups ? c->n_a : c->n;
  if (c->row_begin < std::min(c->row_end, split)) {
    parts[nparts][0] = c->row_begin;
    parts[nparts++][1] = std::min(c->row_end, split) - 1;
  }
  if (std::max(c->row_begin, split) < c->row_end) {
    parts[nparts][0] = std::max(c->row_begin, split);
    parts[nparts++][1] = c->row_end - 1;
  }
  for (int q = 0; q < nparts; ++q)
    for (int e = 0; e < 2; ++e)
      CU(c, cudaMemcpyAsync(c->h_idx + 2 * q + e, c->d_scell.p + parts[q][e], sizeof(unsigned), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  std::vector<char> need((size_t)n2, 0);
  const int R = g.radius;
  for (int q = 0; q < nparts; ++q) {
    const int zlo = (int)(c->h_idx[2 * q] / layer), zhi = (int)(c->h_idx[2 * q + 1] / layer);
    for (int z = zlo - R; z <= zhi + R; ++z) {
      if (g.stencil_pbc) need[(size_t)(((z % n2) + n2) % n2)] = 1;
      else if (z >= 0 && z < n2) need[(size_t)z] = 1;
    }
  }
  // runs of needed layers -> cell ranges -> sorted-index intervals in every group
  struct Run { int za, zb; };
  std::vector<Run> runs;
  for (int z = 0; z < n2;) {
    if (!need[(size_t)z]) { ++z; continue; }
    int zb = z;
    while (zb + 1 < n2 && need[(size_t)zb + 1]) ++zb;
    runs.push_back({z, zb});
    z = zb + 1;
  }
  if (runs.size() == 1 && runs[0].za == 0 && runs[0].zb == n2 - 1) return B200COORD_OK;
  const int ngroups = c->two_groups ? 2 : 1;
  if (runs.size() * ngroups > 6) return B200COORD_OK;
  int m = 0;
  for (int grp = 0; grp < ngroups; ++grp)
    for (const Run& r : runs) {
      const size_t first = (size_t)grp * g.ncell + (size_t)r.za * layer;
      CU(c, cudaMemcpyAsync(c->h_idx + 8 + 2 * m, c->d_cstart.p + first, sizeof(unsigned), cudaMemcpyDeviceToHost, c->st));
      if (r.zb + 1 < n2) {
        const size_t past = (size_t)grp * g.ncell + (size_t)(r.zb + 1) * layer;
        CU(c, cudaMemcpyAsync(c->h_idx + 8 + 2 * m + 1, c->d_cstart.p + past, sizeof(unsigned), cudaMemcpyDeviceToHost, c->st));
      }
      ++m;
    }
  CU(c, cudaStreamSynchronize(c->st));
  IdxRanges out;
  std::memset(&out, 0, sizeof(out));
  m = 0;
  for (int grp = 0; grp < ngroups; ++grp)
    for (const Run& r : runs) {
      const unsigned group_end = (grp == 0 && c->two_groups) ? c->n_a : c->n;
      const unsigned lo = c->h_idx[8 + 2 * m];
      const unsigned hi = (r.zb + 1 < n2) ? c->h_idx[8 + 2 * m + 1] : group_end;
      ++m;
      if (hi <= lo) continue;
      out.lo[out.n] = lo;
      out.len[out.n] = hi - lo;
      out.total += hi - lo;
      out.n++;
    }
  if (out.n == 0) return B200COORD_OK;
  c->needed = out;
  return B200COORD_OK;
}

// work list of the cell-tile sweep for the current sort (cells / no list, FP64, untyped pairings)
int build_tile_work(b200coord_ctx* c) {
  c->tile_ok = false;
  if (!c->tile_on || c->cfg.precision != B200COORD_FP64 || !sweep_tile_supports(c->dsw.type)) return B200COORD_OK;
  const unsigned ngroups = c->two_groups ? 2u : 1u;
  const unsigned nseg = ngroups * (unsigned)c->grid.ncell;
  const unsigned rows = c->row_end - c->row_begin;
  if (!rows) return B200COORD_OK;
  const unsigned rpi = tile_rows_per_item(rows);
  const unsigned split = c->two_groups ? c->n_a : c->n;
  const unsigned rows_a = std::min(c->row_end, split) > c->row_begin ? std::min(c->row_end, split) - c->row_begin : 0u;
  const unsigned rows_b = rows - rows_a;
  c->tile_bound_a = rows_a ? tile_items_bound(rows_a, (unsigned)c->grid.ncell, rpi) : 0u;
  c->tile_bound_b = rows_b ? tile_items_bound(rows_b, (unsigned)c->grid.ncell, rpi) : 0u;
  c->tile_nseg = nseg;
  CU(c, c->d_tilecnt.reserve(nseg + 1));
  CU(c, c->d_tilestart.reserve(nseg + 2));
  CU(c, c->d_tilework.reserve(16 * ((size_t)c->tile_bound_a + c->tile_bound_b + 2)));
  CU(c, c->d_bsum.reserve(nseg / 1024 + 4));
  launch_tile_work(nseg, c->d_cstart.p, c->d_ccount.p, c->row_begin, c->row_end, rpi, c->d_tilecnt.p, c->d_tilestart.p,
                   c->d_bsum.p, reinterpret_cast<TileWork*>(c->d_tilework.p), c->st);
  c->stats.kernel_launches += 5;
  CU_LAST(c, "tile work list");
  c->tile_ok = true;
  return B200COORD_OK;
}

// NeighborList::update (NeighborList.cpp:168-315) on the device
// capacity of the fixed-size rows of the next single-pass rebuild: 1.2 x the longest row seen now; kept when the
// current capacity is still adequate (the list buffer is sized by rows x capacity)
unsigned next_row_cap(unsigned max_row, unsigned current) {
  const unsigned want = ((max_row + max_row / 5 + 16u) + 7u) & ~7u;
  if (current >= max_row + max_row / 10 + 8u && current <= want + want / 4) return current;
  return want;
}

// [DevPbc | DevSwitch] in global memory for the out-of-line exact paths (row patches, band decisions of the builders)
int ensure_params(b200coord_ctx* c) {
  if (!c->params_dirty) return B200COORD_OK;
  CU(c, c->d_params.reserve(sizeof(DevPbc) + sizeof(DevSwitch)));
  CU(c, cudaMemcpyAsync(c->d_params.p, &c->dpbc, sizeof(DevPbc), cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->d_params.p + sizeof(DevPbc), &c->dsw, sizeof(DevSwitch), cudaMemcpyHostToDevice, c->st));
  c->params_dirty = false;
  return B200COORD_OK;
}

int rebuild(b200coord_ctx* c, const double* d_pos) {
  const int mode = c->cfg.nl_mode;
  {
    const int rcp = ensure_params(c);
    if (rcp) return rcp;
  }
  const DevPbc* pbc_g = reinterpret_cast<const DevPbc*>(c->d_params.p);
  if (c->cfg.style == B200COORD_STYLE_PAIR) {
    if (mode == B200COORD_NL_CLASSIC) {
      CU(c, c->d_active.reserve(c->n_a));
      CU(c, cudaEventRecord(c->ev[4], c->st));
      launch_pair_mask(d_pos, c->n_a, c->dpbc, c->cfg.nl_cutoff * c->cfg.nl_cutoff, c->d_active.p, c->st);
      CU(c, cudaEventRecord(c->ev[5], c->st));
      c->ev_valid[2] = true;
      c->stats.kernel_launches += 1;
      c->stats.rebuilds++;
    }
    c->list_valid = true;
    return B200COORD_OK;
  }
  if (mode == B200COORD_NL_NONE) {
    if (!c->sorted_valid) {
      int rc = setup_all_pairs(c);
      if (rc) return rc;
    }
    c->list_valid = true;
    return B200COORD_OK;
  }
  CU(c, cudaEventRecord(c->ev[4], c->st));
  const unsigned rows = c->row_end - c->row_begin;
  const double cut2 = c->cfg.nl_cutoff * c->cfg.nl_cutoff;  // NeighborList.cpp:238
  // super-list: wanted for the FP32-search NLIST path of systems whose sorted indices fit 26 bits
  const bool super_wanted = (mode == B200COORD_NL_CLASSIC) && c->super_on && !c->super_off && c->n <= kSuperIndexMask;
  c->super_delta = 0.1 * c->cfg.nl_cutoff;
  c->search_cutoff = c->cfg.nl_cutoff + (super_wanted ? c->super_delta : 0.0);
  float far2 = INFINITY;
  if (mode == B200COORD_NL_CLASSIC) {
    CU(c, c->d_rowcount.reserve(rows + 1));
    CU(c, c->d_rowstart.reserve(rows + 1));
    CU(c, c->d_rowfar.reserve(2 * (size_t)rows + 2));
    CU(c, c->d_bsum.reserve(rows / 1024 + 2));
    c->far_rows = rows;
    // near/far split of the rows: partners beyond D_MAX + a skin of a quarter of the list's buffer go to the far part
    if (c->far_split && c->sw.dmax > 0.0 && c->sw.dmax < c->cfg.nl_cutoff) {
      const double rf = c->sw.dmax + 0.25 * (c->cfg.nl_cutoff - c->sw.dmax);
      far2 = (float)(rf * rf);
      c->far_skin = rf - c->sw.dmax;
    } else {
      c->far_skin = 0.0;
    }
  }
  // The working list from a candidate source (`launch(mode)`: 0 count, 1 fill, 2 single pass into fixed-capacity rows):
  // single pass when a capacity is known (1.2 x the longest row of the previous rebuild; an overflow is detected on
  // the device and answered with the exact two-pass build), else count + scan + fill.
  auto build_rows = [&](auto&& launch, bool cappable) -> int {
    auto two_pass = [&]() -> int {
      B200_TRACE(c, "rebuild: count pass");
      launch(0);
      launch_scan_rows(c->d_rowcount.p, rows, 3u, c->d_bsum.p, c->d_rowstart.p, c->d_u64.p, c->st);
      CU(c, cudaMemcpyAsync(c->h_u64, c->d_u64.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->st));
      CU(c, cudaMemcpyAsync(c->h_capinfo, c->d_capinfo.p, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->st));
      CU(c, cudaStreamSynchronize(c->st));
      CU_LAST(c, "neighbour list count");
      c->nbr_total = c->h_u64[0];
      {
        // the following rebuilds use fixed-capacity rows: size the buffer for that layout now, so that the second
        // rebuild does not have to free and allocate gigabytes
        size_t need = (size_t)c->nbr_total;
        if (cappable) need = std::max(need, (size_t)rows * next_row_cap(c->h_capinfo[0], 0u));
        CU(c, c->d_nbr.reserve(need + 4));
      }
      B200_TRACE(c, "rebuild: fill pass, %llu entries", (unsigned long long)c->nbr_total);
      launch(1);
      c->stats.kernel_launches += 5;
      return B200COORD_OK;
    };
    CU(c, cudaMemsetAsync(c->d_capinfo.p, 0, 3 * sizeof(unsigned), c->st));
    bool done = false;
    if (cappable && c->row_cap > 0) {
      c->nbr_total = (unsigned long long)rows * c->row_cap;
      CU(c, c->d_nbr.reserve((size_t)c->nbr_total + 4));
      B200_TRACE(c, "rebuild: single pass into rows of %u", c->row_cap);
      launch(2);
      CU(c, cudaMemcpyAsync(c->h_capinfo, c->d_capinfo.p, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->st));
      CU(c, cudaStreamSynchronize(c->st));
      B200_TRACE(c, "rebuild: single pass done, longest row %u overflow %u", c->h_capinfo[0], c->h_capinfo[1]);
      CU_LAST(c, "neighbour list single-pass build");
      c->stats.kernel_launches += 2;
      done = (c->h_capinfo[1] == 0);
      if (!done) CU(c, cudaMemsetAsync(c->d_capinfo.p, 0, 3 * sizeof(unsigned), c->st));
    }
    if (!done) {
      int rc2 = two_pass();
      if (rc2) return rc2;
    }
    if (cappable) c->row_cap = next_row_cap(c->h_capinfo[0], c->row_cap);
    c->max_row = c->h_capinfo[0];
    return B200COORD_OK;
  };
  auto filter_launch = [&](int m) {
    launch_nl_filter(m, c->img_list, d_pos, c->d_perm.p, c->d_lpos.p, c->d_srowstart.p, c->d_srowcount.p, c->d_snbr.p, pbc_g, c->dbox, cut2,
                     c->band_rel, c->n_a, c->two_groups, c->row_begin, c->row_end, c->d_rowcount.p, c->d_rowstart.p,
                     m ? c->d_nbr.p : nullptr, m == 2 ? c->row_cap : 0u, c->d_capinfo.p, far2, c->d_rowfar.p,
                     c->d_rowfar.p + rows, c->filter_flat ? c->super_max_row : 0u, c->filter_minb, c->st);
  };
  bool list_done = false;
  if (super_wanted && c->super_valid) {
    // is the super-list still a superset?  same box, and nobody moved by delta/2 since it was built
    bool ok = (c->box_epoch == c->super_box_epoch);
    if (ok) {
      // frozen permutation: continuous coordinates, their float copy, and the largest displacement since the sort
      CU(c, cudaMemsetAsync(c->d_u64.p + 11, 0, sizeof(unsigned long long), c->st));
      launch_gather_u(true, pos_src_local(d_pos), c->d_perm.p, c->d_abs.p, c->needed, c->d_wpos.p, c->d_braw.p, c->dpbc,
                      c->d_spos.p, c->d_bpos.p, c->d_lpos.p, c->d_u64.p + 11, c->st);
      if (c->comm) {
        // every rank looked at the atoms it needs; all must take the same decision (the permutation is shared):
        // the largest displacement anywhere (non-negative doubles order like their bit patterns)
        NcclApi& api = nccl_api();
        ncclResult_t r = api.AllReduce(c->d_u64.p + 11, c->d_u64.p + 11, 1, ncclUint64, ncclMax, c->comm, c->st);
        B200_TRACE(c, "rebuild: displacement all-reduce enqueued");
        if (r != ncclSuccess) return fail(c, B200COORD_ERR_NCCL, std::string("ncclAllReduce(displacement): ") + api.GetErrorString(r));
        {
          const int rcd = collective_done(c);
          if (rcd) return rcd;
        }
      }
      CU(c, cudaMemcpyAsync(c->h_u64 + 2, c->d_u64.p + 11, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->st));
      CU(c, cudaStreamSynchronize(c->st));
      c->stats.kernel_launches += 1;
      double d2;
      std::memcpy(&d2, c->h_u64 + 2, sizeof(double));
      const double lim = 0.5 * c->super_delta * (1.0 - 1e-3);  // margin: FP32 search, rounding of the displacement
      ok = (d2 < lim * lim);
    }
    if (ok) {
      int rcf = build_rows(filter_launch, true);
      if (rcf) return rcf;
      c->super_streak = 0;
      c->filter_rebuilds++;
      list_done = true;
    } else {
      c->super_valid = false;
      if (++c->super_streak >= 2) {  // twice in a row: this run changes too fast for a super-list to pay
        c->super_off = true;
        c->search_cutoff = c->cfg.nl_cutoff;
      }
    }
  }
  if (!list_done) {
  int rc = setup_grid(c, d_pos);
  if (rc) return rc;
  rc = ensure_cell_arrays(c);
  if (rc) return rc;
  launch_sort(d_pos, c->n, c->n_a, c->two_groups ? 2 : 1, c->grid, c->d_cell_of_slot.p, c->d_ccount.p, c->d_cstart.p,
              c->d_cursor.p, c->d_tmp.p, c->d_perm.p, c->d_scell.p, c->d_bsum.p, c->st);
  launch_invert_perm(c->d_perm.p, c->n, c->d_inv.p, c->st);
  c->stats.kernel_launches += 5;
  c->sorted_valid = true;
  if (mode == B200COORD_NL_CLASSIC) {
    rc = compute_needed(c);
    if (rc) return rc;
  } else {
    needed_all(c);
    rc = build_tile_work(c);
    if (rc) return rc;
  }
  c->sq_valid = false;  // new permutation
  c->stype_valid = false;
  c->super_valid = false;
  if (mode == B200COORD_NL_CLASSIC) {
    const bool build_super = super_wanted && !c->super_off && c->f32_search;
    c->u_mode = c->f32_search;
    c->img_list = c->f32_search && c->n <= kSuperIndexMask;
    c->sort_box_epoch = c->box_epoch;
    CU(c, c->d_bpos.reserve(3 * (size_t)c->n));
    if (c->f32_search) {
      CU(c, c->d_lpos.reserve(c->n));
      CU(c, c->d_wpos.reserve(3 * (size_t)c->n));
      CU(c, c->d_braw.reserve(3 * (size_t)c->n));
      launch_sort_init(d_pos, c->d_perm.p, c->d_abs.p, c->n, c->grid, c->dbox, c->dpbc.type != 0, c->d_lpos.p, c->d_wpos.p,
                       c->d_braw.p, c->d_spos.p, c->d_bpos.p, c->st);
    } else {  // tiny boxes: FP64 search on the caller's positions, general sweep
      launch_gather_track(1, d_pos, c->d_perm.p, c->d_abs.p, c->n, c->d_spos.p, c->d_bpos.p, c->dpbc, c->d_u64.p + 10, c->st);
    }
    c->stats.kernel_launches += 1;
    if (!c->f32_search) CU(c, cudaMemsetAsync(c->d_rowfar.p, 0, 2 * (size_t)rows * sizeof(uint32_t), c->st));
    if (build_super) {
      // the super-list itself: two passes (count, scan, fill) over the cells with the extended cutoff
      const double sc2 = c->search_cutoff * c->search_cutoff;
      CU(c, c->d_srowcount.reserve(rows + 1));
      CU(c, c->d_srowstart.reserve(rows + 1));
      CU(c, cudaMemsetAsync(c->d_capinfo.p, 0, 3 * sizeof(unsigned), c->st));
      auto super_pass = [&](int m) {
        launch_nl_rows_f32(m, true, true, d_pos, c->d_perm.p, c->d_lpos.p, c->d_scell.p, c->d_cstart.p, c->d_ccount.p, c->grid, pbc_g,
                           c->dbox, sc2, c->band_rel, c->n_a, c->two_groups, c->row_begin, c->row_end, c->d_srowcount.p,
                           c->d_srowstart.p, m ? c->d_snbr.p : nullptr, m == 2 ? c->super_cap : 0u, c->d_capinfo.p, INFINITY,
                           nullptr, nullptr, c->st);
      };
      // like the working list: one pass into fixed-capacity rows once a capacity is known (1.2 x the longest row of the
      // previous build), the exact count + scan + fill on the first build and after an overflow
      bool sdone = false;
      if (c->super_cap > 0) {
        CU(c, c->d_snbr.reserve((size_t)rows * c->super_cap + 4));
        super_pass(2);
        CU(c, cudaMemcpyAsync(c->h_capinfo, c->d_capinfo.p, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaStreamSynchronize(c->st));
        CU_LAST(c, "super-list single-pass build");
        c->stats.kernel_launches += 2;
        sdone = (c->h_capinfo[1] == 0);
        if (!sdone) CU(c, cudaMemsetAsync(c->d_capinfo.p, 0, 3 * sizeof(unsigned), c->st));
      }
      if (!sdone) {
        super_pass(0);
        launch_scan_rows(c->d_srowcount.p, rows, 3u, c->d_bsum.p, c->d_srowstart.p, c->d_u64.p, c->st);
        CU(c, cudaMemcpyAsync(c->h_u64, c->d_u64.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaMemcpyAsync(c->h_capinfo, c->d_capinfo.p, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaStreamSynchronize(c->st));
        CU_LAST(c, "super-list count");
        CU(c, c->d_snbr.reserve(std::max((size_t)c->h_u64[0], (size_t)rows * next_row_cap(c->h_capinfo[0], 0u)) + 4));
        super_pass(1);
        c->stats.kernel_launches += 5;
      }
      c->super_cap = next_row_cap(c->h_capinfo[0], c->super_cap);
      c->super_max_row = c->h_capinfo[0];
      c->super_valid = true;
      c->super_box_epoch = c->box_epoch;
      c->super_builds++;
      int rcf = build_rows(filter_launch, true);  // d_lpos of make_local = wrapped position + zero displacement
      if (rcf) return rcf;
    } else if (c->f32_search) {
      int rcf = build_rows([&](int m) {
        launch_nl_rows_f32(m, false, c->img_list, d_pos, c->d_perm.p, c->d_lpos.p, c->d_scell.p, c->d_cstart.p, c->d_ccount.p, c->grid, pbc_g,
                           c->dbox, cut2, c->band_rel, c->n_a, c->two_groups, c->row_begin, c->row_end, c->d_rowcount.p,
                           c->d_rowstart.p, m ? c->d_nbr.p : nullptr, m == 2 ? c->row_cap : 0u, c->d_capinfo.p, far2,
                           c->d_rowfar.p, c->d_rowfar.p + rows, c->st);
      }, true);
      if (rcf) return rcf;
    } else {
      int rcf = build_rows([&](int m) {
        launch_nl_rows(m != 0, c->d_spos.p, c->d_scell.p, c->d_cstart.p, c->d_ccount.p, c->grid, c->dpbc, cut2, c->n_a,
                       c->two_groups, c->row_begin, c->row_end, c->d_rowcount.p, m ? c->d_rowstart.p : nullptr,
                       m ? c->d_nbr.p : nullptr, c->st);
      }, false);
      if (rcf) return rcf;
    }
  }
  }
  if (mode == B200COORD_NL_CLASSIC) {
    c->build_box_epoch = c->box_epoch;
    if (c->img_list) {
      CU(c, c->d_meta.reserve(rows + 1));
      launch_pack_meta(rows, c->d_rowstart.p, c->d_rowcount.p, c->d_rowfar.p, c->d_rowfar.p + rows, c->d_meta.p, c->d_u64.p + 13, c->st);
      c->stats.kernel_launches += 1;
    }
  }
  CU(c, cudaEventRecord(c->ev[5], c->st));
  c->ev_valid[2] = true;
  CU_LAST(c, "neighbour list rebuild");
  {  // rebuild steps already synchronise once to size the list; read the stopwatch here
    float ms = 0.f;
    CU(c, cudaEventSynchronize(c->ev[5]));
    if (cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess) {
      c->build_ms_acc += ms;
      c->build_ms_max = std::max(c->build_ms_max, ms);
      c->build_n++;
    }
  }
  c->stats.rebuilds++;
  c->list_valid = true;
  return B200COORD_OK;
}

// Several contexts in ONE process (b200coord_group_*, worker thread per device): a CUDA call that synchronises a device
// -- the first launch of a kernel under lazy module loading, a growing local-memory pool, cudaFree -- can sit on a
// driver lock while it waits for that device's NCCL kernel, which waits for a peer whose thread needs the same lock to
// launch its part.  As a precaution a thread of such a context does nothing else while a collective of its context is
// in flight.  One process per device (MPI, torchrun) keeps the collectives asynchronous.
int collective_done(b200coord_ctx* c) {
  if (c->in_process) CU(c, cudaStreamSynchronize(c->st));
  return B200COORD_OK;
}

int combine_ranks(b200coord_ctx* c) {
  NcclApi& api = nccl_api();
  ncclResult_t r;
  if (c->cfg.style == B200COORD_STYLE_PAIR) {
    r = api.AllReduce(c->d_out.p, c->d_out.p, (size_t)3 * c->n + 10, ncclDouble, ncclSum, c->comm, c->st);
    if (r != ncclSuccess) return fail(c, B200COORD_ERR_NCCL, std::string("ncclAllReduce: ") + api.GetErrorString(r));
    return collective_done(c);
  }
  // every rank owns complete derivatives for its rows: all-gather the row slices (sorted order), and
  // all-reduce the 10 scalars (virial + value) -- the Comm::Sum of CoordinationBase.cpp:218-224
  if (c->peer_mode) {  // rows were already stored into every peer by the sweep kernel; this all-reduce also orders the ranks
    r = api.AllReduce(c->d_out.p + (size_t)3 * c->n, c->d_out.p + (size_t)3 * c->n, 10, ncclDouble, ncclSum, c->comm, c->st);
    if (r != ncclSuccess) return fail(c, B200COORD_ERR_NCCL, std::string("ncclAllReduce: ") + api.GetErrorString(r));
    return collective_done(c);
  }
  api.GroupStart();
  r = api.AllGather(c->d_sderiv.p + (size_t)3 * c->row_chunk * c->cfg.rank, c->d_sderiv.p, (size_t)3 * c->row_chunk,
                    ncclDouble, c->comm, c->st);
  ncclResult_t r2 = api.AllReduce(c->d_out.p + (size_t)3 * c->n, c->d_out.p + (size_t)3 * c->n, 10, ncclDouble, ncclSum,
                                  c->comm, c->st);
  ncclResult_t r3 = api.GroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess || r3 != ncclSuccess)
    return fail(c, B200COORD_ERR_NCCL, std::string("nccl combine: ") +
                                           api.GetErrorString(r != ncclSuccess ? r : (r2 != ncclSuccess ? r2 : r3)));
  return collective_done(c);
}

// the whole per-step device pipeline on c->st; d_pos device positions, result left in c->d_out
// out: where the 3n derivatives go (null = c->d_out; the 10 tail doubles always land in c->d_out);
// [slot_lo, slot_lo+slot_cnt): the slots of the derivative array the caller is going to read
// d_pos: the whole position array on this device, or null on a step of the distributed engine that keeps its list and
// pulls positions from the owners' slice buffers (`src`)
int run_device(b200coord_ctx* c, const double* d_pos, double* out = nullptr, unsigned slot_lo = 0u,
               unsigned slot_cnt = 0xffffffffu, const PosSrc* src_in = nullptr) {
  const PosSrc src = src_in ? *src_in : pos_src_local(d_pos);
  c->coupled_fresh = false;
  if (c->cfg.pbc && !c->box_set) return fail(c, B200COORD_ERR_STATE, "b200coord_set_box must be called before calculate");
  if (c->dsw.type == B200COORD_PAIR_DHENERGY && !c->have_charges)
    return fail(c, B200COORD_ERR_STATE, "b200coord_set_charges must be called before calculate (DHENERGY)");
  if (c->dsw.type == B200COORD_PAIR_GHBFIX && !c->have_types)
    return fail(c, B200COORD_ERR_STATE, "b200coord_set_types must be called before calculate (GHBFIX)");
  const bool need_rebuild = !c->list_valid || (c->cfg.nl_mode != B200COORD_NL_NONE && c->invalidate);
  if (!d_pos && (need_rebuild || c->cfg.style == B200COORD_STYLE_PAIR))
    return fail(c, B200COORD_ERR_STATE, "internal: a rebuild step needs the whole position array");
  if (need_rebuild) {
    B200_TRACE(c, "rebuild starts");
    int rc = rebuild(c, d_pos);
    if (rc) return rc;
    B200_TRACE(c, "rebuild done");
    c->invalidate = false;
  }
  {
    const int rcp = ensure_params(c);
    if (rcp) return rcp;
  }
  const DevPbc* pbc_g = reinterpret_cast<const DevPbc*>(c->d_params.p);
  const DevSwitch* sw_g = reinterpret_cast<const DevSwitch*>(c->d_params.p + sizeof(DevPbc));
  CU(c, cudaMemsetAsync(c->d_u64.p + 1, 0, sizeof(unsigned long long), c->st));
  CU(c, cudaMemsetAsync(c->d_u64.p + 12, 0, sizeof(unsigned long long), c->st));
  int nblocks;
  double weight;
  if (c->cfg.style == B200COORD_STYLE_PAIR) {
    const unsigned chunk = (c->n_a + (unsigned)c->cfg.nranks - 1) / (unsigned)c->cfg.nranks;
    const unsigned pb = std::min(c->n_a, chunk * (unsigned)c->cfg.rank), pe = std::min(c->n_a, pb + chunk);
    if (c->cfg.nranks > 1) CU(c, cudaMemsetAsync(c->d_out.p, 0, sizeof(double) * 3 * (size_t)c->n, c->st));
    CU(c, c->d_partials.reserve((size_t)kPartialStride * ((pe - pb) / 256 + 2)));
    CU(c, cudaEventRecord(c->ev[2], c->st));
    CU(c, cudaEventRecord(c->sweep_ev[2 * (c->sweep_n % b200coord_ctx::kRing)], c->st));
    nblocks = launch_sweep_pairs(d_pos, c->d_q.p, c->d_types.p, c->d_etas.p, c->ntypes, c->d_abs.p, c->cfg.nl_mode == B200COORD_NL_CLASSIC ? c->d_active.p : nullptr,
                                 c->n_a, pb, pe, c->dpbc, c->dsw, c->d_out.p, c->d_partials.p, c->d_u64.p + 1, c->st);
    CU(c, cudaEventRecord(c->ev[3], c->st));
    weight = 1.0;
    c->stats.kernel_launches += 1;
  } else {
    const bool classic = (c->cfg.nl_mode == B200COORD_NL_CLASSIC);
    // box changed since the sort: the continuous coordinates (and the images) no longer mean anything
    const bool u_now = classic && c->u_mode && c->box_epoch == c->sort_box_epoch;
    if (!d_pos && !(u_now && !need_rebuild))
      return fail(c, B200COORD_ERR_STATE, "internal: this step needs the whole position array");
    if (classic) {
      if (need_rebuild) {  // rebuild() left records, displacement origin and a zero displacement behind
        CU(c, cudaMemsetAsync(c->d_u64.p + 10, 0, sizeof(unsigned long long), c->st));
      } else if (u_now) {
        CU(c, cudaMemsetAsync(c->d_u64.p + 10, 0, sizeof(unsigned long long), c->st));
        launch_gather_u(false, src, c->d_perm.p, c->d_abs.p, c->needed, c->d_wpos.p, c->d_braw.p, c->dpbc, c->d_spos.p,
                        c->d_bpos.p, nullptr, c->d_u64.p + 10, c->st);
      } else if (c->u_mode) {  // far parts are visited anyway (force_far): no displacement needed
        launch_gather(d_pos, c->d_perm.p, c->d_abs.p, c->n, c->d_spos.p, c->st);
      } else {
        CU(c, cudaMemsetAsync(c->d_u64.p + 10, 0, sizeof(unsigned long long), c->st));
        launch_gather_track(2, d_pos, c->d_perm.p, c->d_abs.p, c->n, c->d_spos.p, c->d_bpos.p, c->dpbc, c->d_u64.p + 10, c->st);
      }
    } else {
      launch_gather(d_pos, c->d_perm.p, c->d_abs.p, c->n, c->d_spos.p, c->st);
    }
    SweepArgs a;
    std::memset(&a, 0, sizeof(a));
    a.spos = c->d_spos.p;
    if (c->dsw.type == B200COORD_PAIR_DHENERGY) {
      if (!c->sq_valid) {
        CU(c, c->d_sq.reserve(c->n));
        launch_gather_charges(c->d_q.p, c->d_perm.p, c->n, c->d_sq.p, c->st);
        c->sq_valid = true;
      }
      a.sq = c->d_sq.p;
    }
    if (c->dsw.type == B200COORD_PAIR_GHBFIX) {
      if (!c->stype_valid) {
        CU(c, c->d_stype.reserve(c->n));
        launch_gather_types(c->d_types.p, c->d_perm.p, c->n, c->d_stype.p, c->st);
        c->stype_valid = true;
      }
      a.stype = c->d_stype.p;
      a.etas = c->d_etas.p;
      a.ntypes = c->ntypes;
    }
    a.n_a = c->n_a;
    a.two_groups = c->two_groups;
    a.check_abs = c->check_abs;
    a.row_begin = c->row_begin;
    a.row_end = c->row_end;
    a.row_start = c->d_rowstart.p;
    a.row_count = c->d_rowcount.p;
    a.nbr = c->d_nbr.p;
    a.row_far_off = c->d_rowfar.p;
    a.row_far_cnt = c->d_rowfar.p + c->far_rows;
    // a pair with r^2 above this is beyond D_MAX and outside the band in which the exact patch decides (sweep_math.cuh)
    a.far_skip2 = (c->dsw.band_dmax >= 0.0) ? c->dsw.dmax_2 + c->dsw.band_dmax : INFINITY;
    a.f32 = (c->cfg.precision == B200COORD_FP32) ? 1 : 0;
    if (a.f32 && c->dsw.band_dmax >= 0.0) a.far_skip2 = c->dsw.dmax_2 * (1.0 + 1e-6);  // above D_MAX^2 after rounding to float
    a.disp2_bits = c->d_u64.p + 10;
    {
      // The near/far split was decided on the FP32 r^2 of the search, which is off by up to band_rel * cutoff^2,
      // i.e. r by up to ~band_rel * cutoff: that much of the skin is not available to the atoms.
      const double half = 0.5 * (c->far_skin - c->band_rel * c->cfg.nl_cutoff) * (1.0 - 1e-3);
      a.far_disp2_max = (half > 0.0) ? half * half : 0.0;
    }
    a.force_far = (c->box_epoch != c->build_box_epoch) ? 1 : 0;
    a.idx_mask = c->img_list ? kSuperIndexMask : 0xffffffffu;
    a.row_meta = c->d_meta.p;
    a.pos = src;
    a.executed = c->d_u64.p + 12;
    // image sweep: valid while a listed pair cannot have a second image as close as the stored one, i.e. while
    // NL_CUTOFF + 2 * displacement < half the smallest box height (any lattice vector is at least that long)
    a.img_disp2_max = 0.0;
    const bool img_now = classic && u_now && c->img_list && c->img_on && !a.f32;
    if (img_now) {
      if (c->dpbc.type == 0) {
        a.img_disp2_max = INFINITY;
      } else {
        double hmin = INFINITY;
        for (int k = 0; k < 3; ++k) {
          const double* ib = c->hpbc.inv_box;
          hmin = std::min(hmin, 1.0 / std::sqrt(ib[k] * ib[k] + ib[3 + k] * ib[3 + k] + ib[6 + k] * ib[6 + k]));
        }
        const double lim = 0.5 * (0.5 * hmin * (1.0 - 1e-3) - c->cfg.nl_cutoff);
        a.img_disp2_max = (lim > 0.0) ? lim * lim : 0.0;
      }
      // one block shape for both kernels (they fill the same partial records)
      const unsigned acc_end = c->two_groups ? std::min(c->row_end, c->n_a) : c->row_end;
      const unsigned rows_a = acc_end > c->row_begin ? acc_end - c->row_begin : 0u;
      const unsigned rows_b = c->row_end - std::max(c->row_begin, acc_end);
      a.rows_per_block = sweep_img_rows_per_block(rows_a, rows_b, c->max_row);
      if (a.rows_per_block == 0u) a.img_disp2_max = 0.0;  // rows too long for the image sweep's trip table
      // few GROUPA atoms among many GROUPB atoms (a solute in its solvent): GROUPB rows are a handful of entries each
      a.scatter_b = (c->two_groups && c->cfg.nranks == 1 && c->scatter_on && (unsigned long long)c->n_a * 8ull <= c->n_b) ? 1 : 0;
    }
    a.scell = c->d_scell.p;
    a.cstart = c->d_cstart.p;
    a.ccount = c->d_ccount.p;
    a.grid = c->grid;
    double* rows_now = c->d_sderiv.p;
    if (c->peer_mode) {  // rows stay where they are computed; the un-sort of every rank pulls the ones it returns
      c->parity ^= 1u;
      rows_now = c->peer_rows[c->parity][c->cfg.rank];
    }
    a.sderiv = rows_now;
    a.evals = c->d_u64.p + 1;
    a.pbc_g = pbc_g;
    a.sw_g = sw_g;
    const size_t npart = (size_t)kPartialStride * ((c->row_end - c->row_begin) / 8 + 8 + 148);
    CU(c, c->d_partials.reserve(npart));
    a.partials = c->d_partials.p;
    CU(c, cudaEventRecord(c->ev[2], c->st));
    CU(c, cudaEventRecord(c->sweep_ev[2 * (c->sweep_n % b200coord_ctx::kRing)], c->st));
    if (classic) {
      if (a.img_disp2_max > 0.0) {
        // Two kernels, one of which returns at once (device-side gate on the displacement): the image sweep skips
        // empty rows and the general kernel only stores partials for GROUPA rows, so rows and partials start at zero.
        CU(c, cudaMemsetAsync(c->d_partials.p, 0, npart * sizeof(double), c->st));
        CU(c, cudaMemsetAsync(rows_now + 3 * (size_t)c->row_begin, 0, sizeof(double) * 3 * (size_t)(c->row_end - c->row_begin), c->st));
        nblocks = launch_sweep_img(a, c->dbox, c->dsw, c->img_variant, c->st);
        const int nb2 = launch_sweep_list(a, c->dpbc, c->dsw, c->st);
        if (nb2 < 0) nblocks = nb2;
        c->stats.kernel_launches += 1 + (c->two_groups ? 2 : 1);
      } else {
        nblocks = launch_sweep_list(a, c->dpbc, c->dsw, c->st);
      }
    } else if (c->tile_ok) {
      // cell-tile sweep: the launch covers an upper bound of the item count, blocks past the end leave at once and
      // write no partial record
      CU(c, c->d_partials.reserve((size_t)kPartialStride * ((size_t)c->tile_bound_a + 2)));
      a.partials = c->d_partials.p;
      CU(c, cudaMemsetAsync(c->d_partials.p, 0, sizeof(double) * kPartialStride * ((size_t)c->tile_bound_a + 2), c->st));
      CU(c, cudaMemsetAsync(c->d_u64.p + 15, 0, sizeof(unsigned long long), c->st));
      // a solute in its solvent: GROUPA rows scatter to their partners, GROUPB rows (zeroed here) are not swept
      a.scatter_b = (c->two_groups && c->cfg.nranks == 1 && c->scatter_on && (unsigned long long)c->n_a * 8ull <= c->n_b) ? 1 : 0;
      a.inv = c->d_inv.p;
      if (a.scatter_b)
        CU(c, cudaMemsetAsync(rows_now + 3 * (size_t)c->n_a, 0, sizeof(double) * 3 * (size_t)c->n_b, c->st));
      const unsigned long long* total_dev = c->d_tilestart.p + c->tile_nseg;
      const unsigned long long* split_dev = c->d_tilestart.p + (unsigned)c->grid.ncell;
      nblocks = launch_sweep_tile(a, c->dpbc, c->dsw, reinterpret_cast<const TileWork*>(c->d_tilework.p), c->d_u64.p + 15,
                                  split_dev, total_dev, c->tile_bound_a, c->tile_bound_b, c->st);
    } else {
      nblocks = launch_sweep_cells(a, c->dpbc, c->dsw, c->st);
    }
    CU(c, cudaEventRecord(c->ev[3], c->st));
    weight = c->two_groups ? 1.0 : 0.5;
    c->stats.kernel_launches += 1 + (c->two_groups ? 2 : 1);
  }
  CU(c, cudaEventRecord(c->sweep_ev[2 * (c->sweep_n % b200coord_ctx::kRing) + 1], c->st));
  c->sweep_n++;
  c->ev_valid[1] = true;
  if (nblocks < 0) return fail(c, B200COORD_ERR_UNSUPPORTED, "switching function type has no GPU kernel");
  CU_LAST(c, "pair sweep launch");
  c->sweep_blocks = nblocks;
  CU(c, c->d_fin.reserve(16 * ((size_t)nblocks / 256 + 2)));
  launch_finalize(c->d_partials.p, nblocks, weight, c->d_out.p + (size_t)3 * c->n, c->d_fin.p, c->st);
  c->stats.kernel_launches += 2;
  if (c->comm) {
    B200_TRACE(c, "sweep launched (%d blocks), combining", nblocks);
    int rc = combine_ranks(c);
    if (rc) return rc;
    B200_TRACE(c, "combined");
  }
  if (c->cfg.style != B200COORD_STYLE_PAIR) {
    RowSrc rows;
    for (int r = 0; r < 8; ++r)
      rows.base[r] = (c->peer_mode && r < c->cfg.nranks) ? c->peer_rows[c->parity][r] : c->d_sderiv.p;
    rows.chunk = c->row_chunk ? c->row_chunk : 1u;
    const unsigned lo = std::min(slot_lo, c->n);
    const unsigned cnt = std::min(slot_cnt, c->n - lo);
    launch_unsort_pull(rows, c->d_inv.p, out ? out : c->d_out.p, lo, cnt, c->st);
    c->stats.kernel_launches += 1;
  }
  CU(c, cudaMemcpyAsync(c->h_u64 + 1, c->d_u64.p + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(c->h_u64 + 3, c->d_u64.p + 12, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->st));
  CU_LAST(c, "finalize");
  return B200COORD_OK;
}

void maybe_pin(b200coord_ctx* c, int which, const void* p, size_t bytes) {
  if (!c->pin_host || !p || !bytes) return;
  b200coord_ctx::Pinned& e = c->pinned[which];
  if (e.p == p && e.bytes == bytes) return;
  if (e.p) cudaHostUnregister(const_cast<void*>(e.p));
  e.p = nullptr;
  e.bytes = 0;
  if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) {
    e.p = p;
    e.bytes = bytes;
  }
  cudaGetLastError();  // a failed registration just leaves the copy pageable
}

void refresh_stats(b200coord_ctx* c) {
  // [1] entries of the rows the general / cell kernels swept, [3] entries actually evaluated, [4] entries of the
  // image-mode list (counted when it was built)
  const bool img = (c->cfg.nl_mode == B200COORD_NL_CLASSIC && c->img_list && c->cfg.style != B200COORD_STYLE_PAIR);
  const unsigned long long listed = img ? c->h_u64[4] : c->h_u64[1];
  c->stats.pair_evals = (c->cfg.style == B200COORD_STYLE_PAIR) ? c->h_u64[1] : c->h_u64[3];
  const unsigned long long n = c->n;
  switch (c->cfg.nl_mode) {
    case B200COORD_NL_CLASSIC:
      c->stats.nl_size = (c->cfg.style == B200COORD_STYLE_PAIR) ? c->h_u64[1] : listed / 2 + (c->cfg.rank == 0 ? c->n_self_pairs : 0);
      break;
    case B200COORD_NL_CELLS:
      c->stats.nl_size = c->two_groups ? c->h_u64[1] / 2 : (c->h_u64[1] - (c->row_end - c->row_begin)) / 2;
      break;
    default:
      if (c->cfg.style == B200COORD_STYLE_PAIR) c->stats.nl_size = c->n_a;
      else if (c->two_groups) c->stats.nl_size = (unsigned long long)c->n_a * c->n_b;
      else c->stats.nl_size = n * (n - 1) / 2;
  }
  c->stats.pbc_type = c->hpbc.type;
  c->stats.f32_search = c->f32_search ? 1 : 0;
  c->stats.super_builds = c->super_builds;
  c->stats.filter_rebuilds = c->filter_rebuilds;
}

}  // namespace

// ================================================================================================
extern "C" {

int b200coord_abi_version(void) { return B200COORD_ABI_VERSION; }

int b200coord_switch_parse(const char* definition, b200coord_switch* out, char* err, size_t errlen) {
  if (!definition || !out) return fail(nullptr, B200COORD_ERR_INVALID, "null argument");
  std::string e;
  const int rc = parse_switch(definition, *out, e);
  if (err && errlen) std::snprintf(err, errlen, "%s", e.c_str());
  if (rc) g_last_error = e;
  return rc;
}

int b200coord_switch_rational(int nn, int mm, double r0, double d0, b200coord_switch* out) {
  if (!out) return fail(nullptr, B200COORD_ERR_INVALID, "null argument");
  if (!(r0 > 0.0)) return fail(nullptr, B200COORD_ERR_INVALID, "R_0 should be explicitly specified and positive");
  rational_switch(nn, mm, r0, d0, *out);
  return B200COORD_OK;
}

int b200coord_pairing_dhenergy(double ionic_strength, double temp, double epsilon, double energy_unit, double length_unit,
                               double charge_unit, b200coord_switch* out) {
  if (!out) return fail(nullptr, B200COORD_ERR_INVALID, "null argument");
  if (!(epsilon > 0.0) || !(temp > 0.0) || !(ionic_strength >= 0.0) || !(energy_unit > 0.0) || !(length_unit > 0.0))
    return fail(nullptr, B200COORD_ERR_INVALID, "DHENERGY needs EPSILON > 0, TEMP > 0, I >= 0");
  dhenergy_pairing(ionic_strength, temp, epsilon, energy_unit, length_unit, charge_unit, *out);
  return B200COORD_OK;
}

int b200coord_pairing_ghbfix(double dmax, double d0, double c, b200coord_switch* out) {
  if (!out) return fail(nullptr, B200COORD_ERR_INVALID, "null argument");
  if (!(dmax > d0) || !(c > 0.0) || !(c < 1.0))
    return fail(nullptr, B200COORD_ERR_INVALID, "GHBFIX needs D_MAX > D_0 and 0 < C < 1");
  ghbfix_pairing(dmax, d0, c, *out);
  return B200COORD_OK;
}

int b200coord_switch_describe(const b200coord_switch* sw, char* buf, size_t buflen) {
  if (!sw || !buf || !buflen) return B200COORD_ERR_INVALID;
  std::snprintf(buf, buflen, "%s", describe_switch(*sw).c_str());
  return B200COORD_OK;
}

const char* b200coord_last_error(const b200coord_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int b200coord_create(const b200coord_config* cfg, const b200coord_switch* sw, const unsigned* abs_index,
                     b200coord_ctx** out) {
  if (!cfg || !sw || !out) return fail(nullptr, B200COORD_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->abi_version != B200COORD_ABI_VERSION) return fail(nullptr, B200COORD_ERR_INVALID, "ABI version mismatch");
  if (cfg->style < 0 || cfg->style > 2) return fail(nullptr, B200COORD_ERR_INVALID, "unknown list style");
  // the keyword rules of CoordinationBase.cpp:69-83 and NeighborList.cpp:71-79
  if (cfg->nl_mode == B200COORD_NL_CELLS && cfg->style == B200COORD_STYLE_PAIR)
    return fail(nullptr, B200COORD_ERR_INVALID, "Pair is not compatible with the CELLS implementation of the NL");
  if (cfg->nl_mode != B200COORD_NL_NONE) {
    if (!(cfg->nl_cutoff > 0.0)) return fail(nullptr, B200COORD_ERR_INVALID, "NL_CUTOFF should be explicitly specified and positive");
    if (cfg->nl_stride <= 0) return fail(nullptr, B200COORD_ERR_INVALID, "NL_STRIDE should be explicitly specified and positive");
  }
  if (cfg->style == B200COORD_STYLE_PAIR && cfg->n_group_a != cfg->n_group_b)
    return fail(nullptr, B200COORD_ERR_INVALID,
                "when using PAIR option, the two groups should have the same number of elements");
  if (cfg->style == B200COORD_STYLE_SINGLELIST && cfg->n_group_b != 0)
    return fail(nullptr, B200COORD_ERR_INVALID, "SINGLELIST style takes GROUPA only");
  if ((sw->type < 0 || sw->type >= B200COORD_SW_LEPTON) && sw->type != B200COORD_PAIR_DHENERGY &&
      sw->type != B200COORD_PAIR_GHBFIX)
    return fail(nullptr, B200COORD_ERR_UNSUPPORTED, "switching function is not available on the GPU");
  if (cfg->precision != B200COORD_FP64 && cfg->precision != B200COORD_FP32)
    return fail(nullptr, B200COORD_ERR_INVALID, "precision must be B200COORD_FP64 or B200COORD_FP32");
  const unsigned long long ntot = (unsigned long long)cfg->n_group_a + cfg->n_group_b;
  if (ntot == 0 || ntot > 0x7fffffffull) return fail(nullptr, B200COORD_ERR_INVALID, "atom count out of range");

  b200coord_ctx* c = new b200coord_ctx();
  c->cfg = *cfg;
  if (c->cfg.nranks < 1) c->cfg.nranks = 1;
  if (c->cfg.rank < 0 || c->cfg.rank >= c->cfg.nranks) {
    delete c;
    return fail(nullptr, B200COORD_ERR_INVALID, "rank out of range");
  }
  c->sw = *sw;
  to_dev_switch(*sw, c->dsw);
  c->n_a = cfg->n_group_a;
  c->n_b = cfg->n_group_b;
  c->n = (unsigned)ntot;
  c->two_groups = (cfg->style == B200COORD_STYLE_TWOLIST) ? 1 : 0;
  std::memset(&c->stats, 0, sizeof(c->stats));
  c->abs_host.resize(c->n);
  for (unsigned i = 0; i < c->n; ++i) c->abs_host[i] = abs_index ? abs_index[i] : i;
  {  // do the two groups share atoms (TwoList) / does GROUPA repeat atoms (SingleList)?
    if (c->two_groups) {
      std::vector<unsigned> a(c->abs_host.begin(), c->abs_host.begin() + c->n_a);
      std::vector<unsigned> b(c->abs_host.begin() + c->n_a, c->abs_host.end());
      std::sort(a.begin(), a.end());
      std::sort(b.begin(), b.end());
      size_t i = 0, j = 0;
      while (i < a.size() && j < b.size()) {
        if (a[i] < b[j]) ++i;
        else if (b[j] < a[i]) ++j;
        else {
          size_t i2 = i, j2 = j;
          while (i2 < a.size() && a[i2] == a[i]) ++i2;
          while (j2 < b.size() && b[j2] == b[j]) ++j2;
          c->n_self_pairs += (unsigned long long)(i2 - i) * (j2 - j);
          i = i2;
          j = j2;
        }
      }
    } else if (cfg->style == B200COORD_STYLE_SINGLELIST) {
      std::vector<unsigned> a(c->abs_host);
      std::sort(a.begin(), a.end());
      for (size_t i = 0; i < a.size();) {
        size_t i2 = i;
        while (i2 < a.size() && a[i2] == a[i]) ++i2;
        c->n_self_pairs += (unsigned long long)(i2 - i) * (i2 - i - 1) / 2;
        i = i2;
      }
    }
    c->check_abs = c->n_self_pairs ? 1 : 0;
  }
  c->row_chunk = (c->n + (unsigned)c->cfg.nranks - 1) / (unsigned)c->cfg.nranks;
  c->row_begin = std::min(c->n, c->row_chunk * (unsigned)c->cfg.rank);
  c->row_end = std::min(c->n, c->row_begin + c->row_chunk);
  c->slot_begin = c->row_begin;  // same equal-chunk partition, applied to slots instead of sorted rows
  c->slot_count = c->row_end - c->row_begin;

#define CREATE_CU(expr)                                                                        \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      std::string m__ = std::string(#expr) + ": " + cudaGetErrorString(e__);                   \
      b200coord_destroy(c);                                                                    \
      return fail(nullptr, B200COORD_ERR_CUDA, m__);                                           \
    }                                                                                          \
  } while (0)

  int ndev = 0;
  CREATE_CU(cudaGetDeviceCount(&ndev));
  if (ndev == 0) {
    b200coord_destroy(c);
    return fail(nullptr, B200COORD_ERR_CUDA, "no CUDA device visible; libb200coord has no CPU path");
  }
  if (cfg->device >= 0) {
    CREATE_CU(cudaSetDevice(cfg->device));
    c->device = cfg->device;
  } else {
    CREATE_CU(cudaGetDevice(&c->device));
  }
  CREATE_CU(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
  for (int i = 0; i < 10; ++i) CREATE_CU(cudaEventCreate(&c->ev[i]));
  for (int i = 0; i < 2 * b200coord_ctx::kRing; ++i) CREATE_CU(cudaEventCreate(&c->sweep_ev[i]));
  const size_t n = c->n;
  const size_t padded_rows = (size_t)c->row_chunk * (size_t)c->cfg.nranks;
  CREATE_CU(c->d_pos.reserve(3 * std::max(n, padded_rows)));
  CREATE_CU(c->d_out.reserve(3 * n + 10));
  CREATE_CU(c->d_sderiv.reserve(3 * padded_rows));
  CREATE_CU(c->d_small.reserve(32));
  CREATE_CU(c->d_u64.reserve(16));
  CREATE_CU(c->d_abs.reserve(n));
  CREATE_CU(c->d_perm.reserve(n));
  CREATE_CU(c->d_inv.reserve(n));
  CREATE_CU(cudaHostAlloc((void**)&c->h_idx, 32 * sizeof(unsigned), cudaHostAllocDefault));
  CREATE_CU(c->d_scell.reserve(n));
  CREATE_CU(c->d_cell_of_slot.reserve(n));
  CREATE_CU(c->d_tmp.reserve(n));
  CREATE_CU(c->d_spos.reserve(n));
  CREATE_CU(c->d_rowcount.reserve(1));
  CREATE_CU(c->d_rowstart.reserve(1));
  CREATE_CU(c->d_nbr.reserve(1));
  CREATE_CU(cudaMemsetAsync(c->d_sderiv.p, 0, sizeof(double) * 3 * padded_rows, c->st));
  CREATE_CU(cudaHostAlloc((void**)&c->h_small, 32 * sizeof(double), cudaHostAllocDefault));
  CREATE_CU(cudaHostAlloc((void**)&c->h_u64, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
  CREATE_CU(cudaHostAlloc((void**)&c->h_capinfo, 4 * sizeof(unsigned), cudaHostAllocDefault));
  CREATE_CU(c->d_capinfo.reserve(4));
  c->h_u64[0] = c->h_u64[1] = 0;
  CREATE_CU(cudaMemcpyAsync(c->d_abs.p, c->abs_host.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->st));
  CREATE_CU(cudaStreamSynchronize(c->st));
#undef CREATE_CU
  std::memset(&c->hpbc, 0, sizeof(c->hpbc));
  to_dev_pbc(c->hpbc, false, c->dpbc);
  if (const char* e = std::getenv("B200COORD_PIN_HOST")) c->pin_host = (std::atoi(e) != 0);
  if (const char* e = std::getenv("B200COORD_FILTER_FLAT")) c->filter_flat = (std::atoi(e) != 0);
  if (const char* e = std::getenv("B200COORD_FILTER_MINB")) c->filter_minb = (std::atoi(e) == 3) ? 3 : 2;
  if (const char* e = std::getenv("B200COORD_TEST_FAIL")) {  // "<rank>:<call>"
    int rk = -1, call = -1;
    if (std::sscanf(e, "%d:%d", &rk, &call) == 2 && rk == cfg->rank) c->test_fail_at = call;
  }
  if (const char* e = std::getenv("B200COORD_NO_FAR_SPLIT")) c->far_split = (std::atoi(e) == 0);
  if (const char* e = std::getenv("B200COORD_NO_SUPERLIST")) c->super_on = (std::atoi(e) == 0);
  if (const char* e = std::getenv("B200COORD_NO_IMG_SWEEP")) c->img_on = (std::atoi(e) == 0);
  if (const char* e = std::getenv("B200COORD_NO_TILE_SWEEP")) c->tile_on = (std::atoi(e) == 0);
  if (const char* e = std::getenv("B200COORD_NO_SCATTER")) c->scatter_on = (std::atoi(e) == 0);
  if (const char* e = std::getenv("B200COORD_IMG_VARIANT")) c->img_variant = std::atoi(e);
  needed_all(c);
  *out = c;
  return B200COORD_OK;
}

int b200coord_set_charges(b200coord_ctx* c, const double* charges) {
  if (!c || !charges) return fail(c, B200COORD_ERR_INVALID, "null argument");
  CU(c, cudaSetDevice(c->device));
  CU(c, c->d_q.reserve(c->n));
  CU(c, cudaMemcpyAsync(c->d_q.p, charges, sizeof(double) * (size_t)c->n, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaStreamSynchronize(c->st));  // the caller's array may go away
  c->have_charges = true;
  c->sq_valid = false;
  return B200COORD_OK;
}

int b200coord_set_types(b200coord_ctx* c, const unsigned* types, unsigned ntypes, const double* etas) {
  if (!c || !types || !etas || !ntypes) return fail(c, B200COORD_ERR_INVALID, "null argument");
  for (unsigned i = 0; i < c->n; ++i)
    if (types[i] >= ntypes) return fail(c, B200COORD_ERR_INVALID, "interaction type out of range");
  CU(c, cudaSetDevice(c->device));
  CU(c, c->d_types.reserve(c->n));
  CU(c, c->d_etas.reserve((size_t)ntypes * ntypes));
  CU(c, cudaMemcpyAsync(c->d_types.p, types, sizeof(uint32_t) * (size_t)c->n, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaMemcpyAsync(c->d_etas.p, etas, sizeof(double) * (size_t)ntypes * ntypes, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaStreamSynchronize(c->st));  // the caller's arrays may go away
  c->ntypes = ntypes;
  c->have_types = true;
  c->stype_valid = false;
  return B200COORD_OK;
}

void b200coord_destroy(b200coord_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->st) cudaStreamSynchronize(c->st);
  for (auto& e : c->pinned)
    if (e.p) cudaHostUnregister(const_cast<void*>(e.p));
  cudaGetLastError();
  if (c->comm && !c->comm_aborted && nccl_api().ok) nccl_api().CommDestroy(c->comm);
  c->d_pos.release(); c->d_out.release(); c->d_sderiv.release(); c->d_partials.release(); c->d_small.release();
  c->d_abs.release(); c->d_perm.release(); c->d_scell.release(); c->d_cell_of_slot.release(); c->d_tmp.release();
  c->d_ccount.release(); c->d_cstart.release(); c->d_cursor.release(); c->d_rowcount.release(); c->d_nbr.release();
  c->d_rowstart.release(); c->d_bsum.release(); c->d_u64.release(); c->d_spos.release(); c->d_active.release(); c->d_params.release(); c->d_lpos.release(); c->d_capinfo.release(); c->d_rowfar.release(); c->d_bpos.release(); c->d_meta.release(); c->d_srowstart.release(); c->d_srowcount.release(); c->d_snbr.release(); c->d_wpos.release(); c->d_braw.release(); c->d_q.release(); c->d_sq.release(); c->d_types.release(); c->d_stype.release(); c->d_etas.release();
  if (c->peer_mode && c->peer_ipc)
    for (int par = 0; par < 2; ++par)
      for (int r = 0; r < c->cfg.nranks && r < 8; ++r)
        if (r != c->cfg.rank && c->peer_rows[par][r]) cudaIpcCloseMemHandle(c->peer_rows[par][r]);
  if (c->peer_mode && c->peer_ipc)
    for (int par = 0; par < 2; ++par)
      for (int r = 0; r < c->cfg.nranks && r < 8; ++r)
        if (r != c->cfg.rank && c->peer_pos[par][r]) cudaIpcCloseMemHandle(c->peer_pos[par][r]);
  c->d_sderiv_b.release();
  c->d_pslice[0].release();
  c->d_pslice[1].release();
  c->d_inv.release();
  c->d_cidx.release();
  c->d_fin.release();
  c->d_tilework.release();
  c->d_tilecnt.release();
  c->d_tilestart.release();
  if (c->h_idx) cudaFreeHost(c->h_idx);
  if (c->h_capinfo) cudaFreeHost(c->h_capinfo);
  if (c->h_small) cudaFreeHost(c->h_small);
  if (c->h_u64) cudaFreeHost(c->h_u64);
  for (int i = 0; i < 10; ++i)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 2 * b200coord_ctx::kRing; ++i)
    if (c->sweep_ev[i]) cudaEventDestroy(c->sweep_ev[i]);
  // frames still in flight (b200coord_submit without a collect): their copies touch the caller's buffers
  if (c->st_up) cudaStreamSynchronize(c->st_up);
  if (c->st_down) cudaStreamSynchronize(c->st_down);
  for (int i = 0; i < 2; ++i) {
    c->d_posq[i].release();
    c->d_outq[i].release();
    if (c->q_up[i]) cudaEventDestroy(c->q_up[i]);
    if (c->q_done[i]) cudaEventDestroy(c->q_done[i]);
    if (c->q_down[i]) cudaEventDestroy(c->q_down[i]);
  }
  if (c->h_tailq) cudaFreeHost(c->h_tailq);
  if (c->st_up) cudaStreamDestroy(c->st_up);
  if (c->st_down) cudaStreamDestroy(c->st_down);
  if (c->st) cudaStreamDestroy(c->st);
  delete c;
}

int b200coord_set_box(b200coord_ctx* c, const double box[9]) {
  if (!c || !box) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (c->box_set && std::memcmp(box, c->box_cached, sizeof(c->box_cached)) == 0) return B200COORD_OK;
  std::memcpy(c->box_cached, box, sizeof(c->box_cached));
  c->box_epoch++;
  setup_pbc(box, c->hpbc);
  to_dev_pbc(c->hpbc, c->cfg.pbc != 0, c->dpbc);
  c->params_dirty = true;
  c->box_set = true;
  return B200COORD_OK;
}

int b200coord_prepare(b200coord_ctx* c, long step, int exchange_step, int* will_rebuild) {
  if (!c) return fail(c, B200COORD_ERR_INVALID, "null context");
  int rc = B200COORD_OK;
  const int stride = (c->cfg.nl_mode == B200COORD_NL_NONE) ? 0 : c->cfg.nl_stride;
  if (stride > 0) {  // NeighborList::prepare, NeighborList.cpp:433-456
    if (stride == 1) {
      c->invalidate = true;
      c->firsttime = false;
    } else if (c->firsttime || (step % stride == 0)) {
      c->invalidate = true;
      c->firsttime = false;
    } else {
      c->invalidate = false;
      if (exchange_step)
        rc = fail(c, B200COORD_ERR_STATE,
                  "Neighbor lists should be updated on exchange steps - choose a NL_STRIDE which divides the exchange stride!");
    }
    if (exchange_step) c->firsttime = true;
  }
  if (will_rebuild) *will_rebuild = (stride > 0 && c->invalidate) ? 1 : 0;
  return rc;
}

int b200coord_update_list(b200coord_ctx* c, const double* pos) {
  if (!c || !pos) return fail(c, B200COORD_ERR_INVALID, "null argument");
  CU(c, cudaSetDevice(c->device));
  if (c->cfg.pbc && !c->box_set) return fail(c, B200COORD_ERR_STATE, "b200coord_set_box must be called before update_list");
  CU(c, cudaMemcpyAsync(c->d_pos.p, pos, sizeof(double) * 3 * (size_t)c->n, cudaMemcpyHostToDevice, c->st));
  int rc = rebuild(c, c->d_pos.p);
  if (rc) return rc;
  CU(c, cudaStreamSynchronize(c->st));
  c->invalidate = false;
  return B200COORD_OK;
}

int b200coord_calculate(b200coord_ctx* c, const double* pos, double* value, double* deriv, double* virial) {
  if (!c || !pos || !value || !deriv || !virial) return fail(c, B200COORD_ERR_INVALID, "null argument");
  CU(c, cudaSetDevice(c->device));
  const size_t n3 = 3 * (size_t)c->n;
  maybe_pin(c, 0, pos, sizeof(double) * n3);
  maybe_pin(c, 1, deriv, sizeof(double) * n3);
  CU(c, cudaEventRecord(c->ev[0], c->st));
  CU(c, cudaMemcpyAsync(c->d_pos.p, pos, sizeof(double) * n3, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaEventRecord(c->ev[1], c->st));
  int rc = run_device(c, c->d_pos.p);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->ev[6], c->st));
  CU(c, cudaMemcpyAsync(deriv, c->d_out.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(c->h_small, c->d_out.p + n3, sizeof(double) * 10, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaEventRecord(c->ev[7], c->st));
  CU(c, cudaStreamSynchronize(c->st));
  c->ev_valid[0] = c->ev_valid[3] = true;
  for (int i = 0; i < 9; ++i) virial[i] = c->h_small[i];
  *value = c->h_small[9];
  refresh_stats(c);
  return B200COORD_OK;
}

int b200coord_calculate_device(b200coord_ctx* c, const double* d_pos, double* d_out) {
  if (!c || !d_pos || !d_out) return fail(c, B200COORD_ERR_INVALID, "null argument");
  CU(c, cudaSetDevice(c->device));
  const bool direct = (c->cfg.style != B200COORD_STYLE_PAIR);  // the un-sort writes the caller's buffer itself
  int rc = run_device(c, d_pos, direct ? d_out : nullptr);
  if (rc) return rc;
  if (direct)
    CU(c, cudaMemcpyAsync(d_out + 3 * (size_t)c->n, c->d_out.p + 3 * (size_t)c->n, sizeof(double) * 10, cudaMemcpyDeviceToDevice, c->st));
  else
    CU(c, cudaMemcpyAsync(d_out, c->d_out.p, sizeof(double) * (3 * (size_t)c->n + 10), cudaMemcpyDeviceToDevice, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  refresh_stats(c);
  return B200COORD_OK;
}

int b200coord_enqueue_device(b200coord_ctx* c, const double* d_pos, double* d_out) {
  if (!c || !d_pos || !d_out) return fail(c, B200COORD_ERR_INVALID, "null argument");
  CU(c, cudaSetDevice(c->device));
  const bool direct = (c->cfg.style != B200COORD_STYLE_PAIR);  // the un-sort writes the caller's buffer itself
  int rc = run_device(c, d_pos, direct ? d_out : nullptr);
  if (rc) return rc;
  if (direct)
    CU(c, cudaMemcpyAsync(d_out + 3 * (size_t)c->n, c->d_out.p + 3 * (size_t)c->n, sizeof(double) * 10, cudaMemcpyDeviceToDevice, c->st));
  else
    CU(c, cudaMemcpyAsync(d_out, c->d_out.p, sizeof(double) * (3 * (size_t)c->n + 10), cudaMemcpyDeviceToDevice, c->st));
  return B200COORD_OK;
}

// ---- frames known in advance (trajectory post-processing: plumed driver reads the next frame while this one is
// computed, src/cltools/Driver.cpp).  Two steps are in flight: the upload of frame k+1 and the download of result k-1
// run on their own streams under the sweep of frame k.  Results are the ones b200coord_calculate returns, bit for bit.
namespace {
int finish_pending(b200coord_ctx* c, int slot) {
  b200coord_ctx::Pending& p = c->pend[slot];
  if (!p.busy) return B200COORD_OK;
  CU(c, cudaEventSynchronize(c->q_down[slot]));
  const double* t = c->h_tailq + 10 * slot;
  for (int i = 0; i < 9; ++i) p.virial[i] = t[i];
  *p.value = t[9];
  p.busy = false;
  return B200COORD_OK;
}
}  // namespace

int b200coord_submit(b200coord_ctx* c, const double* pos, double* value, double* deriv, double* virial) {
  if (!c || !pos || !value || !deriv || !virial) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (c->cfg.nranks > 1) return fail(c, B200COORD_ERR_UNSUPPORTED, "b200coord_submit is a single-context call");
  CU(c, cudaSetDevice(c->device));
  const size_t n3 = 3 * (size_t)c->n;
  if (!c->st_up) {
    CU(c, cudaStreamCreateWithFlags(&c->st_up, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&c->st_down, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CU(c, cudaEventCreateWithFlags(&c->q_up[i], cudaEventDisableTiming));
      CU(c, cudaEventCreateWithFlags(&c->q_done[i], cudaEventDisableTiming));
      CU(c, cudaEventCreateWithFlags(&c->q_down[i], cudaEventDisableTiming));
      CU(c, c->d_posq[i].reserve(n3));
      CU(c, c->d_outq[i].reserve(n3 + 10));
    }
    CU(c, cudaHostAlloc((void**)&c->h_tailq, 20 * sizeof(double), cudaHostAllocDefault));
  }
  const int slot = (int)(c->q_next & 1u);
  int rc = finish_pending(c, slot);  // the step that used this slot two submits ago
  if (rc) return rc;
  maybe_pin(c, 0, pos, sizeof(double) * n3);
  CU(c, cudaMemcpyAsync(c->d_posq[slot].p, pos, sizeof(double) * n3, cudaMemcpyHostToDevice, c->st_up));
  CU(c, cudaEventRecord(c->q_up[slot], c->st_up));
  CU(c, cudaStreamWaitEvent(c->st, c->q_up[slot], 0));
  rc = b200coord_enqueue_device(c, c->d_posq[slot].p, c->d_outq[slot].p);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->q_done[slot], c->st));
  CU(c, cudaStreamWaitEvent(c->st_down, c->q_done[slot], 0));
  CU(c, cudaMemcpyAsync(deriv, c->d_outq[slot].p, sizeof(double) * n3, cudaMemcpyDeviceToHost, c->st_down));
  CU(c, cudaMemcpyAsync(c->h_tailq + 10 * slot, c->d_outq[slot].p + n3, sizeof(double) * 10, cudaMemcpyDeviceToHost, c->st_down));
  CU(c, cudaEventRecord(c->q_down[slot], c->st_down));
  c->pend[slot].busy = true;
  c->pend[slot].value = value;
  c->pend[slot].virial = virial;
  c->q_next++;
  return B200COORD_OK;
}

int b200coord_collect(b200coord_ctx* c) {
  if (!c) return fail(c, B200COORD_ERR_INVALID, "null argument");
  CU(c, cudaSetDevice(c->device));
  for (int i = 0; i < 2; ++i) {  // oldest first
    const int rc = finish_pending(c, (int)((c->q_next + (unsigned)i) & 1u));
    if (rc) return rc;
  }
  CU(c, cudaStreamSynchronize(c->st));
  refresh_stats(c);
  return B200COORD_OK;
}

int b200coord_stream_mark(b200coord_ctx* c, int which) {
  if (!c || which < 0 || which > 1) return fail(c, B200COORD_ERR_INVALID, "bad stopwatch mark");
  CU(c, cudaSetDevice(c->device));
  if (which == 0) {
    c->sweep_n = 0;
    c->build_ms_acc = c->build_ms_max = 0.f;
    c->build_n = 0;
  }
  CU(c, cudaEventRecord(c->ev[8 + which], c->st));
  return B200COORD_OK;
}

int b200coord_stream_elapsed_ms(b200coord_ctx* c, float* ms) {
  if (!c || !ms) return fail(c, B200COORD_ERR_INVALID, "null argument");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventSynchronize(c->ev[9]));
  CU(c, cudaEventElapsedTime(ms, c->ev[8], c->ev[9]));
  refresh_stats(c);
  return B200COORD_OK;
}

int b200coord_my_slice(const b200coord_ctx* c, unsigned* slot_begin, unsigned* slot_count) {
  if (!c || !slot_begin || !slot_count) return B200COORD_ERR_INVALID;
  *slot_begin = c->slot_begin;
  *slot_count = c->slot_count;
  return B200COORD_OK;
}

// one barrier over the ranks on the context's stream (a one-element all-reduce)
static int rank_barrier(b200coord_ctx* c) {
  NcclApi& api = nccl_api();
  ncclResult_t r = api.AllReduce(c->d_small.p + 24, c->d_small.p + 24, 1, ncclDouble, ncclSum, c->comm, c->st);
  if (r != ncclSuccess) return fail(c, B200COORD_ERR_NCCL, std::string("ncclAllReduce(barrier): ") + api.GetErrorString(r));
  return collective_done(c);
}

// does the coming step keep its list AND its continuous coordinates?  Then a rank only needs the positions of the
// atoms around its rows, and pulls them from the owners' slice buffers over NVLink instead of receiving everything.
static bool step_can_pull(const b200coord_ctx* c) {
  const bool need_rebuild = !c->list_valid || (c->cfg.nl_mode != B200COORD_NL_NONE && c->invalidate);
  return c->peer_mode && c->comm && !need_rebuild && c->cfg.style != B200COORD_STYLE_PAIR &&
         c->cfg.nl_mode == B200COORD_NL_CLASSIC && c->u_mode && c->box_epoch == c->sort_box_epoch;
}

int b200coord_calculate_distributed(b200coord_ctx* c, const double* pos_slice, double* value, double* deriv_slice,
                                    double* virial) {
  if (!c || !pos_slice || !value || !deriv_slice || !virial) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (c->cfg.nranks > 1 && !c->comm) return fail(c, B200COORD_ERR_STATE, "b200coord_comm_init has not been called");
  CU(c, cudaSetDevice(c->device));
  const size_t off = 3 * (size_t)c->slot_begin, cnt = 3 * (size_t)c->slot_count;
  const bool sliced = (c->cfg.style != B200COORD_STYLE_PAIR);
  maybe_pin(c, 0, pos_slice, sizeof(double) * cnt);
  maybe_pin(c, 1, deriv_slice, sizeof(double) * cnt);
  int rc;
  if (c->test_fail_at >= 0 && c->dist_calls++ == c->test_fail_at)
    return fail(c, B200COORD_ERR_STATE, "failure injected by B200COORD_TEST_FAIL");
  B200_TRACE(c, "calculate_distributed: pull=%d list_valid=%d invalidate=%d", (int)step_can_pull(c), (int)c->list_valid, (int)c->invalidate);
  CU(c, cudaEventRecord(c->ev[0], c->st));
  if (step_can_pull(c)) {
    c->pos_parity ^= 1u;
    if (cnt) CU(c, cudaMemcpyAsync(c->d_pslice[c->pos_parity].p, pos_slice, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->st));
    CU(c, cudaEventRecord(c->ev[1], c->st));
    rc = rank_barrier(c);  // every rank's slice is in place before anybody gathers from it
    B200_TRACE(c, "barrier passed");
    if (rc) return rc;
    PosSrc src;
    for (int r = 0; r < 8; ++r) src.base[r] = c->peer_pos[c->pos_parity][r < c->cfg.nranks ? r : 0];
    src.chunk = c->row_chunk ? c->row_chunk : 1u;
    rc = run_device(c, nullptr, nullptr, c->slot_begin, c->slot_count, &src);
  } else {
    if (cnt) CU(c, cudaMemcpyAsync(c->d_pos.p + off, pos_slice, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->st));
    CU(c, cudaEventRecord(c->ev[1], c->st));
    if (c->comm) {  // positions of all ranks over NVLink (in place: every rank's slice sits at rank*chunk)
      NcclApi& api = nccl_api();
      ncclResult_t r = api.AllGather(c->d_pos.p + (size_t)3 * c->row_chunk * c->cfg.rank, c->d_pos.p, (size_t)3 * c->row_chunk,
                                     ncclDouble, c->comm, c->st);
      if (r != ncclSuccess) return fail(c, B200COORD_ERR_NCCL, std::string("ncclAllGather(positions): ") + api.GetErrorString(r));
      rc = collective_done(c);
      if (rc) return rc;
      B200_TRACE(c, "positions all-gathered");
    }
    rc = run_device(c, c->d_pos.p, nullptr, sliced ? c->slot_begin : 0u, sliced ? c->slot_count : 0xffffffffu);
  }
  if (rc) return rc;
  CU(c, cudaEventRecord(c->ev[6], c->st));
  if (cnt) CU(c, cudaMemcpyAsync(deriv_slice, c->d_out.p + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaMemcpyAsync(c->h_small, c->d_out.p + 3 * (size_t)c->n, sizeof(double) * 10, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaEventRecord(c->ev[7], c->st));
  B200_TRACE(c, "step enqueued, waiting for the stream");
  CU(c, cudaStreamSynchronize(c->st));
  B200_TRACE(c, "step done");
  c->ev_valid[0] = c->ev_valid[3] = true;
  for (int i = 0; i < 9; ++i) virial[i] = c->h_small[i];
  *value = c->h_small[9];
  refresh_stats(c);
  return B200COORD_OK;
}

int b200coord_enqueue_device_distributed(b200coord_ctx* c, const double* d_pos, double* d_out_slice) {
  if (!c || !d_pos || !d_out_slice) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (c->cfg.nranks > 1 && !c->comm) return fail(c, B200COORD_ERR_STATE, "b200coord_comm_init has not been called");
  if (c->cfg.style == B200COORD_STYLE_PAIR) return fail(c, B200COORD_ERR_INVALID, "PAIR style returns whole arrays: use b200coord_enqueue_device");
  CU(c, cudaSetDevice(c->device));
  // the un-sort writes slots [slot_begin, +slot_count) of an array that starts slot_begin slots before the caller's
  double* base = d_out_slice - 3 * (size_t)c->slot_begin;
  int rc = run_device(c, d_pos, base, c->slot_begin, c->slot_count);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(d_out_slice + 3 * (size_t)c->slot_count, c->d_out.p + 3 * (size_t)c->n, sizeof(double) * 10,
                        cudaMemcpyDeviceToDevice, c->st));
  return B200COORD_OK;
}

int b200coord_device_count(int* n) {
  if (!n) return B200COORD_ERR_INVALID;
  *n = 0;
  if (cudaGetDeviceCount(n) != cudaSuccess) {
    cudaGetLastError();
    *n = 0;
    return fail(nullptr, B200COORD_ERR_CUDA, "cudaGetDeviceCount failed");
  }
  return B200COORD_OK;
}

int b200coord_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return B200COORD_ERR_INVALID;
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return fail(nullptr, B200COORD_ERR_CUDA, "cudaSetDevice failed");
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return fail(nullptr, B200COORD_ERR_CUDA, "no CUDA device");
  const double t = measure_dfma_tflops(nullptr, sms, 5);
  if (t <= 0.0) return fail(nullptr, B200COORD_ERR_CUDA, "DFMA microbenchmark failed");
  *tflops = t;
  return B200COORD_OK;
}

int b200coord_get_stats(const b200coord_ctx* cc, b200coord_stats* out) {
  if (!cc || !out) return B200COORD_ERR_INVALID;
  b200coord_ctx* c = const_cast<b200coord_ctx*>(cc);
  float ms;
  if (c->ev_valid[0] && cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->stats.last_h2d_ms = ms;
  if (c->ev_valid[1] && cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]) == cudaSuccess) c->stats.last_sweep_ms = ms;
  if (c->ev_valid[2] && cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess) c->stats.last_build_ms = ms;
  if (c->ev_valid[3] && cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]) == cudaSuccess) c->stats.last_d2h_ms = ms;
  {
    const unsigned m = std::min<unsigned>(c->sweep_n, b200coord_ctx::kRing);
    float acc = 0.f;
    unsigned got = 0;
    for (unsigned i = 0; i < m; ++i)
      if (cudaEventElapsedTime(&ms, c->sweep_ev[2 * i], c->sweep_ev[2 * i + 1]) == cudaSuccess) {
        acc += ms;
        ++got;
      }
    c->stats.sweep_ms_sum = acc;
    c->stats.sweep_count = got;
    c->stats.build_ms_sum = c->build_ms_acc;
    c->stats.build_count = c->build_n;
    c->stats.build_ms_max = c->build_ms_max;
    cudaGetLastError();
  }
  *out = c->stats;
  return B200COORD_OK;
}

int b200coord_nl_pairs(b200coord_ctx* c, unsigned* pairs, unsigned long long capacity, unsigned long long* nout) {
  if (!c || !nout) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (!c->list_valid) return fail(c, B200COORD_ERR_STATE, "no neighbour list has been built yet");
  if (c->cfg.nranks > 1) return fail(c, B200COORD_ERR_STATE, "nl_pairs is only available on a single-rank context");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->st));
  std::vector<std::pair<unsigned, unsigned>> out;
  const unsigned n = c->n, n_a = c->n_a;
  if (c->cfg.style == B200COORD_STYLE_PAIR) {
    std::vector<uint8_t> act;
    if (c->cfg.nl_mode == B200COORD_NL_CLASSIC) {
      act.resize(n_a);
      CU(c, cudaMemcpy(act.data(), c->d_active.p, n_a, cudaMemcpyDeviceToHost));
    }
    for (unsigned k = 0; k < n_a; ++k)
      if (act.empty() || act[k]) out.emplace_back(k, k + n_a);
  } else {
    std::vector<uint32_t> perm(n);
    CU(c, cudaMemcpy(perm.data(), c->d_perm.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    auto emit = [&](unsigned si, unsigned sj) {  // slots -> reference (i0,i1)
      if (c->two_groups) {
        if (si < n_a) out.emplace_back(si, sj);
      } else if (si < sj) {
        out.emplace_back(si, sj);
      }
    };
    if (c->cfg.nl_mode == B200COORD_NL_CLASSIC) {
      std::vector<uint32_t> cnt(n);
      std::vector<unsigned long long> st(n);
      CU(c, cudaMemcpy(cnt.data(), c->d_rowcount.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      CU(c, cudaMemcpy(st.data(), c->d_rowstart.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      std::vector<uint32_t> nbr((size_t)c->nbr_total);
      if (c->nbr_total)
        CU(c, cudaMemcpy(nbr.data(), c->d_nbr.p, (size_t)c->nbr_total * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      std::vector<uint32_t> far(2 * (size_t)n);
      if (n) CU(c, cudaMemcpy(far.data(), c->d_rowfar.p, 2 * (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      const uint32_t mask = c->img_list ? kSuperIndexMask : 0xffffffffu;  // image-mode entries: index | image << 26
      for (unsigned k = 0; k < n; ++k) {
        for (uint32_t e = 0; e < cnt[k]; ++e) emit(perm[k], perm[nbr[st[k] + e] & mask]);
        for (uint32_t e = 0; e < far[n + k]; ++e) emit(perm[k], perm[nbr[st[k] + far[k] + e] & mask]);
      }
      // pairs of one and the same atom are at distance 0 <= cutoff: the reference lists them
      if (c->n_self_pairs) {
        for (unsigned i = 0; i < (c->two_groups ? n_a : n); ++i)
          for (unsigned j = (c->two_groups ? n_a : i + 1); j < n; ++j)
            if (c->abs_host[i] == c->abs_host[j]) out.emplace_back(i, j);
      }
    } else {
      const size_t m = (size_t)(c->two_groups ? 2 : 1) * c->grid.ncell;
      std::vector<uint32_t> cs(m), cn(m), sc(n);
      CU(c, cudaMemcpy(cs.data(), c->d_cstart.p, m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      CU(c, cudaMemcpy(cn.data(), c->d_ccount.p, m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      CU(c, cudaMemcpy(sc.data(), c->d_scell.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      const DevGrid& g = c->grid;
      for (unsigned k = 0; k < n; ++k) {
        const unsigned grp = k < n_a ? 0 : 1, other = c->two_groups ? 1 - grp : 0;
        if (c->two_groups && grp == 1) continue;
        int cc[3], lo[3], hi[3];
        const int cell = (int)sc[k];
        cc[2] = cell / (g.n[0] * g.n[1]);
        const int rem = cell - cc[2] * g.n[0] * g.n[1];
        cc[1] = rem / g.n[0];
        cc[0] = rem - cc[1] * g.n[0];
        for (int a = 0; a < 3; ++a) {
          const int nn = g.n[a];
          int mn = cc[a] + ((nn < 2) ? 0 : -1), mx = cc[a] + ((nn < 3 && g.stencil_pbc) ? 1 : 2);
          if (!g.stencil_pbc) {
            mn = std::max(mn, 0);
            mx = std::min(mx, nn);
          }
          lo[a] = mn;
          hi[a] = mx;
        }
        auto wrap = [](int v, int nn) { return v < 0 ? nn - 1 : v % nn; };
        for (int x = lo[0]; x < hi[0]; ++x)
          for (int y = lo[1]; y < hi[1]; ++y)
            for (int z = lo[2]; z < hi[2]; ++z) {
              const size_t ci = (size_t)other * g.ncell + wrap(x, g.n[0]) + wrap(y, g.n[1]) * g.n[0] +
                                (size_t)wrap(z, g.n[2]) * g.n[0] * g.n[1];
              for (uint32_t e = 0; e < cn[ci]; ++e) {
                const unsigned j = cs[ci] + e;
                if (j != k) emit(perm[k], perm[j]);
              }
            }
      }
    }
  }
  std::sort(out.begin(), out.end());
  *nout = out.size();
  if (pairs) {
    const unsigned long long m = std::min<unsigned long long>(capacity, out.size());
    for (unsigned long long i = 0; i < m; ++i) {
      pairs[2 * i] = out[i].first;
      pairs[2 * i + 1] = out[i].second;
    }
  }
  return B200COORD_OK;
}

int b200coord_nl_pairs_device(b200coord_ctx* c, unsigned* d_pairs, unsigned long long capacity, unsigned long long* npairs) {
  if (!c || !npairs) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (!c->list_valid) return fail(c, B200COORD_ERR_STATE, "no neighbour list has been built yet");
  if (c->cfg.nl_mode != B200COORD_NL_CLASSIC || c->cfg.style == B200COORD_STYLE_PAIR)
    return fail(c, B200COORD_ERR_UNSUPPORTED, "nl_pairs_device hands out the distance-filtered list (NLIST, not PAIR)");
  CU(c, cudaSetDevice(c->device));
  const unsigned rows = c->row_end - c->row_begin;
  DevBuf<uint32_t> cnt;
  DevBuf<unsigned long long> start;
  CU(c, cnt.reserve(rows + 1));
  CU(c, start.reserve(rows + 2));
  CU(c, c->d_bsum.reserve(rows / 1024 + 4));
  const uint32_t mask = c->img_list ? kSuperIndexMask : 0xffffffffu;
  launch_export_pairs(false, rows, c->row_begin, c->n_a, c->two_groups, c->d_perm.p, c->d_rowstart.p, c->d_rowcount.p,
                      c->d_rowfar.p, c->d_rowfar.p + c->far_rows, c->d_nbr.p, mask, cnt.p, nullptr, nullptr, 0ull, c->st);
  launch_scan_rows(cnt.p, rows, 0u, c->d_bsum.p, start.p, start.p + rows, c->st);
  unsigned long long total = 0;
  CU(c, cudaMemcpyAsync(&total, start.p + rows, sizeof(total), cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  *npairs = total;
  if (d_pairs && capacity) {
    launch_export_pairs(true, rows, c->row_begin, c->n_a, c->two_groups, c->d_perm.p, c->d_rowstart.p, c->d_rowcount.p,
                        c->d_rowfar.p, c->d_rowfar.p + c->far_rows, c->d_nbr.p, mask, cnt.p, start.p, d_pairs, capacity, c->st);
    CU(c, cudaStreamSynchronize(c->st));
  }
  c->stats.kernel_launches += 5;
  CU_LAST(c, "nl_pairs_device");
  cnt.release();
  start.release();
  return B200COORD_OK;
}

int b200coord_comm_unique_id(char id[B200COORD_UNIQUE_ID_BYTES]) {
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(nullptr, B200COORD_ERR_NCCL, "libnccl.so.2 could not be loaded");
  static_assert(sizeof(ncclUniqueId) == B200COORD_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  ncclResult_t r = api.GetUniqueId(&u);
  if (r != ncclSuccess) return fail(nullptr, B200COORD_ERR_NCCL, std::string("ncclGetUniqueId: ") + api.GetErrorString(r));
  std::memcpy(id, &u, sizeof(u));
  return B200COORD_OK;
}

int b200coord_comm_init(b200coord_ctx* c, const char id[B200COORD_UNIQUE_ID_BYTES]) {
  if (!c || !id) return fail(c, B200COORD_ERR_INVALID, "null argument");
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(c, B200COORD_ERR_NCCL, "libnccl.so.2 could not be loaded");
  if (c->cfg.nranks <= 1) return B200COORD_OK;
  CU(c, cudaSetDevice(c->device));
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  ncclResult_t r = api.CommInitRank(&c->comm, c->cfg.nranks, u, c->cfg.rank);
  if (r != ncclSuccess) {
    c->comm = nullptr;
    return fail(c, B200COORD_ERR_NCCL, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
  }
  return B200COORD_OK;
}

int b200coord_peer_export(b200coord_ctx* c, char handle[B200COORD_PEER_HANDLE_BYTES]) {
  if (!c || !handle) return fail(c, B200COORD_ERR_INVALID, "null argument");
  static_assert(4 * sizeof(cudaIpcMemHandle_t) == B200COORD_PEER_HANDLE_BYTES, "IPC handle size");
  if (c->cfg.style == B200COORD_STYLE_PAIR) return fail(c, B200COORD_ERR_INVALID, "PAIR style has no row exchange");
  CU(c, cudaSetDevice(c->device));
  const size_t padded = (size_t)3 * c->row_chunk * (size_t)c->cfg.nranks;
  CU(c, c->d_sderiv_b.reserve(padded));
  CU(c, cudaMemsetAsync(c->d_sderiv_b.p, 0, sizeof(double) * padded, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  // the two slice buffers of this rank's positions (3 * row_chunk doubles each)
  for (int par = 0; par < 2; ++par) {
    CU(c, c->d_pslice[par].reserve((size_t)3 * c->row_chunk + 3));
    CU(c, cudaMemsetAsync(c->d_pslice[par].p, 0, sizeof(double) * ((size_t)3 * c->row_chunk + 3), c->st));
  }
  CU(c, cudaStreamSynchronize(c->st));
  cudaIpcMemHandle_t h[4];
  CU(c, cudaIpcGetMemHandle(&h[0], c->d_sderiv.p));
  CU(c, cudaIpcGetMemHandle(&h[1], c->d_sderiv_b.p));
  CU(c, cudaIpcGetMemHandle(&h[2], c->d_pslice[0].p));
  CU(c, cudaIpcGetMemHandle(&h[3], c->d_pslice[1].p));
  std::memcpy(handle, h, sizeof(h));
  return B200COORD_OK;
}

int b200coord_peer_attach(b200coord_ctx* c, const char* all) {
  if (!c || !all) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (!c->comm) return fail(c, B200COORD_ERR_STATE, "b200coord_comm_init must come first");
  if (c->cfg.nranks > 8) return fail(c, B200COORD_ERR_INVALID, "peer exchange supports up to 8 ranks");
  if (!c->d_sderiv_b.p) return fail(c, B200COORD_ERR_STATE, "b200coord_peer_export must come first");
  CU(c, cudaSetDevice(c->device));
  for (int r = 0; r < c->cfg.nranks; ++r) {
    if (r == c->cfg.rank) {
      c->peer_rows[0][r] = c->d_sderiv.p;
      c->peer_rows[1][r] = c->d_sderiv_b.p;
      c->peer_pos[0][r] = c->d_pslice[0].p;
      c->peer_pos[1][r] = c->d_pslice[1].p;
      continue;
    }
    cudaIpcMemHandle_t h[4];
    std::memcpy(h, all + (size_t)r * B200COORD_PEER_HANDLE_BYTES, sizeof(h));
    for (int par = 0; par < 2; ++par) {
      void* ptr = nullptr;
      CU(c, cudaIpcOpenMemHandle(&ptr, h[par], cudaIpcMemLazyEnablePeerAccess));
      c->peer_rows[par][r] = static_cast<double*>(ptr);
      ptr = nullptr;
      CU(c, cudaIpcOpenMemHandle(&ptr, h[2 + par], cudaIpcMemLazyEnablePeerAccess));
      c->peer_pos[par][r] = static_cast<double*>(ptr);
    }
  }
  c->peer_mode = true;
  c->peer_ipc = true;
  c->parity = 0;
  return B200COORD_OK;
}

int b200coord_peer_attach_local(b200coord_ctx* c, b200coord_ctx* const* all, int n) {
  if (!c || !all) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (!c->comm) return fail(c, B200COORD_ERR_STATE, "b200coord_comm_init must come first");
  if (n != c->cfg.nranks || n > 8) return fail(c, B200COORD_ERR_INVALID, "peer exchange supports up to 8 ranks");
  CU(c, cudaSetDevice(c->device));
  for (int r = 0; r < n; ++r) {
    b200coord_ctx* o = all[r];
    if (!o || !o->d_sderiv_b.p || !o->d_pslice[0].p) return fail(c, B200COORD_ERR_STATE, "b200coord_peer_export must come first on every context");
    if (r != c->cfg.rank && o->device != c->device) {
      int can = 0;
      CU(c, cudaDeviceCanAccessPeer(&can, c->device, o->device));
      if (!can) return fail(c, B200COORD_ERR_CUDA, "the devices of this group cannot access each other's memory");
      cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(c, B200COORD_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
    c->peer_rows[0][r] = o->d_sderiv.p;
    c->peer_rows[1][r] = o->d_sderiv_b.p;
    c->peer_pos[0][r] = o->d_pslice[0].p;
    c->peer_pos[1][r] = o->d_pslice[1].p;
  }
  c->peer_mode = true;
  c->peer_ipc = false;
  c->parity = 0;
  return B200COORD_OK;
}

// ================================================================================================
// Several GPUs inside ONE process: one context per device, one worker thread per context (NCCL collectives need every
// rank to enqueue concurrently), i-atoms sharded over them like MPI ranks (CoordinationBase.cpp:152-170).
struct b200coord_group {
  std::vector<b200coord_ctx*> ctx;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::function<int(int)> task;
  unsigned long epoch = 0;
  int pending = 0;
  bool quit = false;
  std::vector<int> rc;
  std::string err;
  unsigned n = 0;
  bool broken = false;  // a member failed: its peers' communicators were aborted, the group only reports the error
  int broken_rc = 0;
};

namespace {
void group_worker(b200coord_group* g, int r) {
  unsigned long seen = 0;
  for (;;) {
    std::function<int(int)> fn;
    {
      std::unique_lock<std::mutex> lk(g->mu);
      g->cv_go.wait(lk, [&] { return g->quit || g->epoch != seen; });
      if (g->quit) return;
      seen = g->epoch;
      fn = g->task;
    }
    const int rc = fn(r);
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->rc[r] = rc;
      if (rc && !g->broken && g->ctx.size() > 1) {
        // The peers of a member that fails wait for it in their next collective, forever.  Abort their communicators
        // (ncclCommAbort ends the operations in flight) so that every worker comes back; the group is unusable after.
        g->broken = true;
        g->broken_rc = rc;
        g->err = "device " + std::to_string(g->ctx[(size_t)r] ? g->ctx[(size_t)r]->device : -1) + " (rank " + std::to_string(r) +
                 "): " + (g->ctx[(size_t)r] ? g->ctx[(size_t)r]->err : g_last_error);
        NcclApi& api = nccl_api();
        if (api.CommAbort)
          for (size_t o = 0; o < g->ctx.size(); ++o) {
            b200coord_ctx* x = g->ctx[o];
            if (x && x->comm && !x->comm_aborted) {
              x->comm_aborted = true;
              api.CommAbort(x->comm);
            }
          }
      }
      if (--g->pending == 0) g->cv_done.notify_all();
    }
  }
}
// run fn(rank) on every worker, wait for all; first non-zero code wins
int group_run(b200coord_group* g, std::function<int(int)> fn) {
  {
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->broken) {
      g_last_error = g->err;
      return g->broken_rc ? g->broken_rc : B200COORD_ERR_STATE;
    }
    g->task = std::move(fn);
    g->pending = (int)g->ctx.size();
    g->epoch++;
  }
  g->cv_go.notify_all();
  std::unique_lock<std::mutex> lk(g->mu);
  g->cv_done.wait(lk, [&] { return g->pending == 0; });
  if (g->broken) {
    g_last_error = g->err;
    return g->broken_rc;
  }
  for (size_t r = 0; r < g->ctx.size(); ++r)
    if (g->rc[r]) {
      g->err = "rank " + std::to_string(r) + ": " + (g->ctx[r] ? g->ctx[r]->err : g_last_error);
      return g->rc[r];
    }
  return B200COORD_OK;
}
}  // namespace

int b200coord_group_create(const b200coord_config* cfg, const b200coord_switch* sw, const unsigned* abs_index,
                           const int* devices, int ndevices, b200coord_group** out) {
  if (!cfg || !sw || !devices || !out) return fail(nullptr, B200COORD_ERR_INVALID, "null argument");
  *out = nullptr;
  if (ndevices < 1 || ndevices > 8) return fail(nullptr, B200COORD_ERR_INVALID, "a group holds 1 to 8 devices");
  if (cfg->style == B200COORD_STYLE_PAIR && ndevices > 1)
    return fail(nullptr, B200COORD_ERR_INVALID, "PAIR style runs on one device");
  auto g = std::make_unique<b200coord_group>();
  g->ctx.assign((size_t)ndevices, nullptr);
  g->rc.assign((size_t)ndevices, 0);
  g->n = cfg->n_group_a + cfg->n_group_b;
  for (int r = 0; r < ndevices; ++r) {
    b200coord_config c = *cfg;
    c.device = devices[r];
    c.rank = r;
    c.nranks = ndevices;
    const int rc = b200coord_create(&c, sw, abs_index, &g->ctx[(size_t)r]);
    if (rc) {
      for (auto* x : g->ctx) b200coord_destroy(x);
      return rc;
    }
    g->ctx[(size_t)r]->in_process = (ndevices > 1);
  }
  for (int r = 0; r < ndevices; ++r) g->workers.emplace_back(group_worker, g.get(), r);
  b200coord_group* gp = g.release();
  if (ndevices > 1) {
    char id[B200COORD_UNIQUE_ID_BYTES];
    int rc = b200coord_comm_unique_id(id);
    if (!rc) rc = group_run(gp, [&](int r) { return b200coord_comm_init(gp->ctx[(size_t)r], id); });
    char scratch[8][B200COORD_PEER_HANDLE_BYTES];
    if (!rc) rc = group_run(gp, [&](int r) { return b200coord_peer_export(gp->ctx[(size_t)r], scratch[r]); });
    if (!rc) rc = group_run(gp, [&](int r) { return b200coord_peer_attach_local(gp->ctx[(size_t)r], gp->ctx.data(), (int)gp->ctx.size()); });
    if (rc) {
      g_last_error = gp->err;
      b200coord_group_destroy(gp);
      return rc;
    }
  }
  *out = gp;
  return B200COORD_OK;
}

void b200coord_group_destroy(b200coord_group* g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->quit = true;
  }
  g->cv_go.notify_all();
  for (auto& t : g->workers) t.join();
  for (auto* c : g->ctx)
    if (c && c->st) {
      cudaSetDevice(c->device);
      cudaStreamSynchronize(c->st);
    }
  for (auto* c : g->ctx) b200coord_destroy(c);
  delete g;
}

int b200coord_group_size(const b200coord_group* g) { return g ? (int)g->ctx.size() : 0; }
b200coord_ctx* b200coord_group_context(b200coord_group* g, int rank) {
  return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[(size_t)rank] : nullptr;
}
const char* b200coord_group_last_error(const b200coord_group* g) { return g ? g->err.c_str() : g_last_error.c_str(); }

int b200coord_group_set_box(b200coord_group* g, const double box[9]) {
  if (!g) return B200COORD_ERR_INVALID;
  for (auto* c : g->ctx) {
    const int rc = b200coord_set_box(c, box);
    if (rc) { g->err = c->err; return rc; }
  }
  return B200COORD_OK;
}
int b200coord_group_prepare(b200coord_group* g, long step, int exchange_step, int* will_rebuild) {
  if (!g) return B200COORD_ERR_INVALID;
  int rc = B200COORD_OK;
  for (auto* c : g->ctx) {
    const int r = b200coord_prepare(c, step, exchange_step, will_rebuild);
    if (r) { g->err = c->err; rc = r; }
  }
  return rc;
}
int b200coord_group_set_charges(b200coord_group* g, const double* charges) {
  if (!g) return B200COORD_ERR_INVALID;
  return group_run(g, [&](int r) { return b200coord_set_charges(g->ctx[(size_t)r], charges); });
}
int b200coord_group_set_types(b200coord_group* g, const unsigned* types, unsigned ntypes, const double* etas) {
  if (!g) return B200COORD_ERR_INVALID;
  return group_run(g, [&](int r) { return b200coord_set_types(g->ctx[(size_t)r], types, ntypes, etas); });
}
int b200coord_group_calculate(b200coord_group* g, const double* pos, double* value, double* deriv, double* virial) {
  if (!g || !pos || !value || !deriv || !virial) return B200COORD_ERR_INVALID;
  if (g->ctx.size() == 1) return b200coord_calculate(g->ctx[0], pos, value, deriv, virial);
  double vals[8] = {0}, virs[8][9];
  const int rc = group_run(g, [&](int r) {
    b200coord_ctx* c = g->ctx[(size_t)r];
    return b200coord_calculate_distributed(c, pos + 3 * (size_t)c->slot_begin, &vals[r], deriv + 3 * (size_t)c->slot_begin, virs[r]);
  });
  if (rc) return rc;
  *value = vals[0];  // every rank holds the all-reduced value and virial
  for (int i = 0; i < 9; ++i) virial[i] = virs[0][i];
  return B200COORD_OK;
}

// ---- coupling with an MD engine that keeps positions and forces on the device. The reference passes
// host pointers through plumed_cmd (patches/gromacs-2025.0.diff/.../plumedforceprovider.cpp:171-204); here the engine
// publishes its device arrays under a name, the action that carries GPU_COUPLING=<name> looks them up, reads the
// positions where they are and adds force-on-CV x derivative to the engine's force array (Colvar::apply,
// src/core/Colvar.cpp:50-60) without either array visiting the host.
namespace {
struct CouplingEntry {
  int device = 0;
  const double* d_pos = nullptr;
  double* d_force = nullptr;
  size_t natoms = 0;
};
std::mutex g_coupling_mutex;
std::map<std::string, CouplingEntry>& coupling_table() {
  static std::map<std::string, CouplingEntry> t;
  return t;
}
}  // namespace

int b200coord_coupling_publish(const char* key, int device, const double* d_pos, double* d_force, size_t natoms) {
  if (!key || !*key || !d_pos || !d_force || natoms == 0) return fail(nullptr, B200COORD_ERR_INVALID, "coupling_publish: null or empty argument");
  cudaPointerAttributes at;
  for (const void* q : {static_cast<const void*>(d_pos), static_cast<const void*>(d_force)}) {
    if (cudaPointerGetAttributes(&at, q) != cudaSuccess || at.type != cudaMemoryTypeDevice || at.device != device) {
      cudaGetLastError();
      return fail(nullptr, B200COORD_ERR_INVALID, "coupling_publish: not a device pointer of the stated device");
    }
  }
  std::lock_guard<std::mutex> lk(g_coupling_mutex);
  CouplingEntry& e = coupling_table()[key];
  e.device = device;
  e.d_pos = d_pos;
  e.d_force = d_force;
  e.natoms = natoms;
  return B200COORD_OK;
}

int b200coord_coupling_withdraw(const char* key) {
  if (!key) return fail(nullptr, B200COORD_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(g_coupling_mutex);
  return coupling_table().erase(key) ? B200COORD_OK : fail(nullptr, B200COORD_ERR_STATE, std::string("no coupling named ") + key);
}

int b200coord_coupling_lookup(const char* key, int* device, const double** d_pos, double** d_force, size_t* natoms) {
  if (!key) return fail(nullptr, B200COORD_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(g_coupling_mutex);
  auto it = coupling_table().find(key);
  if (it == coupling_table().end()) return fail(nullptr, B200COORD_ERR_STATE, std::string("no coupling named ") + key);
  if (device) *device = it->second.device;
  if (d_pos) *d_pos = it->second.d_pos;
  if (d_force) *d_force = it->second.d_force;
  if (natoms) *natoms = it->second.natoms;
  return B200COORD_OK;
}

int b200coord_coupled_set_index(b200coord_ctx* c, const unsigned* index) {
  if (!c) return fail(c, B200COORD_ERR_INVALID, "null argument");
  CU(c, cudaSetDevice(c->device));
  c->coupled_fresh = false;
  if (!index) {
    c->cidx_set = false;
    return B200COORD_OK;
  }
  CU(c, c->d_cidx.reserve(c->n));
  CU(c, cudaMemcpyAsync(c->d_cidx.p, index, sizeof(unsigned) * (size_t)c->n, cudaMemcpyHostToDevice, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  c->cidx_set = true;
  return B200COORD_OK;
}

int b200coord_calculate_coupled(b200coord_ctx* c, const double* d_pos_all, double* value, double* virial) {
  if (!c || !d_pos_all || !value || !virial) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (c->cfg.nranks > 1) return fail(c, B200COORD_ERR_UNSUPPORTED, "device coupling is one context on the engine's device");
  CU(c, cudaSetDevice(c->device));
  c->coupled_fresh = false;
  const double* pos = d_pos_all;
  if (c->cidx_set) {
    launch_coupled_gather(d_pos_all, c->d_cidx.p, c->n, c->d_pos.p, c->st);
    pos = c->d_pos.p;
  }
  int rc = run_device(c, pos);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->h_small, c->d_out.p + 3 * (size_t)c->n, sizeof(double) * 10, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  for (int i = 0; i < 9; ++i) virial[i] = c->h_small[i];
  *value = c->h_small[9];
  refresh_stats(c);
  c->coupled_fresh = true;
  return B200COORD_OK;
}

int b200coord_apply_coupled(b200coord_ctx* c, double factor, double* d_force_all) {
  if (!c || !d_force_all) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (!c->coupled_fresh) return fail(c, B200COORD_ERR_STATE, "apply_coupled needs the derivatives of a b200coord_calculate_coupled");
  CU(c, cudaSetDevice(c->device));
  launch_coupled_apply(c->d_out.p, c->cidx_set ? c->d_cidx.p : nullptr, c->n, factor, d_force_all, c->st);
  CU_LAST(c, "coupled apply");
  CU(c, cudaStreamSynchronize(c->st));  // the engine may read its forces as soon as this returns
  return B200COORD_OK;
}

int b200coord_coupled_derivatives(b200coord_ctx* c, double* deriv) {
  if (!c || !deriv) return fail(c, B200COORD_ERR_INVALID, "null argument");
  if (!c->coupled_fresh) return fail(c, B200COORD_ERR_STATE, "no derivatives of a coupled step on the device");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(deriv, c->d_out.p, sizeof(double) * 3 * (size_t)c->n, cudaMemcpyDeviceToHost, c->st));
  CU(c, cudaStreamSynchronize(c->st));
  return B200COORD_OK;
}

int b200coord_host_alloc(size_t bytes, void** ptr) {
  if (!ptr) return B200COORD_ERR_INVALID;
  cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(nullptr, B200COORD_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return B200COORD_OK;
}
int b200coord_host_free(void* ptr) {
  if (ptr && cudaFreeHost(ptr) != cudaSuccess) return B200COORD_ERR_CUDA;
  return B200COORD_OK;
}
int b200coord_device_alloc(size_t bytes, void** dptr) {
  if (!dptr) return B200COORD_ERR_INVALID;
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
  if (e != cudaSuccess) return fail(nullptr, B200COORD_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  return B200COORD_OK;
}
int b200coord_device_free(void* dptr) {
  if (dptr && cudaFree(dptr) != cudaSuccess) return B200COORD_ERR_CUDA;
  return B200COORD_OK;
}
int b200coord_memcpy_h2d(void* dptr, const void* hptr, size_t bytes) {
  cudaError_t e = cudaMemcpy(dptr, hptr, bytes, cudaMemcpyHostToDevice);
  return e == cudaSuccess ? B200COORD_OK : fail(nullptr, B200COORD_ERR_CUDA, cudaGetErrorString(e));
}
int b200coord_memcpy_d2h(void* hptr, const void* dptr, size_t bytes) {
  cudaError_t e = cudaMemcpy(hptr, dptr, bytes, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? B200COORD_OK : fail(nullptr, B200COORD_ERR_CUDA, cudaGetErrorString(e));
}
int b200coord_device_synchronize(void) {
  cudaError_t e = cudaDeviceSynchronize();
  return e == cudaSuccess ? B200COORD_OK : fail(nullptr, B200COORD_ERR_CUDA, cudaGetErrorString(e));
}

}  // extern "C"
