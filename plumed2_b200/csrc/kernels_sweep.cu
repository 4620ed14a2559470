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
#include <cstdlib>
#include <type_traits>

#include "sweep_kernels.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// PAIR style: pair k = (k, k+n_a) (NeighborList.cpp:150-152); each atom slot occurs in exactly one pair
template <int K, int PBC>
__global__ void __launch_bounds__(kSweepThreads)
    k_sweep_pairs(const double* __restrict__ pos, const double* __restrict__ charges, const uint32_t* __restrict__ types,
                  const double* __restrict__ etas, unsigned ntypes, const uint32_t* __restrict__ abs_index, const uint8_t* __restrict__ active,
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
      double dx = pj.x - pos[ia], dy = pj.y - pos[ia + 1], dz = pj.z - pos[ia + 2];
      min_image_fast<PBC>(pbc, dx, dy, dz);
      double s, df;
      const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
      eval_switch<K>(sw, r2, s, df);
      if (K == K_DH || K == K_GHB) {
        const double qq = (K == K_DH) ? charges[k] * charges[k + n_a] : etas[types[k] * ntypes + types[k + n_a]];
        s *= qq;
        df *= qq;
      }
      if (on_boundary(sw, r2)) {  // one pair per thread: the exact evaluation is simply done in place
        const ExactPair o = exact_pair<K>(pbc, sw, pos[ia], pos[ia + 1], pos[ia + 2], pj.x, pj.y, pj.z, false);
        dx = o.dx;
        dy = o.dy;
        dz = o.dz;
        s = o.s;
        df = o.df;
      }
      const double gx = df * dx, gy = df * dy, gz = df * dz;
      fx = -gx;
      fy = -gy;
      fz = -gz;
      acc.val += s;
      acc.vxx = gx * dx;
      acc.vxy = gx * dy;
      acc.vxz = gx * dz;
      acc.vyy = gy * dy;
      acc.vyz = gy * dz;
      acc.vzz = gz * dz;
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
// fixed-order sum of the block partials; virial = -weight * sum(c), value = weight * sum(s).
// Level 1: block b adds the records [b * 256, (b + 1) * 256) -- 64 groups x 16 components, group g takes records g, g + 64,
// ... and the 64 group sums are added in index order.  Level 2 (one block) adds the level-1 sums in index order.  The
// result does not depend on scheduling.
__global__ void __launch_bounds__(1024) k_finalize1(const double* __restrict__ partials, int nblocks, double* __restrict__ sums) {
  __shared__ double sm[64][16];
  const int comp = threadIdx.x & 15, grp = threadIdx.x >> 4;
  const int b0 = blockIdx.x * 256, b1 = min(nblocks, b0 + 256);
  double t = 0.0;
  if (comp < 10)
    for (int b = b0 + grp; b < b1; b += 64) t += partials[(size_t)b * kPartialStride + comp];
  sm[grp][comp] = t;
  __syncthreads();
  if (threadIdx.x < 10) {
    double r = 0.0;
    for (int g2 = 0; g2 < 64; ++g2) r += sm[g2][threadIdx.x];
    sums[(size_t)blockIdx.x * 16 + threadIdx.x] = r;
  }
}
__global__ void __launch_bounds__(32) k_finalize2(const double* __restrict__ sums, int n, double weight, double* __restrict__ tail) {
  if (threadIdx.x < 10) {
    double r = 0.0;
    for (int b = 0; b < n; ++b) r += sums[(size_t)b * 16 + threadIdx.x];
    if (threadIdx.x == 0) tail[9] = r * weight;
    else tail[threadIdx.x - 1] = -(r * weight);
  }
}

// Derivatives back in the caller's order: slot s <- row inv[s].  With several ranks the row lives in the buffer of the
// rank that swept it (rows.base[owner], peer memory over NVLink): every rank pulls exactly the rows of the atoms it
// returns, 24 bytes each, instead of every rank receiving every row.
__global__ void __launch_bounds__(256)
    k_unsort_pull(RowSrc rows, const uint32_t* __restrict__ inv, double* __restrict__ out, unsigned slot_lo, unsigned slot_cnt) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= slot_cnt) return;
  const unsigned slot = slot_lo + t;
  const uint32_t k = inv[slot];
  const double* __restrict__ src = rows.base[k / rows.chunk] + 3 * (size_t)k;
  const size_t o = 3 * (size_t)slot;
  out[o] = src[0];
  out[o + 1] = src[1];
  out[o + 2] = src[2];
}

// ------------------------------------------------------------------------------------------------
// Coupling with an MD engine whose arrays live on the device (capi.cu: b200coord_calculate_coupled / _apply_coupled):
// the action's atoms are picked out of the engine's position array, and the chain rule of Colvar::apply
// (src/core/Colvar.cpp:50-60: force on atom = force on the CV x derivative) is added to the engine's force array.
__global__ void __launch_bounds__(256)
    k_coupled_gather(const double* __restrict__ pos_all, const uint32_t* __restrict__ index, unsigned n, double* __restrict__ pos) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double* __restrict__ src = pos_all + 3 * (size_t)index[t];
  const size_t o = 3 * (size_t)t;
  pos[o] = src[0];
  pos[o + 1] = src[1];
  pos[o + 2] = src[2];
}

__global__ void __launch_bounds__(256)
    k_coupled_apply(const double* __restrict__ deriv, const uint32_t* __restrict__ index, unsigned n, double factor,
                    double* __restrict__ force_all) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  // an atom listed in both groups owns two derivative rows: atomics keep the two adds apart
  double* dst = force_all + 3 * (size_t)(index ? index[t] : t);
  const size_t o = 3 * (size_t)t;
  atomicAdd(dst, factor * deriv[o]);
  atomicAdd(dst + 1, factor * deriv[o + 1]);
  atomicAdd(dst + 2, factor * deriv[o + 2]);
}

// ------------------------------------------------------------------------------------------------
// dispatch (FP64 here; the FP32 instances live in kernels_sweep_f32.cu)
int launch_sweep_list_f32(const SweepArgs& a, const DevPbc& pbc, const DevSwitchT<float>& sw, cudaStream_t st);
int launch_sweep_cells_f32(const SweepArgs& a, const DevPbc& pbc, const DevSwitchT<float>& sw, cudaStream_t st);

int launch_sweep_list(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st) {
  if (a.f32) return launch_sweep_list_f32(a, pbc, to_f32(sw), st);
  return run_sweep_kind<true, double>(a, pbc, sw, st);
}
int launch_sweep_cells(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, cudaStream_t st) {
  if (a.f32) return launch_sweep_cells_f32(a, pbc, to_f32(sw), st);
  return run_sweep_kind<false, double>(a, pbc, sw, st);
}

template <int K>
static int run_pairs(const double* pos, const double* charges, const uint32_t* types, const double* etas, unsigned ntypes,
                     const uint32_t* abs_index, const uint8_t* active, unsigned n_a, unsigned pb,
                     unsigned pe, const DevPbc& pbc, const DevSwitch& sw, double* out, double* partials,
                     unsigned long long* evals, cudaStream_t st) {
  const int nblocks = (int)((pe - pb + kSweepThreads - 1) / kSweepThreads);
  if (nblocks == 0) return 0;
  switch (pbc.type) {
    case 0: k_sweep_pairs<K, 0><<<nblocks, kSweepThreads, 0, st>>>(pos, charges, types, etas, ntypes, abs_index, active, n_a, pb, pe, pbc, sw, out, partials, evals); break;
    case 1: k_sweep_pairs<K, 1><<<nblocks, kSweepThreads, 0, st>>>(pos, charges, types, etas, ntypes, abs_index, active, n_a, pb, pe, pbc, sw, out, partials, evals); break;
    default: k_sweep_pairs<K, 2><<<nblocks, kSweepThreads, 0, st>>>(pos, charges, types, etas, ntypes, abs_index, active, n_a, pb, pe, pbc, sw, out, partials, evals); break;
  }
  return nblocks;
}

int launch_sweep_pairs(const double* pos, const double* charges, const uint32_t* types, const double* etas, unsigned ntypes,
                       const uint32_t* abs_index, const uint8_t* active, unsigned n_a,
                       unsigned pair_begin, unsigned pair_end, const DevPbc& pbc, const DevSwitch& sw, double* out,
                       double* partials, unsigned long long* evals, cudaStream_t st) {
#define B200_PAIR_CASE(KK) \
  case KK: return run_pairs<KK>(pos, charges, types, etas, ntypes, abs_index, active, n_a, pair_begin, pair_end, pbc, sw, out, partials, evals, st);
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
    B200_PAIR_CASE(K_DH)
    B200_PAIR_CASE(K_GHB)
    default: return -1;
  }
#undef B200_PAIR_CASE
}

void launch_finalize(const double* partials, int nblocks, double weight, double* out_tail, double* scratch, cudaStream_t st) {
  const int n1 = nblocks > 0 ? (nblocks + 255) / 256 : 1;
  k_finalize1<<<n1, 1024, 0, st>>>(partials, nblocks, scratch);
  k_finalize2<<<1, 32, 0, st>>>(scratch, n1, weight, out_tail);
}

void launch_unsort_pull(const RowSrc& rows, const uint32_t* inv, double* out, unsigned slot_lo, unsigned slot_cnt,
                        cudaStream_t st) {
  if (slot_cnt) k_unsort_pull<<<(slot_cnt + 255) / 256, 256, 0, st>>>(rows, inv, out, slot_lo, slot_cnt);
}

void launch_coupled_gather(const double* pos_all, const uint32_t* index, unsigned n, double* pos, cudaStream_t st) {
  if (n) k_coupled_gather<<<(n + 255) / 256, 256, 0, st>>>(pos_all, index, n, pos);
}

void launch_coupled_apply(const double* deriv, const uint32_t* index, unsigned n, double factor, double* force_all,
                          cudaStream_t st) {
  if (n) k_coupled_apply<<<(n + 255) / 256, 256, 0, st>>>(deriv, index, n, factor, force_all);
}

}  // namespace b200
