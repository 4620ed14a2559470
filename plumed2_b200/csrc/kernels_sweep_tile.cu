// Instances and launch wrappers of the cell-tile sweep (sweep_tile.cuh): NLISTCELLS and "no list", FP64, every
// switching function except the two typed pairings (DHENERGY / GHBFIX keep the row kernel).
#include <algorithm>

#include "sweep_tile.cuh"

namespace b200 {

unsigned tile_rows_per_item(unsigned rows) {
  unsigned r = rows / (148u * 4u);  // small systems: enough items for every SM
  if (r < 4u) r = 4u;
  if (r > (unsigned)kTileRows) r = (unsigned)kTileRows;
  return r;
}

// upper bound of the number of items of `rows` rows spread over at most `nseg` segments
unsigned tile_items_bound(unsigned rows, unsigned nseg, unsigned rows_per_item) {
  return std::min(rows, nseg) + rows / rows_per_item + 1u;
}

void launch_tile_work(unsigned nseg, const uint32_t* cstart, const uint32_t* ccount, unsigned row_begin, unsigned row_end,
                      unsigned rows_per_item, uint32_t* nitems /*nseg*/, unsigned long long* item_start /*nseg + 1*/,
                      unsigned long long* bsum, TileWork* work, cudaStream_t st) {
  if (!nseg) return;
  k_tile_count<<<(nseg + 255) / 256, 256, 0, st>>>(nseg, cstart, ccount, row_begin, row_end, rows_per_item, nitems);
  launch_scan_rows(nitems, nseg, 0u, bsum, item_start, item_start + nseg, st);  // [nseg] = total
  k_tile_fill<<<(nseg + 255) / 256, 256, 0, st>>>(nseg, cstart, ccount, row_begin, row_end, rows_per_item, item_start, work);
}

template <int K, int PBC, bool ACC>
static void launch_one(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, const TileWork* work,
                       const unsigned long long* first_dev, const unsigned long long* end_dev, unsigned blocks, cudaStream_t st) {
  if (!blocks) return;
  constexpr size_t kDyn = (size_t)kTileCap * sizeof(SPos);
  // the attribute belongs to the function ON ONE DEVICE: several contexts of one process (b200coord_group_*) each
  // need it (a per-process flag made the second device's launch fail with "invalid argument")
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaFuncSetAttribute(k_sweep_tile<K, PBC, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDyn);
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  k_sweep_tile<K, PBC, ACC><<<blocks, kSweepThreads, kDyn, st>>>(a, pbc, sw, work, first_dev, end_dev);
}

template <int K, int PBC>
static int run_tile(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, const TileWork* work,
                    const unsigned long long* zero_dev, const unsigned long long* split_dev, const unsigned long long* total_dev,
                    unsigned bound_a, unsigned bound_b, cudaStream_t st) {
  // GROUPA items (all items of a single group) accumulate value + virial: one partial record per block
  launch_one<K, PBC, true>(a, pbc, sw, work, zero_dev, a.two_groups ? split_dev : total_dev, bound_a, st);
  // scatter_b: the GROUPA rows have added +dd to their partners; the virial is complete (df d (x) d per pair)
  if (a.two_groups && !a.scatter_b) launch_one<K, PBC, false>(a, pbc, sw, work, split_dev, total_dev, bound_b, st);
  return (int)bound_a;
}

template <int K>
static int run_tile_pbc(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, const TileWork* work,
                        const unsigned long long* zero_dev, const unsigned long long* split_dev,
                        const unsigned long long* total_dev, unsigned bound_a, unsigned bound_b, cudaStream_t st) {
  switch (pbc.type) {
    case 0: return run_tile<K, 0>(a, pbc, sw, work, zero_dev, split_dev, total_dev, bound_a, bound_b, st);
    case 1: return run_tile<K, 1>(a, pbc, sw, work, zero_dev, split_dev, total_dev, bound_a, bound_b, st);
    default: return run_tile<K, 2>(a, pbc, sw, work, zero_dev, split_dev, total_dev, bound_a, bound_b, st);
  }
}

bool sweep_tile_supports(int sw_type) {
  const int k = kind_of(sw_type);
  return k >= 0 && k != K_DH && k != K_GHB;
}

int launch_sweep_tile(const SweepArgs& a, const DevPbc& pbc, const DevSwitch& sw, const TileWork* work,
                      const unsigned long long* zero_dev, const unsigned long long* split_dev, const unsigned long long* total_dev,
                      unsigned bound_a, unsigned bound_b, cudaStream_t st) {
#define B200_TILE_CASE(KK) \
  case KK: return run_tile_pbc<KK>(a, pbc, sw, work, zero_dev, split_dev, total_dev, bound_a, bound_b, st);
  switch (kind_of(sw.type)) {
    B200_TILE_CASE(K_FIX6)
    B200_TILE_CASE(K_FIXN)
    B200_TILE_CASE(K_RAT_R2)
    B200_TILE_CASE(K_RAT_R)
    B200_TILE_CASE(K_EXP)
    B200_TILE_CASE(K_GAUSS)
    B200_TILE_CASE(K_FASTGAUSS)
    B200_TILE_CASE(K_SMAP)
    B200_TILE_CASE(K_CUBIC)
    B200_TILE_CASE(K_TANH)
    B200_TILE_CASE(K_COS)
    B200_TILE_CASE(K_NATIVEQ)
    default: return -1;
  }
#undef B200_TILE_CASE
}

}  // namespace b200
