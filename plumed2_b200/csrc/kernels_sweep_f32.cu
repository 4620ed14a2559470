// FP32 instances of the list and cell sweeps (B200COORD_FP32: FP64 minimum image, FP32 pair arithmetic and row
// sums, FP64 accumulation across rows; see sweep_math.cuh).  The PAIR style keeps its FP64 kernel: one pair per
// atom is not worth a second flavour.
#include "sweep_kernels.cuh"

namespace b200 {

int launch_sweep_list_f32(const SweepArgs& a, const DevPbc& pbc, const DevSwitchT<float>& sw, cudaStream_t st) {
  return run_sweep_kind<true, float>(a, pbc, sw, st);
}
int launch_sweep_cells_f32(const SweepArgs& a, const DevPbc& pbc, const DevSwitchT<float>& sw, cudaStream_t st) {
  return run_sweep_kind<false, float>(a, pbc, sw, st);
}

}  // namespace b200
