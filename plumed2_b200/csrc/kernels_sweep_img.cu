// FP64 instances of the image-mode list sweep (sweep_img.cuh), one translation unit of its own so that it
// compiles in parallel with the general sweeps.
#include <algorithm>

#include "sweep_img.cuh"

namespace b200 {

template <int MINB>
static int run_img_kind(const SweepArgs& a, const DevSwitch& sw, const ImgShifts& sh, const NearBands& nb, cudaStream_t st) {
  switch (kind_of(sw.type)) {
    case K_FIX6: return run_sweep_img<K_FIX6, MINB>(a, sw, sh, nb, st);
    case K_FIXN: return run_sweep_img<K_FIXN, MINB>(a, sw, sh, nb, st);
    case K_RAT_R2: return run_sweep_img<K_RAT_R2, MINB>(a, sw, sh, nb, st);
    case K_RAT_R: return run_sweep_img<K_RAT_R, MINB>(a, sw, sh, nb, st);
    case K_EXP: return run_sweep_img<K_EXP, MINB>(a, sw, sh, nb, st);
    case K_GAUSS: return run_sweep_img<K_GAUSS, MINB>(a, sw, sh, nb, st);
    case K_FASTGAUSS: return run_sweep_img<K_FASTGAUSS, MINB>(a, sw, sh, nb, st);
    case K_SMAP: return run_sweep_img<K_SMAP, MINB>(a, sw, sh, nb, st);
    case K_CUBIC: return run_sweep_img<K_CUBIC, MINB>(a, sw, sh, nb, st);
    case K_TANH: return run_sweep_img<K_TANH, MINB>(a, sw, sh, nb, st);
    case K_COS: return run_sweep_img<K_COS, MINB>(a, sw, sh, nb, st);
    case K_NATIVEQ: return run_sweep_img<K_NATIVEQ, MINB>(a, sw, sh, nb, st);
    case K_DH: return run_sweep_img<K_DH, MINB>(a, sw, sh, nb, st);
    case K_GHB: return run_sweep_img<K_GHB, MINB>(a, sw, sh, nb, st);
    default: return -1;
  }
}

// high-word buckets around D_MAX^2 and D_0^2 in which a pair is handed to the exact patch (which re-checks with the
// 1e-10 band of on_boundary)
static NearBands make_bands(const DevSwitch& sw) {
  NearBands nb;
  auto hi = [](double v) -> uint32_t {
    unsigned long long b;
    static_assert(sizeof(b) == sizeof(v), "double size");
    __builtin_memcpy(&b, &v, sizeof(b));
    return (uint32_t)(b >> 32);
  };
  auto band = [&](int q, double centre, double half) {
    if (half >= 0.0 && centre > 0.0 && centre - half > 0.0) {
      nb.lo[q] = hi(centre - half);
      nb.span[q] = hi(centre + half) - nb.lo[q];
    } else {  // no such boundary: h - 0xffffffff <= 0 only for the NaN pattern 0xffffffff
      nb.lo[q] = 0xffffffffu;
      nb.span[q] = 0u;
    }
  };
  band(0, sw.dmax_2, sw.band_dmax);
  band(1, sw.d0_2, sw.band_d0);
  return nb;
}

unsigned sweep_img_rows_per_block(unsigned rows_a, unsigned rows_b, unsigned max_row) {
  // one shape for the GROUPA and the GROUPB launch: the smaller of the two picks
  unsigned rpb = img_rows_per_block(rows_a ? rows_a : rows_b, max_row);
  if (rows_a && rows_b) rpb = std::min(rpb, img_rows_per_block(rows_b, max_row));
  return rpb;
}

int launch_sweep_img(const SweepArgs& a, const DevPbc& box, const DevSwitch& sw, int variant, cudaStream_t st) {
  ImgShifts sh;
  for (int code = 0; code < 64; ++code) {
    const unsigned c = (unsigned)code ^ kImageCentre;
    const int w[3] = {(int)(c & 3u) - 1, (int)((c >> 2) & 3u) - 1, (int)((c >> 4) & 3u) - 1};
    for (int k = 0; k < 3; ++k) sh.v[code][k] = w[0] * box.box[k] + w[1] * box.box[3 + k] + w[2] * box.box[6 + k];
  }
  const NearBands nb = make_bands(sw);
  if (variant == 3) return run_img_kind<3>(a, sw, sh, nb, st);
  return run_img_kind<2>(a, sw, sh, nb, st);
}

}  // namespace b200
