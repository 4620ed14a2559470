// Host-side set-up for libb200coord (see host_setup.hpp).  Compiled with -ffp-contract=off: the box
// inverse, the reduced lattice and the shift vectors feed the bit-exact neighbour-list test on the
// device, so they are computed with the same operation order as the reference's Tensor/Pbc code.
#include "host_setup.hpp"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <utility>
#include <vector>

namespace b200 {

namespace {

struct V3 {
  double x[3];
  double& operator[](int i) { return x[i]; }
  double operator[](int i) const { return x[i]; }
};

inline double norm2(const V3& v) { return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]; }           // LoopUnroller.h:146
inline double dot(const V3& a, const V3& b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }  // LoopUnroller.h:153

double det3(const double* d) {  // Tensor.h:384-392
  return d[0] * d[4] * d[8] + d[1] * d[5] * d[6] + d[2] * d[3] * d[7] - d[0] * d[5] * d[7] - d[1] * d[3] * d[8] -
         d[2] * d[4] * d[6];
}

void invert3(const double* m, double* out) {  // Tensor.h:416-425
  const double invdet = 1.0 / det3(m);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      out[3 * j + i] = invdet * (m[3 * i1 + j1] * m[3 * i2 + j2] - m[3 * i1 + j2] * m[3 * i2 + j1]);
    }
}

void sort_by_length(V3 v[3]) {  // LatticeReduction::sort :32-58
  double m[3] = {norm2(v[0]), norm2(v[1]), norm2(v[2])};
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (m[i] > m[j]) {
        std::swap(v[i], v[j]);
        std::swap(m[i], m[j]);
      }
}

void gauss_reduce(V3& a, V3& b) {  // LatticeReduction::reduce(a,b) :60-86
  const double tol = 1.0 + 1e-14;
  double ma = norm2(a), mb = norm2(b);
  for (unsigned it = 0; it < 1000000u; ++it) {
    if (mb > ma) {
      std::swap(a, b);
      std::swap(ma, mb);
    }
    const double k = std::floor(dot(a, b) / mb + 0.5);
    for (int c = 0; c < 3; ++c) a[c] -= b[c] * k;
    ma = norm2(a);
    if (mb <= ma * tol) break;
  }
  std::swap(a, b);
}

}  // namespace

void reduce_lattice(double rows[9]) {  // LatticeReduction::reduceFast :144-192
  const double tol = 1.0 + 1e-14;
  V3 v[3];
  std::memcpy(v, rows, sizeof(v));
  for (unsigned it = 0; it < 1000000u; ++it) {
    sort_by_length(v);
    gauss_reduce(v[0], v[1]);
    const double b11 = norm2(v[0]), b22 = norm2(v[1]);
    const double b12 = dot(v[0], v[1]), b13 = dot(v[0], v[2]), b23 = dot(v[1], v[2]);
    const double z = b11 * b22 - b12 * b12;
    const double y2 = -(b11 * b23 - b12 * b13) / z;
    const double y1 = -(b22 * b13 - b12 * b23) / z;
    const int x1lo = (int)std::floor(y1), x2lo = (int)std::floor(y2);
    bool have = false;
    double mbest = 0;
    V3 best{};
    for (int x1 = x1lo; x1 <= x1lo + 1; ++x1)
      for (int x2 = x2lo; x2 <= x2lo + 1; ++x2) {
        V3 t;
        for (int c = 0; c < 3; ++c) t[c] = (v[2][c] + x2 * v[1][c]) + x1 * v[0][c];
        const double mt = norm2(t);
        if (!have || mt < mbest) {
          mbest = mt;
          best = t;
          have = true;
        }
      }
    if (norm2(best) * tol >= norm2(v[2])) break;
    v[2] = best;
  }
  sort_by_length(v);
  std::memcpy(rows, v, sizeof(v));
}

void setup_pbc(const double box[9], HostPbc& p) {  // Pbc::setBox :165-212
  p = HostPbc();
  std::memcpy(p.box, box, sizeof(p.box));
  const double tiny = 1e-28;
  const double det = det3(box);
  if (det * det < tiny) return;  // type stays "unset"
  const bool cxy = box[1] * box[1] < tiny && box[3] * box[3] < tiny;
  const bool cxz = box[2] * box[2] < tiny && box[6] * box[6] < tiny;
  const bool cyz = box[5] * box[5] < tiny && box[7] * box[7] < tiny;
  invert3(p.box, p.inv_box);
  std::memcpy(p.reduced, p.box, sizeof(p.reduced));
  if (cxy && cxz && cyz) {
    p.type = 1;
    invert3(p.reduced, p.inv_reduced);
    return;
  }
  p.type = 2;
  reduce_lattice(p.reduced);
  invert3(p.reduced, p.inv_reduced);

  // Pbc::buildShifts :59-135 -- candidate lattice translations per octant of the scaled vector
  const double* R = p.reduced;
  double G[9];  // R * R^T, k innermost (Tensor.h:428-437)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += R[3 * i + k] * R[3 * j + k];
      G[3 * i + j] = s;
    }
  for (int l = -1; l <= 1; ++l)
    for (int m = -1; m <= 1; ++m)
      for (int n = -1; n <= 1; ++n) {
        const int is[3] = {l, m, n};
        const int nonzero = (l != 0) + (m != 0) + (n != 0);
        if (nonzero == 0 || nonzero == 3) continue;
        const V3 ds{{(double)l, (double)m, (double)n}};
        V3 cosdir;
        for (int i = 0; i < 3; ++i) {
          double s = 0.0;
          for (int j = 0; j < 3; ++j) s += G[3 * i + j] * ds[j];
          cosdir[i] = s;
        }
        const double dp = dot(ds, cosdir);
        const double ref = norm2(ds) * norm2(cosdir);
        if (std::fabs(ref - dp * dp) < tiny) continue;  // face perpendicular to its axis: never needed
        V3 shift;  // transpose(reduced) * ds
        for (int i = 0; i < 3; ++i) {
          double s = 0.0;
          for (int j = 0; j < 3; ++j) s += R[3 * j + i] * ds[j];
          shift[i] = s;
        }
        for (int oct = 0; oct < 8; ++oct) {
          const int block[3] = {2 * ((oct >> 2) & 1) - 1, 2 * ((oct >> 1) & 1) - 1, 2 * (oct & 1) - 1};
          bool too_far = false;
          for (int s = 0; s < 3; ++s) too_far |= (is[s] * block[s] > 0);
          if (too_far) continue;
          bool useful = false;
          for (int s = 0; s < 3; ++s)
            if (((1 - is[s] * is[s]) * block[s]) * cosdir[s] < -tiny) useful = true;
          if (!useful) continue;
          if (p.nshift[oct] < kMaxShift) {
            for (int c = 0; c < 3; ++c) p.shifts[oct][p.nshift[oct]][c] = shift[c];
            p.nshift[oct]++;
          }
        }
      }
}

void cell_grid(const double inv_box[9], double cutoff, unsigned ncells[3]) {  // LinkCells::createCells :99-122
  for (int k = 0; k < 3; ++k) {
    const V3 col{{inv_box[k], inv_box[3 + k], inv_box[6 + k]}};  // row k of transpose(invBox)
    const double v = std::floor(1.0 / std::sqrt(norm2(col)) / cutoff);
    ncells[k] = v >= 1.0 ? (unsigned)v : 1u;
  }
}

// ------------------------------------------------------------------------------------------------
// switching functions
namespace {

double ipow(double base, int e) {  // Tools::fastpow, Tools.h:581-595
  if (e < 0) {
    e = -e;
    base = 1.0 / base;
  }
  double r = 1.0;
  while (e) {
    if (e & 1) r *= base;
    e >>= 1;
    base *= base;
  }
  return r;
}

void init_data(b200coord_switch& s, int type, double d0, double dmax, double r0) {  // Data::init :84-94
  std::memset(&s, 0, sizeof(s));
  s.type = type;
  s.stretch = 1.0;
  s.nn = 6;
  s.mm = 12;
  s.nnf = 3;
  s.mmf = 6;
  s.beta = 50.0;
  s.lambda = 1.8;
  s.d0 = d0;
  s.dmax = dmax;
  s.dmax_2 = (dmax < std::sqrt(DBL_MAX)) ? dmax * dmax : DBL_MAX;
  s.invr0 = 1.0 / r0;
  s.invr0_2 = s.invr0 * s.invr0;
}

int fixed_power(int type) {
  switch (type) {
    case B200COORD_SW_RATIONALFIX12: return 12;
    case B200COORD_SW_RATIONALFIX10: return 10;
    case B200COORD_SW_RATIONALFIX8: return 8;
    case B200COORD_SW_RATIONALFIX6: return 6;
    case B200COORD_SW_RATIONALFIX4: return 4;
    case B200COORD_SW_RATIONALFIX2: return 2;
    default: return 0;
  }
}

void make_rational(b200coord_switch& s, double d0, double dmax, double r0, int N, int M) {  // rationalFactory :306-347
  const bool even = (N % 2 == 0) && (M % 2 == 0) && (d0 == 0.0);
  const bool n2m = (2 * N == M) || (M == 0);
  if (n2m && even && N <= 12 && N >= 2) {
    static const int kinds[7] = {-1, B200COORD_SW_RATIONALFIX2, B200COORD_SW_RATIONALFIX4, B200COORD_SW_RATIONALFIX6,
                                 B200COORD_SW_RATIONALFIX8, B200COORD_SW_RATIONALFIX10, B200COORD_SW_RATIONALFIX12};
    init_data(s, kinds[N / 2], d0, dmax, r0);
    return;
  }
  init_data(s, B200COORD_SW_RATIONAL, d0, dmax, r0);  // rational<>::init :230-256
  s.nn = N;
  s.mm = (M == 0) ? 2 * N : M;
  const double n = s.nn, m = s.mm;
  s.preRes = n / m;
  s.preDfunc = 0.5 * n * (n - m) / m;
  s.preSecDev = (n * (m * m - 3.0 * m * (-1 + n) + n * (-3 + 2 * n))) / (6.0 * m);
  s.nnf = s.nn / 2;
  s.mmf = s.mm / 2;
  const double nf = s.nnf, mf = s.mmf;
  s.preDfuncF = 0.5 * nf * (nf - mf) / mf;
  s.preSecDevF = (nf * (mf * mf - 3.0 * mf * (-1 + nf) + nf * (-3 + 2 * nf))) / (6.0 * mf);
  if (n2m)
    s.type = even ? B200COORD_SW_RATIONALSIMPLEFAST : B200COORD_SW_RATIONALSIMPLE;
  else
    s.type = even ? B200COORD_SW_RATIONALFAST : B200COORD_SW_RATIONAL;
}

// f(x) only (no derivative) of the reduced distance x>0; SwitchingFunction.cpp:185-507
double shape(const b200coord_switch& s, double x) {
  const int fp = fixed_power(s.type);
  if (fp) return 1.0 / (1.0 + ipow(x, fp - 1) * x);
  switch (s.type) {
    case B200COORD_SW_RATIONALSIMPLE:
    case B200COORD_SW_RATIONALSIMPLEFAST:
      return 1.0 / (1.0 + ipow(x, s.nn - 1) * x);
    case B200COORD_SW_RATIONAL:
    case B200COORD_SW_RATIONALFAST: {
      const double lo = 1.0 - 5.0e10 * DBL_EPSILON, hi = 1.0 + 5.0e10 * DBL_EPSILON;
      if (!(x > lo && x < hi)) {
        const double num = 1.0 - ipow(x, s.nn - 1) * x;
        const double iden = 1.0 / (1.0 - ipow(x, s.mm - 1) * x);
        return num * iden;
      }
      const double dx = x - 1.0;
      return s.preRes + dx * (s.preDfunc + 0.5 * dx * s.preSecDev);
    }
    case B200COORD_SW_EXPONENTIAL: return std::exp(-x);
    case B200COORD_SW_GAUSSIAN:
    case B200COORD_SW_FASTGAUSSIAN: return std::exp(-0.5 * x * x);
    case B200COORD_SW_SMAP: return std::pow(1.0 + s.c * ipow(x, s.a), s.d);
    case B200COORD_SW_CUBIC: {
      const double t1 = x - 1.0, t2 = 1.0 + 2.0 * x;
      return t1 * t1 * t2;
    }
    case B200COORD_SW_TANH: return 1.0 - std::tanh(x);
    case B200COORD_SW_COSINUS: return x <= 1.0 ? 0.5 * (std::cos(x * M_PI) + 1.0) : 0.0;
    default: return 0.0;
  }
}

void setup_stretch(b200coord_switch& s) {  // SwitchInterface::setupStretch :63-72
  if (s.dmax < DBL_MAX) {
    s.stretch = 1.0;
    s.shift = 0.0;
    const double s0 = switch_value_host(s, 0.0);
    const double sd = switch_value_host(s, s.dmax);
    s.stretch = 1.0 / (s0 - sd);
    s.shift = -sd * s.stretch;
  }
}

// "KEY=value" extraction over a word list (stands in for Tools::parse on the SWITCH string)
struct Words {
  std::vector<std::string> w;
  int find(const std::string& key) const {
    for (size_t i = 0; i < w.size(); ++i)
      if (w[i].compare(0, key.size() + 1, key + "=") == 0) return (int)i;
    return -1;
  }
  // 1 = parsed, 0 = absent, -1 = malformed
  int number(const std::string& key, double& v) {
    const int i = find(key);
    if (i < 0) return 0;
    const std::string txt = w[i].substr(key.size() + 1);
    w.erase(w.begin() + i);
    char* end = nullptr;
    const double x = std::strtod(txt.c_str(), &end);
    if (end == txt.c_str() || *end != 0) return -1;
    v = x;
    return 1;
  }
  int integer(const std::string& key, int& v) {
    double x;
    const int r = number(key, x);
    if (r == 1) {
      if (x != std::floor(x)) return -1;
      v = (int)x;
    }
    return r;
  }
  bool flag(const std::string& key) {
    for (size_t i = 0; i < w.size(); ++i)
      if (w[i] == key) {
        w.erase(w.begin() + i);
        return true;
      }
    return false;
  }
};

}  // namespace

double switch_value_host(const b200coord_switch& s, double r) {  // baseSwitch::calculate :135-149 (value only)
  if (s.type == B200COORD_SW_NATIVEQ) {                         // nativeqSwitch::calculate :524-549
    if (r > s.dmax) return 0.0;
    double res = 1.0;
    if (r > s.d0) res = 1.0 / (1.0 + std::exp(s.beta * (r - s.lambda * s.ref)));
    return res * s.stretch + s.shift;
  }
  if (r > s.dmax) return 0.0;
  const double x = (r - s.d0) * s.invr0;
  if (x > 0.0) return shape(s, x) * s.stretch + s.shift;
  return s.stretch + s.shift;
}

void ghbfix_pairing(double dmax, double d0, double c, b200coord_switch& out) {
  std::memset(&out, 0, sizeof(out));
  out.type = B200COORD_PAIR_GHBFIX;
  out.d0 = d0;
  out.dmax = dmax;
  out.dmax_2 = dmax * dmax;  // dmax_squared, :99
  const double dmax2 = dmax - d0;  // :106
  out.preRes = (-c * dmax2 * dmax2) / ((1 - c) * dmax2 * dmax2);  // A :108
  out.preDfunc = (2 * dmax2) / ((1 - c) * dmax2 * dmax2);         // B :109
  out.preSecDev = -1 / ((1 - c) * dmax2 * dmax2);                 // C :110
  out.d = 1 / (c * dmax2 * dmax2);                                // D :111
  out.c = c * dmax2;                                              // joint of the two pieces, :206
}

void dhenergy_pairing(double I, double T, double epsilon, double energy_unit, double length_unit, double charge_unit,
                      b200coord_switch& out) {
  std::memset(&out, 0, sizeof(out));
  out.type = B200COORD_PAIR_DHENERGY;
  out.d0 = 0.0;
  out.dmax = DBL_MAX;  // no cutoff: every listed pair contributes
  out.dmax_2 = DBL_MAX;
  const double constant = 138.935458111 / energy_unit / length_unit * charge_unit * charge_unit;  // :119
  out.beta = std::sqrt(I / (epsilon * T)) * 502.903741125 * length_unit;                           // :120 (k)
  out.lambda = constant / epsilon;
}

void rational_switch(int nn, int mm, double r0, double d0, b200coord_switch& out) {
  if (mm == 0) mm = 2 * nn;
  const double dmax = d0 + r0 * std::pow(0.00001, 1. / (nn - mm));
  make_rational(out, d0, dmax, r0, nn, mm);
  setup_stretch(out);
}

int parse_switch(const std::string& definition, b200coord_switch& out, std::string& err) {
  err.clear();
  Words ws;
  {
    std::string clean = definition;
    for (char& c : clean)
      if (c == '{' || c == '}') c = ' ';
    std::istringstream is(clean);
    std::string tok;
    while (is >> tok) ws.w.push_back(tok);
  }
  out.type = B200COORD_SW_NOT_INITIALIZED;
  if (ws.w.empty()) {
    err = "missing all input for switching function";
    return B200COORD_ERR_PARSE;
  }
  const std::string name = ws.w[0];
  ws.w.erase(ws.w.begin());
  double d0 = 0.0, dmax = DBL_MAX;
  if (ws.number("D_0", d0) < 0) err = "could not parse D_0";
  if (ws.number("D_MAX", dmax) < 0) err = "could not parse D_MAX";
  ws.flag("STRETCH");
  const bool stretch = !ws.flag("NOSTRETCH");
  if (name == "CUBIC") {
    init_data(out, B200COORD_SW_CUBIC, d0, dmax, dmax - d0);
  } else {
    double r0 = 0.0;
    if (ws.number("R_0", r0) != 1) err = "R_0 is required for " + name;
    if (name == "RATIONAL") {
      int nn = 6, mm = 0;
      if (ws.integer("NN", nn) < 0) err = "could not parse NN";
      if (ws.integer("MM", mm) < 0) err = "could not parse MM";
      make_rational(out, d0, dmax, r0, nn, mm);
    } else if (name == "SMAP") {
      int a = 0, b = 0;
      if (ws.integer("A", a) != 1) err = "A is required for " + name;
      if (ws.integer("B", b) != 1) err = "B is required for " + name;
      init_data(out, B200COORD_SW_SMAP, d0, dmax, r0);
      out.a = a;
      out.b = b;
      if (a != 0 && b != 0) {
        out.c = std::pow(2., (double)a / (double)b) - 1.0;
        out.d = -(double)b / (double)a;
      }
    } else if (name == "Q") {
      double beta = 50.0, lambda = 1.8, ref = 0.0;
      if (ws.number("BETA", beta) < 0) err = "could not parse BETA";
      if (ws.number("LAMBDA", lambda) < 0) err = "could not parse LAMBDA";
      if (ws.number("REF", ref) != 1) err = "REF is required for " + name;
      init_data(out, B200COORD_SW_NATIVEQ, d0, dmax, r0);
      out.beta = beta;
      out.lambda = lambda;
      out.ref = ref;
    } else if (name == "EXP") {
      init_data(out, B200COORD_SW_EXPONENTIAL, d0, dmax, r0);
    } else if (name == "GAUSSIAN") {
      if (r0 == 1.0 && d0 == 0.0)
        init_data(out, B200COORD_SW_FASTGAUSSIAN, 0.0, dmax, 1.0);
      else
        init_data(out, B200COORD_SW_GAUSSIAN, d0, dmax, r0);
    } else if (name == "TANH") {
      init_data(out, B200COORD_SW_TANH, d0, dmax, r0);
    } else if (name == "COSINUS") {
      init_data(out, B200COORD_SW_COSINUS, d0, dmax, r0);
    } else if (name == "CUSTOM" || name == "MATHEVAL") {
      err = "SWITCH=" + name + " (lepton expression) cannot be evaluated on the GPU; use the built-in COORDINATION for it";
      out.type = B200COORD_SW_LEPTON;
      return B200COORD_ERR_UNSUPPORTED;
    } else {
      err = "cannot understand switching function type '" + name + "'";
      return B200COORD_ERR_PARSE;
    }
  }
  if (!ws.w.empty()) {
    err = "found the following rogue keywords in switching function input : ";
    for (const auto& k : ws.w) err += k + " ";
  }
  if (!err.empty()) return B200COORD_ERR_PARSE;
  if (stretch && dmax != DBL_MAX) setup_stretch(out);
  return B200COORD_OK;
}

std::string describe_switch(const b200coord_switch& s) {
  static const char* names[] = {"rational", "rational", "rational", "rational", "rational", "rational",
                                "rational", "rational", "rational", "rational", "exponential", "gaussian",
                                "fastgaussian", "smap", "cubic", "tanh", "cosinus", "nativeq", "lepton", "unset"};
  std::ostringstream os;
  if (s.type == B200COORD_PAIR_GHBFIX) {
    os << "GHBFIX pairing: d0=" << s.d0 << " dmax=" << s.dmax << " joint at d0+" << s.c;
    return os.str();
  }
  if (s.type == B200COORD_PAIR_DHENERGY) {
    os << "Debye-Hueckel pairing: screening length " << (s.beta > 0.0 ? 1.0 / s.beta : INFINITY) << ", constant/epsilon " << s.lambda;
    return os.str();
  }
  const int t = (s.type >= 0 && s.type <= B200COORD_SW_NOT_INITIALIZED) ? s.type : B200COORD_SW_NOT_INITIALIZED;
  os << 1.0 / s.invr0 << ".  Using " << names[t] << " switching function with parameters d0=" << s.d0;
  const int fp = fixed_power(s.type);
  if (fp)
    os << " nn=" << fp << " mm=" << 2 * fp;
  else if (t >= B200COORD_SW_RATIONAL && t <= B200COORD_SW_RATIONALSIMPLEFAST)
    os << " nn=" << s.nn << " mm=" << s.mm;
  else if (t == B200COORD_SW_SMAP)
    os << " a=" << s.a << " b=" << s.b;
  else if (t == B200COORD_SW_NATIVEQ)
    os << " beta=" << s.beta << " lambda=" << s.lambda << " ref=" << s.ref;
  if (s.dmax < DBL_MAX) os << " dmax=" << s.dmax << " stretch=" << s.stretch << " shift=" << s.shift;
  return os.str();
}

}  // namespace b200
