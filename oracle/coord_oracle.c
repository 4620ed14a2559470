/* TEST INFRASTRUCTURE ONLY -- see coord_oracle.h.  Plain-C restatement of the reference's CPU
 * COORDINATION path.  Compile with -ffp-contract=off (oracle/Makefile does) so that the operation
 * order written here is the operation order executed, as in the reference's non-FMA x86-64 build.
 *
 * All "file:line" citations are relative to /root/reference/src/.
 */
#include "coord_oracle.h"
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.141592653589793238462643383279502884197169399375105820974944592307
#define ORC_EPS DBL_EPSILON /* tools/Tools.h:55 */

/* ------------------------------------------------------------------ small vector helpers */
/* tools/LoopUnroller.h:146-152 : modulo2 = (x*x + y*y) + z*z */
static inline double mod2(const double v[3]) { return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]; }
/* tools/LoopUnroller.h:153-160 : dot = (a0*b0 + a1*b1) + a2*b2 */
static inline double dot3(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
/* tools/Tensor.h:451-458 : row-vector x matrix, accumulation from 0 in j order */
static inline void vecmat(const double a[3], const double m[9], double out[3]) {
  for (int i = 0; i < 3; i++) {
    double t = 0.0;
    for (int j = 0; j < 3; j++) t += a[j] * m[3 * j + i];
    out[i] = t;
  }
}
/* tools/Tensor.h:440-447 : matrix x column-vector */
static inline void matvec(const double m[9], const double b[3], double out[3]) {
  for (int i = 0; i < 3; i++) {
    double t = 0.0;
    for (int j = 0; j < 3; j++) t += m[3 * i + j] * b[j];
    out[i] = t;
  }
}
/* tools/Tensor.h:384-392 */
static double det3(const double d[9]) {
  return d[0] * d[4] * d[8] + d[1] * d[5] * d[6] + d[2] * d[3] * d[7] - d[0] * d[5] * d[7] - d[1] * d[3] * d[8] -
         d[2] * d[4] * d[6];
}
/* tools/Tensor.h:416-425 */
static void inv3(const double m[9], double t[9]) {
  double invdet = 1.0 / det3(m);
  for (unsigned i = 0; i < 3; i++)
    for (unsigned j = 0; j < 3; j++)
      t[3 * j + i] = invdet * (m[3 * ((i + 1) % 3) + (j + 1) % 3] * m[3 * ((i + 2) % 3) + (j + 2) % 3] -
                               m[3 * ((i + 1) % 3) + (j + 2) % 3] * m[3 * ((i + 2) % 3) + (j + 1) % 3]);
}
static void transpose3(const double m[9], double t[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = m[3 * j + i];
}
/* tools/Tensor.h:428-437 : t(i,j) += a(i,k)*b(k,j), k innermost */
static void matmat(const double a[9], const double b[9], double t[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += a[3 * i + k] * b[3 * k + j];
      t[3 * i + j] = s;
    }
}

/* ------------------------------------------------------------------ Tools */
/* tools/Tools.h:545-571 (round_toward_zero branch: int() truncates) */
double orc_tools_pbc(double x) {
  x += 100.0;
  if (x >= 0) return x - (int)(x + 0.5);
  return x - (int)(x - 0.5);
}

/* tools/Tools.h:581-595 (runtime) and :597-617 (template) -- same multiplication sequence */
static double fastpow(double base, int e) {
  if (e < 0) {
    e = -e;
    base = 1.0 / base;
  }
  double result = 1.0;
  while (e) {
    if (e & 1) result *= base;
    e >>= 1;
    base *= base;
  }
  return result;
}

/* ------------------------------------------------------------------ switching functions */
/* Data::init, tools/SwitchingFunction.cpp:84-94 ; defaults tools/SwitchingFunction.h:58-91 */
static void data_init(orc_switch* s, int type, double D0, double DMAX, double R0) {
  memset(s, 0, sizeof(*s));
  s->type = type;
  s->stretch = 1.0;
  s->shift = 0.0;
  s->nn = 6;
  s->mm = 12;
  s->nnf = 3;
  s->mmf = 6;
  s->beta = 50.0;
  s->lambda = 1.8;
  s->d0 = D0;
  s->dmax = DMAX;
  s->dmax_2 = (DMAX < sqrt(DBL_MAX)) ? DMAX * DMAX : DBL_MAX;
  s->invr0 = 1.0 / R0;
  s->invr0_2 = s->invr0 * s->invr0;
}

/* rationalFactory, tools/SwitchingFunction.cpp:306-347 + rational<>::init :230-256 */
static void rational_factory(orc_switch* s, double D0, double DMAX, double R0, int N, int M) {
  int fast = (N % 2 == 0) && (M % 2 == 0) && (D0 == 0.0);
  if (((2 * N) == M || M == 0) && fast && N <= 12) {
    int t = -1;
    switch (N) {
      case 12: t = ORC_SW_RATIONALFIX12; break;
      case 10: t = ORC_SW_RATIONALFIX10; break;
      case 8: t = ORC_SW_RATIONALFIX8; break;
      case 6: t = ORC_SW_RATIONALFIX6; break;
      case 4: t = ORC_SW_RATIONALFIX4; break;
      case 2: t = ORC_SW_RATIONALFIX2; break;
      default: break;
    }
    if (t >= 0) {
      data_init(s, t, D0, DMAX, R0);
      return;
    }
  }
  data_init(s, ORC_SW_RATIONAL, D0, DMAX, R0);
  s->nn = N;
  s->mm = (M == 0) ? N * 2 : M;
  s->preRes = (double)s->nn / s->mm;
  s->preDfunc = 0.5 * s->nn * (s->nn - s->mm) / (double)s->mm;
  s->preSecDev = (s->nn * (s->mm * s->mm - 3.0 * s->mm * (-1 + s->nn) + s->nn * (-3 + 2 * s->nn))) / (6.0 * s->mm);
  s->nnf = s->nn / 2;
  s->mmf = s->mm / 2;
  s->preDfuncF = 0.5 * s->nnf * (s->nnf - s->mmf) / (double)s->mmf;
  s->preSecDevF =
      (s->nnf * (s->mmf * s->mmf - 3.0 * s->mmf * (-1 + s->nnf) + s->nnf * (-3 + 2 * s->nnf))) / (6.0 * s->mmf);
  if (2 * N == M || M == 0)
    s->type = fast ? ORC_SW_RATIONALSIMPLEFAST : ORC_SW_RATIONALSIMPLE;
  else
    s->type = fast ? ORC_SW_RATIONALFAST : ORC_SW_RATIONAL;
}

/* fixedRational<N>::doRational<POW>, tools/SwitchingFunction.cpp:190-196 */
static inline void do_rational_fixed(int POW, double rdist, double* result, double* dfunc) {
  const double rNdist = fastpow(rdist, POW - 1);
  *result = 1.0 / (1.0 + rNdist * rdist);
  *dfunc = -POW * rNdist * (*result) * (*result);
}

/* rational<>::doRational, tools/SwitchingFunction.cpp:258-283.
 * result/dfunc come in preset to preRes/preDfunc (used only by the Taylor patch). */
static inline void do_rational(int simplified, double rdist, double secDev, int N, int M, double* dfunc,
                               double* result) {
  const double moreThanOne = 1.0 + 5.0e10 * ORC_EPS;
  const double lessThanOne = 1.0 - 5.0e10 * ORC_EPS;
  if (simplified) {
    const double rNdist = fastpow(rdist, N - 1);
    *result = 1.0 / (1.0 + rNdist * rdist);
    *dfunc = -N * rNdist * (*result) * (*result);
  } else {
    if (!((rdist > lessThanOne) && (rdist < moreThanOne))) {
      const double rNdist = fastpow(rdist, N - 1);
      const double rMdist = fastpow(rdist, M - 1);
      const double num = 1.0 - rNdist * rdist;
      const double iden = 1.0 / (1.0 - rMdist * rdist);
      *result = num * iden;
      *dfunc = ((M * (*result) * rMdist) - (N * rNdist)) * iden;
    } else {
      const double x = (rdist - 1.0);
      *result = *result + x * (*dfunc + 0.5 * x * secDev);
      *dfunc = *dfunc + x * secDev;
    }
  }
}

static int fixed_N(int type) {
  switch (type) {
    case ORC_SW_RATIONALFIX12: return 12;
    case ORC_SW_RATIONALFIX10: return 10;
    case ORC_SW_RATIONALFIX8: return 8;
    case ORC_SW_RATIONALFIX6: return 6;
    case ORC_SW_RATIONALFIX4: return 4;
    case ORC_SW_RATIONALFIX2: return 2;
    default: return 0;
  }
}

/* the per-type f(rdist), f'(rdist): tools/SwitchingFunction.cpp:185-507 */
static void sw_function(const orc_switch* s, double rdist, double* result, double* dfunc) {
  switch (s->type) {
    case ORC_SW_RATIONALFIX12:
    case ORC_SW_RATIONALFIX10:
    case ORC_SW_RATIONALFIX8:
    case ORC_SW_RATIONALFIX6:
    case ORC_SW_RATIONALFIX4:
    case ORC_SW_RATIONALFIX2:
      do_rational_fixed(fixed_N(s->type), rdist, result, dfunc); /* :198-201 */
      break;
    case ORC_SW_RATIONAL:
    case ORC_SW_RATIONALFAST:
      *result = s->preRes;
      *dfunc = s->preDfunc;
      do_rational(0, rdist, s->preSecDev, s->nn, s->mm, dfunc, result); /* :284-287 */
      break;
    case ORC_SW_RATIONALSIMPLE:
    case ORC_SW_RATIONALSIMPLEFAST:
      *result = s->preRes;
      *dfunc = s->preDfunc;
      do_rational(1, rdist, s->preSecDev, s->nn, s->mm, dfunc, result);
      break;
    case ORC_SW_EXPONENTIAL: /* :375-387 */
      *result = exp(-rdist);
      *dfunc = -*result;
      break;
    case ORC_SW_GAUSSIAN:     /* :389-401 */
    case ORC_SW_FASTGAUSSIAN: /* :403-413 */
      *result = exp(-0.5 * rdist * rdist);
      *dfunc = -rdist * (*result);
      break;
    case ORC_SW_SMAP: { /* :434-455 */
      const double sx = s->c * fastpow(rdist, s->a);
      *result = pow(1.0 + sx, s->d);
      *dfunc = -s->b * sx / rdist * (*result) / (1.0 + sx);
    } break;
    case ORC_SW_CUBIC: { /* :457-469 */
      const double tmp1 = rdist - 1.0;
      const double tmp2 = 1.0 + 2.0 * rdist;
      *dfunc = 2 * tmp1 * tmp2 + 2 * tmp1 * tmp1;
      *result = tmp1 * tmp1 * tmp2;
    } break;
    case ORC_SW_TANH: { /* :471-486 */
      const double tmp1 = tanh(rdist);
      *dfunc = tmp1 * tmp1 - 1.0;
      *result = 1.0 - tmp1;
    } break;
    case ORC_SW_COSINUS: /* :488-507 */
      *result = 0.0;
      *dfunc = 0.0;
      if (rdist <= 1.0) {
        double rdistPI = rdist * ORC_PI;
        *result = 0.5 * (cos(rdistPI) + 1.0);
        *dfunc = -0.5 * ORC_PI * sin(rdistPI);
      }
      break;
    default:
      *result = 0.0;
      *dfunc = 0.0;
  }
}

/* baseSwitch::calculate :135-149 + applystretch :124-130 ; nativeq :524-549 */
double orc_switch_calculate(const orc_switch* s, double distance, double* dfunc) {
  if (s->type == ORC_SW_NATIVEQ) {
    double res = 0.0, df = 0.0;
    if (distance <= s->dmax) {
      res = 1.0;
      if (distance > s->d0) {
        const double rdist = s->beta * (distance - s->lambda * s->ref);
        double exprdist = exp(rdist);
        res = 1.0 / (1.0 + exprdist);
        df = -s->beta / (exprdist + 2.0 + 1.0 / exprdist) / distance;
        df *= s->stretch;
      }
      res = res * s->stretch + s->shift;
    }
    *dfunc = df;
    return res;
  }
  const double rdist = (distance - s->d0) * s->invr0;
  if (distance > s->dmax) {
    *dfunc = 0.0;
    return 0.0;
  }
  if (rdist > 0.0) {
    double f, d;
    sw_function(s, rdist, &f, &d);
    f = f * s->stretch + s->shift;
    d *= s->stretch;
    d *= s->invr0;
    d /= distance;
    *dfunc = d;
    return f;
  }
  *dfunc = 0.0;
  return s->stretch + s->shift;
}

/* SwitchingFunction::calculateSqr :1167-1169 -> per-type calculateSqr */
double orc_switch_calculate_sqr(const orc_switch* s, double distance2, double* dfunc) {
  double result = 0.0, df = 0.0;
  switch (s->type) {
    case ORC_SW_RATIONALFIX12:
    case ORC_SW_RATIONALFIX10:
    case ORC_SW_RATIONALFIX8:
    case ORC_SW_RATIONALFIX6:
    case ORC_SW_RATIONALFIX4:
    case ORC_SW_RATIONALFIX2: /* fixedRational<N>::calculateSqr :203-215 */
      if (distance2 <= s->dmax_2) {
        const double rdist = distance2 * s->invr0_2;
        do_rational_fixed(fixed_N(s->type) / 2, rdist, &result, &df);
        df *= 2 * s->invr0_2;
        result = result * s->stretch + s->shift;
        df *= s->stretch;
      }
      *dfunc = df;
      return result;
    case ORC_SW_RATIONALFAST:
    case ORC_SW_RATIONALSIMPLEFAST: /* rational<fast,*>::calculateSqr :289-303 */
      if (distance2 <= s->dmax_2) {
        const double rdist = distance2 * s->invr0_2;
        result = s->preRes;
        df = s->preDfuncF;
        do_rational(s->type == ORC_SW_RATIONALSIMPLEFAST, rdist, s->preSecDevF, s->nnf, s->mmf, &df, &result);
        df *= 2 * s->invr0_2;
        result = result * s->stretch + s->shift;
        df *= s->stretch;
      }
      *dfunc = df;
      return result;
    case ORC_SW_FASTGAUSSIAN: /* :414-431 */
      if (distance2 < s->dmax_2) {
        result = 1.0;
        if (distance2 > 0.0) {
          result = exp(-0.5 * distance2);
          df = -result;
          result = result * s->stretch + s->shift;
          df *= s->stretch;
        }
      }
      *dfunc = df;
      return result;
    default: /* baseSwitch::calculateSqr :181-183 */
      return orc_switch_calculate(s, sqrt(distance2), dfunc);
  }
}

/* SwitchInterface::setupStretch, tools/SwitchingFunction.cpp:63-72 */
static void setup_stretch(orc_switch* s) {
  if (s->dmax < DBL_MAX) {
    double dummy;
    s->stretch = 1.0;
    s->shift = 0.0;
    double s0 = orc_switch_calculate(s, 0.0, &dummy);
    double sd = orc_switch_calculate(s, s->dmax, &dummy);
    s->stretch = 1.0 / (s0 - sd);
    s->shift = -sd * s->stretch;
  }
}

/* SwitchingFunction::set(nn,mm,r0,d0), tools/SwitchingFunction.cpp:1176-1184 */
void orc_switch_set_rational(orc_switch* s, int nn, int mm, double r0, double d0) {
  if (mm == 0) mm = 2 * nn;
  double dmax = d0 + r0 * pow(0.00001, 1. / (nn - mm));
  rational_factory(s, d0, dmax, r0, nn, mm);
  setup_stretch(s);
}

/* --- tiny word parser standing in for Tools::getWords / Tools::parse / Tools::parseFlag */
#define MAXW 64
typedef struct {
  char* w[MAXW];
  int n;
} words_t;
static int find_key(words_t* ws, const char* key) { /* index of "KEY=..." or -1 */
  size_t kl = strlen(key);
  for (int i = 0; i < ws->n; i++)
    if (strncmp(ws->w[i], key, kl) == 0 && ws->w[i][kl] == '=') return i;
  return -1;
}
static void drop_word(words_t* ws, int i) {
  for (int k = i; k + 1 < ws->n; k++) ws->w[k] = ws->w[k + 1];
  ws->n--;
}
/* returns 1 parsed, 0 absent, -1 malformed */
static int parse_double(words_t* ws, const char* key, double* v) {
  int i = find_key(ws, key);
  if (i < 0) return 0;
  char* end;
  const char* txt = ws->w[i] + strlen(key) + 1;
  double x = strtod(txt, &end);
  int ok = (end != txt && *end == 0);
  drop_word(ws, i);
  if (!ok) return -1;
  *v = x;
  return 1;
}
static int parse_int(words_t* ws, const char* key, int* v) {
  double x;
  int r = parse_double(ws, key, &x);
  if (r == 1) {
    if (x != floor(x)) return -1;
    *v = (int)x;
  }
  return r;
}
static int parse_flag(words_t* ws, const char* key) {
  for (int i = 0; i < ws->n; i++)
    if (strcmp(ws->w[i], key) == 0) {
      drop_word(ws, i);
      return 1;
    }
  return 0;
}

/* SwitchingFunction::set(definition, errormsg), tools/SwitchingFunction.cpp:1055-1159 */
int orc_switch_set(orc_switch* s, const char* definition, char* err, int errlen) {
  char buf[1024];
  words_t ws;
  ws.n = 0;
  if (err && errlen > 0) err[0] = 0;
#define SETERR(...)                                  \
  do {                                               \
    if (err && errlen > 0) snprintf(err, errlen, __VA_ARGS__); \
  } while (0)
  strncpy(buf, definition, sizeof(buf) - 1);
  buf[sizeof(buf) - 1] = 0;
  for (char* p = buf; *p; p++)
    if (*p == '{' || *p == '}') *p = ' ';
  for (char* tok = strtok(buf, " \t\n"); tok && ws.n < MAXW; tok = strtok(NULL, " \t\n")) ws.w[ws.n++] = tok;
  if (ws.n < 1) {
    SETERR("missing all input for switching function");
    return 1;
  }
  char name[64];
  strncpy(name, ws.w[0], sizeof(name) - 1);
  name[sizeof(name) - 1] = 0;
  drop_word(&ws, 0);
  int bad = 0;
  double d0 = 0.0, dmax = DBL_MAX;
  if (parse_double(&ws, "D_0", &d0) < 0) { SETERR("could not parse D_0"); bad = 1; }
  if (parse_double(&ws, "D_MAX", &dmax) < 0) { SETERR("could not parse D_MAX"); bad = 1; }
  int dostretch = 1;
  parse_flag(&ws, "STRETCH");
  if (parse_flag(&ws, "NOSTRETCH")) dostretch = 0;
  s->type = ORC_SW_NOT_INITIALIZED;
  if (strcmp(name, "CUBIC") == 0) {
    data_init(s, ORC_SW_CUBIC, d0, dmax, dmax - d0); /* cubicSwitch::init :458-461 */
  } else {
    double r0 = 0.0;
    if (parse_double(&ws, "R_0", &r0) != 1) { SETERR("R_0 is required for %s", name); bad = 1; }
    if (strcmp(name, "RATIONAL") == 0) {
      int nn = 6, mm = 0;
      if (parse_int(&ws, "NN", &nn) < 0) { SETERR("could not parse NN"); bad = 1; }
      if (parse_int(&ws, "MM", &mm) < 0) { SETERR("could not parse MM"); bad = 1; }
      rational_factory(s, d0, dmax, r0, nn, mm);
    } else if (strcmp(name, "SMAP") == 0) {
      int a = 0, b = 0;
      if (parse_int(&ws, "A", &a) != 1) { SETERR("A is required for %s", name); bad = 1; }
      if (parse_int(&ws, "B", &b) != 1) { SETERR("B is required for %s", name); bad = 1; }
      data_init(s, ORC_SW_SMAP, d0, dmax, r0); /* smapSwitch::init :435-447 */
      s->a = a;
      s->b = b;
      s->c = pow(2., (double)a / (double)b) - 1.0;
      s->d = -(double)b / (double)a;
    } else if (strcmp(name, "Q") == 0) {
      double beta = 50.0, lambda = 1.8, ref = 0.0;
      if (parse_double(&ws, "BETA", &beta) < 0) { SETERR("could not parse BETA"); bad = 1; }
      if (parse_double(&ws, "LAMBDA", &lambda) < 0) { SETERR("could not parse LAMBDA"); bad = 1; }
      if (parse_double(&ws, "REF", &ref) != 1) { SETERR("REF is required for %s", name); bad = 1; }
      data_init(s, ORC_SW_NATIVEQ, d0, dmax, r0);
      s->beta = beta;
      s->lambda = lambda;
      s->ref = ref;
    } else if (strcmp(name, "EXP") == 0) {
      data_init(s, ORC_SW_EXPONENTIAL, d0, dmax, r0);
    } else if (strcmp(name, "GAUSSIAN") == 0) {
      if (r0 == 1.0 && d0 == 0.0)
        data_init(s, ORC_SW_FASTGAUSSIAN, 0.0, dmax, 1.0);
      else
        data_init(s, ORC_SW_GAUSSIAN, d0, dmax, r0);
    } else if (strcmp(name, "TANH") == 0) {
      data_init(s, ORC_SW_TANH, d0, dmax, r0);
    } else if (strcmp(name, "COSINUS") == 0) {
      data_init(s, ORC_SW_COSINUS, d0, dmax, r0);
    } else if (strcmp(name, "MATHEVAL") == 0 || strcmp(name, "CUSTOM") == 0) {
      SETERR("CUSTOM/MATHEVAL (lepton) switching functions are outside the oracle's scope");
      return 2;
    } else {
      SETERR("cannot understand switching function type '%s'", name);
      return 1;
    }
  }
  if (ws.n > 0) {
    char tmp[512];
    tmp[0] = 0;
    for (int i = 0; i < ws.n; i++) {
      strncat(tmp, ws.w[i], sizeof(tmp) - strlen(tmp) - 2);
      strcat(tmp, " ");
    }
    SETERR("found the following rogue keywords in switching function input : %s", tmp);
    bad = 1;
  }
  if (bad) return 1;
  if (dostretch && dmax != DBL_MAX) setup_stretch(s);
  return 0;
#undef SETERR
}

/* ------------------------------------------------------------------ LatticeReduction */
static const double LR_EPS = 1e-14; /* tools/LatticeReduction.cpp:30 */

/* LatticeReduction::sort :32-58 (the consistency asserts are dropped) */
static void lr_sort(double v[3][3]) {
  double m[3];
  for (int i = 0; i < 3; i++) m[i] = mod2(v[i]);
  for (int i = 0; i < 3; i++)
    for (int j = i + 1; j < 3; j++)
      if (m[i] > m[j]) {
        double t[3];
        memcpy(t, v[i], sizeof(t));
        memcpy(v[i], v[j], sizeof(t));
        memcpy(v[j], t, sizeof(t));
        double tm = m[i];
        m[i] = m[j];
        m[j] = tm;
      }
}
/* LatticeReduction::reduce(a,b) :60-86 */
static void lr_reduce2v(double a[3], double b[3]) {
  const double onePlusEpsilon = (1.0 + LR_EPS);
  double ma = mod2(a), mb = mod2(b);
  unsigned counter = 0;
  for (;;) {
    if (mb > ma) {
      double t[3];
      memcpy(t, a, sizeof(t));
      memcpy(a, b, sizeof(t));
      memcpy(b, t, sizeof(t));
      double tm = ma;
      ma = mb;
      mb = tm;
    }
    double f = floor(dot3(a, b) / mb + 0.5);
    for (int k = 0; k < 3; k++) a[k] -= b[k] * f;
    ma = mod2(a);
    if (mb <= ma * onePlusEpsilon) break;
    if (++counter > 1000000u) break;
  }
  double t[3];
  memcpy(t, a, sizeof(t));
  memcpy(a, b, sizeof(t));
  memcpy(b, t, sizeof(t));
}
/* LatticeReduction::reduceFast :144-192 */
void orc_lattice_reduce(double t[9]) {
  const double onePlusEpsilon = (1.0 + LR_EPS);
  double v[3][3];
  memcpy(v, t, sizeof(v));
  unsigned counter = 0;
  for (;;) {
    lr_sort(v);
    lr_reduce2v(v[0], v[1]);
    double b11 = mod2(v[0]);
    double b22 = mod2(v[1]);
    double b12 = dot3(v[0], v[1]);
    double b13 = dot3(v[0], v[2]);
    double b23 = dot3(v[1], v[2]);
    double z = b11 * b22 - b12 * b12;
    double y2 = -(b11 * b23 - b12 * b13) / z;
    double y1 = -(b22 * b13 - b12 * b23) / z;
    int x1min = (int)floor(y1);
    int x1max = x1min + 1;
    int x2min = (int)floor(y2);
    int x2max = x2min + 1;
    int first = 1;
    double mbest = 0, best[3] = {0, 0, 0};
    for (int x1 = x1min; x1 <= x1max; x1++)
      for (int x2 = x2min; x2 <= x2max; x2++) {
        double trial[3];
        /* trial=v[2]+x2*v[1]+x1*v[0] : (v2 + x2*v1) + x1*v0 */
        for (int k = 0; k < 3; k++) trial[k] = (v[2][k] + x2 * v[1][k]) + x1 * v[0][k];
        double mtrial = mod2(trial);
        if (first || mtrial < mbest) {
          mbest = mtrial;
          memcpy(best, trial, sizeof(best));
          first = 0;
        }
      }
    if (mod2(best) * onePlusEpsilon >= mod2(v[2])) break;
    if (++counter > 1000000u) break;
    memcpy(v[2], best, sizeof(best));
  }
  lr_sort(v);
  memcpy(t, v, sizeof(v));
}

/* ------------------------------------------------------------------ Pbc */
/* Pbc::buildShifts, tools/Pbc.cpp:59-135 */
static void build_shifts(orc_pbc* p) {
  const double small = 1e-28;
  for (int o = 0; o < 8; o++) p->nshift[o] = 0;
  double rt[9], rrt[9];
  transpose3(p->reduced, rt);
  matmat(p->reduced, rt, rrt);
  for (int l = -1; l <= 1; l++)
    for (int m = -1; m <= 1; m++)
      for (int n = -1; n <= 1; n++) {
        const int ishift[3] = {l, m, n};
        double dshift[3] = {(double)l, (double)m, (double)n};
        unsigned count = 0;
        for (int s = 0; s < 3; s++)
          if (ishift[s] != 0) count++;
        if (count == 0 || count == 3) continue;
        double cosdir[3];
        matvec(rrt, dshift, cosdir);
        double dp = dot3(dshift, cosdir);
        double ref = mod2(dshift) * mod2(cosdir);
        if (fabs(ref - dp * dp) < small) continue;
        for (int i = 0; i < 2; i++)
          for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
              const int block[3] = {2 * i - 1, 2 * j - 1, 2 * k - 1};
              int skip = 0;
              for (int s = 0; s < 3; s++)
                if (ishift[s] * block[s] > 0) skip = 1;
              if (skip) continue;
              skip = 1;
              for (int s = 0; s < 3; s++)
                if (((1 - ishift[s] * ishift[s]) * block[s]) * cosdir[s] < -small) skip = 0;
              if (skip) continue;
              int o = 4 * i + 2 * j + k;
              if (p->nshift[o] < ORC_MAXSHIFT) {
                matvec(rt, dshift, p->shifts[o][p->nshift[o]]);
                p->nshift[o]++;
              }
            }
      }
}

/* Pbc::setBox, tools/Pbc.cpp:165-212 */
void orc_pbc_set_box(orc_pbc* p, const double b[9]) {
  memset(p, 0, sizeof(*p));
  memcpy(p->box, b, 9 * sizeof(double));
  const double boxEpsilon = 1e-28;
  p->type = ORC_PBC_UNSET;
  double det = det3(p->box);
  if (det * det < boxEpsilon) return;
  int cxy = 0, cxz = 0, cyz = 0;
  if (b[1] * b[1] < boxEpsilon && b[3] * b[3] < boxEpsilon) cxy = 1;
  if (b[2] * b[2] < boxEpsilon && b[6] * b[6] < boxEpsilon) cxz = 1;
  if (b[5] * b[5] < boxEpsilon && b[7] * b[7] < boxEpsilon) cyz = 1;
  inv3(p->box, p->invBox);
  p->type = (cxy && cxz && cyz) ? ORC_PBC_ORTHO : ORC_PBC_GENERIC;
  memcpy(p->reduced, p->box, sizeof(p->reduced));
  if (p->type == ORC_PBC_ORTHO) {
    inv3(p->reduced, p->invReduced);
  } else {
    orc_lattice_reduce(p->reduced);
    inv3(p->reduced, p->invReduced);
    build_shifts(p);
  }
}

/* Pbc::distance(v1,v2), tools/Pbc.cpp:362-415 ; delta = v2 - v1 (tools/Vector.h:303-306) */
void orc_pbc_distance(const orc_pbc* p, const double v1[3], const double v2[3], double d[3]) {
  for (int i = 0; i < 3; i++) d[i] = v2[i] - v1[i];
  if (p->type == ORC_PBC_UNSET) {
    return;
  } else if (p->type == ORC_PBC_ORTHO) {
    for (int i = 0; i < 3; i++) d[i] = orc_tools_pbc(d[i] * p->invBox[4 * i]) * p->box[4 * i];
  } else {
    double s[3];
    vecmat(d, p->invReduced, s);
    for (int i = 0; i < 3; i++) s[i] = orc_tools_pbc(s[i]);
    vecmat(s, p->reduced, d);
    if ((fabs(s[0]) + fabs(s[1]) + fabs(s[2]) > 0.5)) {
      int o = 4 * (s[0] > 0 ? 1 : 0) + 2 * (s[1] > 0 ? 1 : 0) + (s[2] > 0 ? 1 : 0);
      double best[3] = {d[0], d[1], d[2]};
      double lbest = mod2(best);
      for (int i = 0; i < p->nshift[o]; i++) {
        double trial[3] = {d[0] + p->shifts[o][i][0], d[1] + p->shifts[o][i][1], d[2] + p->shifts[o][i][2]};
        double ltrial = mod2(trial);
        if (ltrial < lbest) {
          lbest = ltrial;
          memcpy(best, trial, sizeof(best));
        }
      }
      memcpy(d, best, sizeof(best));
    }
  }
}

/* Pbc::fullSearch, tools/Pbc.cpp:137-163 (brute-force reference used by regtest/basic/rt-make-1) */
void orc_pbc_full_search(const orc_pbc* p, double d[3]) {
  if (p->type == ORC_PBC_UNSET) return;
  double irt[9], rt[9], s[3];
  transpose3(p->invReduced, irt);
  transpose3(p->reduced, rt);
  matvec(irt, d, s);
  for (int i = 0; i < 3; i++) s[i] = orc_tools_pbc(s[i]);
  matvec(rt, s, d);
  const int smax = 4;
  const double* a0 = p->reduced;
  const double* a1 = p->reduced + 3;
  const double* a2 = p->reduced + 6;
  double best[3] = {d[0], d[1], d[2]};
  double lbest = mod2(d);
  for (int i = -smax; i <= smax; i++)
    for (int j = -smax; j <= smax; j++)
      for (int k = -smax; k <= smax; k++) {
        double trial[3];
        for (int c = 0; c < 3; c++) trial[c] = ((d[c] + i * a0[c]) + j * a1[c]) + k * a2[c];
        double ltrial = mod2(trial);
        if (ltrial < lbest) {
          memcpy(best, trial, sizeof(best));
          lbest = ltrial;
        }
      }
  memcpy(d, best, sizeof(best));
}

/* ------------------------------------------------------------------ LinkCells */
/* LinkCells::createCells, tools/LinkCells.cpp:99-122 */
static void lc_create_cells(orc_linkcells* lc, const double box[9]) {
  orc_pbc_set_box(&lc->mypbc, box);
  for (int k = 0; k < 3; k++) {
    /* row k of transpose(invBox) = column k of invBox */
    double row[3] = {lc->mypbc.invBox[k], lc->mypbc.invBox[3 + k], lc->mypbc.invBox[6 + k]};
    double v = floor(1.0 / sqrt(mod2(row)) / lc->cutoff);
    lc->ncells[k] = (unsigned)v;
    if (lc->ncells[k] == 0) lc->ncells[k] = 1;
  }
  lc->nstride[0] = 1;
  lc->nstride[1] = lc->ncells[0];
  lc->nstride[2] = lc->ncells[0] * lc->ncells[1];
}

/* LinkCells::setupCells(pos,pbc) :85-97 -> setupCells(pos) :49-73 or setupCells(pbc) :75-83 */
void orc_linkcells_setup(orc_linkcells* lc, double cutoff, const double* pos, size_t n, const orc_pbc* pbc) {
  memset(lc, 0, sizeof(*lc));
  lc->cutoff = cutoff;
  int allzero = 1;
  for (int k = 0; k < 9; k++)
    if (pbc->box[k] != 0.0) allzero = 0;
  if (allzero) {
    lc->nopbc = 1;
    double box[9] = {0};
    for (unsigned k = 0; k < 3; ++k) {
      double minp = pos[k], maxp = pos[k];
      for (size_t i = 1; i < n; ++i) {
        if (pos[3 * i + k] > maxp) maxp = pos[3 * i + k];
        if (pos[3 * i + k] < minp) minp = pos[3 * i + k];
      }
      if (cutoff < sqrt(DBL_MAX))
        box[4 * k] = cutoff * (1 + ceil((maxp - minp) / cutoff));
      else
        box[4 * k] = maxp - minp + 1;
      lc->origin[k] = (minp + maxp) / 2;
    }
    lc_create_cells(lc, box);
  } else {
    lc->nopbc = 0;
    lc_create_cells(lc, pbc->box);
  }
}

/* LinkCells::findMyCell(pos) :277-292 */
static void lc_find_my_cell(const orc_linkcells* lc, const double pos[3], unsigned celn[3]) {
  double mypos[3] = {pos[0], pos[1], pos[2]};
  if (lc->nopbc)
    for (int k = 0; k < 3; k++) mypos[k] = mypos[k] - lc->origin[k];
  double ibt[9], fpos[3];
  transpose3(lc->mypbc.invBox, ibt);
  matvec(ibt, mypos, fpos); /* Pbc::realToScaled :472-474 */
  for (unsigned j = 0; j < 3; ++j) celn[j] = (unsigned)floor((orc_tools_pbc(fpos[j]) + 0.5) * lc->ncells[j]);
}
unsigned orc_linkcells_find_cell(const orc_linkcells* lc, const double pos[3]) {
  unsigned c[3];
  lc_find_my_cell(lc, pos, c);
  return c[0] + c[1] * lc->nstride[1] + c[2] * lc->nstride[2]; /* :306-311 */
}

/* min_cell / max_cell :183-193 and addRequiredCells :195-239 (called with ncells_required==0) */
unsigned orc_linkcells_required(const orc_linkcells* lc, const unsigned celn[3], int usePbc, unsigned* out) {
  int lo[3], hi[3];
  for (int n = 0; n < 3; n++) {
    int nc = (int)lc->ncells[n];
    int c = (int)celn[n];
    int mn = c + ((nc < 2) ? 0 : -1);
    if (!usePbc && mn < 0) mn = 0;
    int mx = c + ((nc < 3 && usePbc) ? 1 : 2);
    if (!usePbc && mx > nc) mx = nc;
    lo[n] = mn;
    hi[n] = mx;
  }
  unsigned cnt = 0;
#define LINKC_PBC(n, num) (((n) < 0) ? (num)-1 : (n) % (num))
  for (int nx = lo[0]; nx < hi[0]; ++nx) {
    int xval = LINKC_PBC(nx, (int)lc->ncells[0]) * (int)lc->nstride[0];
    for (int ny = lo[1]; ny < hi[1]; ++ny) {
      int yval = LINKC_PBC(ny, (int)lc->ncells[1]) * (int)lc->nstride[1];
      for (int nz = lo[2]; nz < hi[2]; ++nz) {
        int zval = LINKC_PBC(nz, (int)lc->ncells[2]) * (int)lc->nstride[2];
        out[cnt++] = (unsigned)(xval + yval + zval);
      }
    }
  }
#undef LINKC_PBC
  return cnt;
}

/* counting sort of a group into cells: LinkCells::resetCollection :138-181 */
typedef struct {
  unsigned *starts, *tots, *lists;
} cellcoll;
static void lc_collect(const orc_linkcells* lc, const double* pos, unsigned first, unsigned n, cellcoll* cc) {
  unsigned nct = lc->ncells[0] * lc->ncells[1] * lc->ncells[2];
  unsigned* allcells = (unsigned*)malloc(sizeof(unsigned) * (n ? n : 1));
  cc->starts = (unsigned*)calloc(nct, sizeof(unsigned));
  cc->tots = (unsigned*)calloc(nct, sizeof(unsigned));
  cc->lists = (unsigned*)malloc(sizeof(unsigned) * (n ? n : 1));
  for (unsigned i = 0; i < n; i++) {
    allcells[i] = orc_linkcells_find_cell(lc, pos + 3 * (size_t)(first + i));
    cc->tots[allcells[i]]++;
  }
  unsigned tot = 0;
  for (unsigned c = 0; c < nct; c++) {
    cc->starts[c] = tot;
    tot += cc->tots[c];
    cc->tots[c] = 0;
  }
  for (unsigned j = 0; j < n; j++) {
    unsigned myind = cc->starts[allcells[j]] + cc->tots[allcells[j]];
    cc->lists[myind] = first + j;
    cc->tots[allcells[j]]++;
  }
  free(allcells);
}
static void cc_free(cellcoll* cc) {
  free(cc->starts);
  free(cc->tots);
  free(cc->lists);
}

/* ------------------------------------------------------------------ NeighborList */
orc_nl* orc_nl_create(int style, unsigned n0, unsigned n1, int do_pbc, int use_cells, double cutoff, unsigned stride) {
  orc_nl* nl = (orc_nl*)calloc(1, sizeof(orc_nl));
  nl->style = style;
  nl->nlist0 = n0;
  nl->nlist1 = (style == ORC_NL_SINGLELIST) ? 0 : n1;
  nl->do_pbc = do_pbc;
  nl->use_cells = use_cells;
  nl->cutoff = cutoff;
  nl->stride = stride;
  /* NeighborList ctors :43-101 */
  if (style == ORC_NL_PAIR)
    nl->nallpairs = n0;
  else if (style == ORC_NL_TWOLIST)
    nl->nallpairs = (size_t)n0 * n1;
  else
    nl->nallpairs = (size_t)n0 * (n0 - 1) / 2;
  /* initialize() :105-141 : with no NL (stride==0) or Pair the list is every pair, prefilled.
   * We do not materialise it: orc_nl_index_pair() is used on the fly (same order). */
  nl->list_built = 0;
  return nl;
}
void orc_nl_free(orc_nl* nl) {
  if (!nl) return;
  free(nl->pairs);
  free(nl);
}

/* NeighborList::getIndexPair :147-166, in 64-bit arithmetic (the reference's `unsigned` 8*ii+1
 * overflows for N>32768; the row-major upper-triangle ORDER is what is restated). */
void orc_nl_index_pair(const orc_nl* nl, size_t ipair, unsigned* i0, unsigned* i1) {
  switch (nl->style) {
    case ORC_NL_PAIR:
      *i0 = (unsigned)ipair;
      *i1 = (unsigned)(ipair + nl->nlist0);
      break;
    case ORC_NL_TWOLIST:
      *i0 = (unsigned)(ipair / nl->nlist1);
      *i1 = (unsigned)(ipair % nl->nlist1 + nl->nlist0);
      break;
    default: {
      size_t ii = nl->nallpairs - 1 - ipair;
      size_t K = (size_t)floor((sqrt((double)(8 * ii + 1)) + 1) / 2);
      /* guard the sqrt rounding for very large ii (not reachable in the reference's 32-bit range) */
      while (K * (K - 1) / 2 > ii) K--;
      while ((K + 1) * K / 2 <= ii) K++;
      size_t jj = ii - K * (K - 1) / 2;
      *i0 = (unsigned)(nl->nlist0 - 1 - K);
      *i1 = (unsigned)(nl->nlist0 - 1 - jj);
    }
  }
}

static void nl_push(orc_nl* nl, unsigned a, unsigned b) {
  if (nl->npairs == nl->cap) {
    nl->cap = nl->cap ? nl->cap * 2 : 1024;
    nl->pairs = (unsigned*)realloc(nl->pairs, nl->cap * 2 * sizeof(unsigned));
  }
  nl->pairs[2 * nl->npairs] = a;
  nl->pairs[2 * nl->npairs + 1] = b;
  nl->npairs++;
}

/* NeighborList::update :168-315 */
void orc_nl_update(orc_nl* nl, const orc_pbc* pbc, const double* pos) {
  nl->npairs = 0;
  size_t ntot = (size_t)nl->nlist0 + nl->nlist1;
  if (nl->use_cells) {
    orc_linkcells lc;
    orc_linkcells_setup(&lc, nl->cutoff, pos, ntot, pbc); /* :177-182 */
    unsigned nct = lc.ncells[0] * lc.ncells[1] * lc.ncells[2];
    unsigned req[27];
    cellcoll A, B;
    lc_collect(&lc, pos, 0, nl->nlist0, &A);
    if (nl->style == ORC_NL_TWOLIST) lc_collect(&lc, pos, nl->nlist0, nl->nlist1, &B);
    cellcoll* other = (nl->style == ORC_NL_TWOLIST) ? &B : &A;
    for (unsigned c = 0; c < nct; ++c) {
      if (A.tots[c] == 0) continue;
      /* findMyCell(cellIndex) :294-304 */
      unsigned cell[3];
      cell[2] = c / lc.nstride[2];
      unsigned rem = c % lc.nstride[2];
      cell[1] = rem / lc.nstride[1];
      cell[0] = rem % lc.nstride[1];
      unsigned nreq = orc_linkcells_required(&lc, cell, nl->do_pbc, req);
      for (unsigned ia = 0; ia < A.tots[c]; ia++) {
        unsigned a = A.lists[A.starts[c] + ia];
        for (unsigned cb = 0; cb < nreq; ++cb) {
          unsigned oc = req[cb];
          for (unsigned ib = 0; ib < other->tots[oc]; ib++) {
            unsigned b = other->lists[other->starts[oc] + ib];
            if (nl->style == ORC_NL_SINGLELIST && !(b > a)) continue; /* :224 */
            nl_push(nl, a, b);
          }
        }
      }
    }
    cc_free(&A);
    if (nl->style == ORC_NL_TWOLIST) cc_free(&B);
  } else {
    const double d2 = nl->cutoff * nl->cutoff; /* :238 */
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if (nl->nallpairs < 100000) nt = 1;
    orc_nl* parts = (orc_nl*)calloc((size_t)nt, sizeof(orc_nl));
#pragma omp parallel num_threads(nt)
    {
      int t = 0;
#ifdef _OPENMP
      t = omp_get_thread_num();
#endif
      size_t chunk = (nl->nallpairs + nt - 1) / nt;
      size_t lo = (size_t)t * chunk, hi = lo + chunk;
      if (hi > nl->nallpairs) hi = nl->nallpairs;
      for (size_t i = lo; i < hi; ++i) {
        unsigned i0, i1;
        double d[3];
        orc_nl_index_pair(nl, i, &i0, &i1);
        if (nl->do_pbc)
          orc_pbc_distance(pbc, pos + 3 * (size_t)i0, pos + 3 * (size_t)i1, d);
        else
          for (int k = 0; k < 3; k++) d[k] = pos[3 * (size_t)i1 + k] - pos[3 * (size_t)i0 + k];
        double value = mod2(d);
        if (value <= d2) nl_push(&parts[t], i0, i1); /* :256 */
      }
    }
    for (int t = 0; t < nt; t++) {
      for (size_t k = 0; k < parts[t].npairs; k++) nl_push(nl, parts[t].pairs[2 * k], parts[t].pairs[2 * k + 1]);
      free(parts[t].pairs);
    }
    free(parts);
  }
  nl->list_built = 1;
}

static int cmp_pair(const void* a, const void* b) {
  const unsigned* x = (const unsigned*)a;
  const unsigned* y = (const unsigned*)b;
  if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
  if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
  return 0;
}

/* NOT a reference algorithm: same pair set as the classic branch above (value<=d2 with the same
 * arithmetic), candidates found through a cell grid.  Pairs come out sorted by (i0,i1). */
void orc_nl_update_classic_cells(orc_nl* nl, const orc_pbc* pbc, const double* pos) {
  nl->npairs = 0;
  if (nl->style == ORC_NL_PAIR) {
    orc_nl_update(nl, pbc, pos);
    return;
  }
  size_t ntot = (size_t)nl->nlist0 + nl->nlist1;
  orc_pbc zero;
  memset(&zero, 0, sizeof(zero));
  int periodic = nl->do_pbc && pbc->type != ORC_PBC_UNSET;
  orc_linkcells lc;
  /* a little slack on the cell width guards the floor() in createCells against rounding */
  orc_linkcells_setup(&lc, nl->cutoff * (1.0 + 1e-9), pos, ntot, periodic ? pbc : &zero);
  unsigned nct = lc.ncells[0] * lc.ncells[1] * lc.ncells[2];
  const double d2 = nl->cutoff * nl->cutoff;
  cellcoll O;
  unsigned ofirst = (nl->style == ORC_NL_TWOLIST) ? nl->nlist0 : 0;
  unsigned on = (nl->style == ORC_NL_TWOLIST) ? nl->nlist1 : nl->nlist0;
  lc_collect(&lc, pos, ofirst, on, &O);
  (void)nct;
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  orc_nl* parts = (orc_nl*)calloc((size_t)nt, sizeof(orc_nl));
#pragma omp parallel num_threads(nt)
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    unsigned req[27];
#pragma omp for schedule(static)
    for (long a = 0; a < (long)nl->nlist0; a++) {
      unsigned cell[3];
      lc_find_my_cell(&lc, pos + 3 * (size_t)a, cell);
      unsigned nreq = orc_linkcells_required(&lc, cell, periodic, req);
      for (unsigned cb = 0; cb < nreq; cb++) {
        unsigned oc = req[cb];
        for (unsigned ib = 0; ib < O.tots[oc]; ib++) {
          unsigned b = O.lists[O.starts[oc] + ib];
          if (nl->style == ORC_NL_SINGLELIST && !(b > (unsigned)a)) continue;
          double d[3];
          if (nl->do_pbc)
            orc_pbc_distance(pbc, pos + 3 * (size_t)a, pos + 3 * (size_t)b, d);
          else
            for (int k = 0; k < 3; k++) d[k] = pos[3 * (size_t)b + k] - pos[3 * (size_t)a + k];
          if (mod2(d) <= d2) nl_push(&parts[t], (unsigned)a, b);
        }
      }
    }
  }
  for (int t = 0; t < nt; t++) {
    for (size_t k = 0; k < parts[t].npairs; k++) nl_push(nl, parts[t].pairs[2 * k], parts[t].pairs[2 * k + 1]);
    free(parts[t].pairs);
  }
  free(parts);
  cc_free(&O);
  qsort(nl->pairs, nl->npairs, 2 * sizeof(unsigned), cmp_pair);
  nl->list_built = 1;
}

/* NeighborList::size :369-375 */
size_t orc_nl_size(const orc_nl* nl) { return nl->list_built ? nl->npairs : nl->nallpairs; }
const unsigned* orc_nl_pairs(const orc_nl* nl) { return nl->pairs; }

/* NeighborList::prepare :433-456 (the requestAtoms side effects are the caller's business) */
void orc_nl_prepare(const orc_nl* nl, long step, int exchange_step, int* firsttime, int* invalidate) {
  if (nl->stride > 0) {
    if (nl->stride == 1) {
      *invalidate = 1;
      *firsttime = 0;
    } else if (*firsttime || (step % (long)nl->stride == 0)) {
      *invalidate = 1;
      *firsttime = 0;
    } else {
      *invalidate = 0;
    }
    if (exchange_step) *firsttime = 1;
  }
}

/* ------------------------------------------------------------------ CoordinationBase::calculate */
/* colvar/CoordinationBase.cpp:142-232 */
/* DHEnergy::pairing, colvar/DHEnergy.cpp:130-143 (beta = k, lambda = constant, ref = epsilon) */
static double dhenergy_pairing(const orc_switch* s, double distance2, double qi, double qj, double* dfunc) {
  double distance = sqrt(distance2);
  double invdistance = 1.0 / distance;
  double tmp = exp(-s->beta * distance) * invdistance * s->lambda * qi * qj / s->ref;
  double dtmp = -(s->beta + invdistance) * tmp;
  *dfunc = dtmp * invdistance;
  return tmp;
}

/* GHBFIX::pairing, colvar/GHBFIX.cpp:186-220 (preRes = A, preDfunc = B, preSecDev = C, d = D, c = c*dmax2) */
static double ghbfix_pairing(const orc_switch* s, double distance2, double scale, double* dfunc) {
  double result;
  if (distance2 > s->dmax_2) {
    *dfunc = 0.0;
    return 0.;
  }
  double distance = sqrt(distance2);
  const double rdist = (distance - s->d0);
  if (rdist <= 0.) {
    result = -1.;
    *dfunc = 0.0;
  } else {
    result = -1.;
    *dfunc = 0.0;
    if (rdist > s->c) {
      result += (s->preRes + s->preDfunc * rdist + s->preSecDev * rdist * rdist);
      *dfunc += s->preDfunc + 2 * s->preSecDev * rdist;
    } else if (rdist > 0.0) {
      result += s->d * (rdist * rdist);
      *dfunc += 2 * s->d * rdist;
    }
    *dfunc /= distance;
  }
  result *= scale;
  *dfunc *= scale;
  return result;
}

/* GHBFIX::GHBFIX, colvar/GHBFIX.cpp:98-113 */
void orc_ghbfix_setup(orc_switch* sw, double dmax, double d0, double c) {
  memset(sw, 0, sizeof(*sw));
  sw->type = ORC_PAIR_GHBFIX;
  sw->dmax = dmax;
  sw->dmax_2 = dmax * dmax;
  sw->d0 = d0;
  const double dmax2 = dmax - d0;
  sw->preRes = (-c * dmax2 * dmax2) / ((1 - c) * dmax2 * dmax2);
  sw->preDfunc = (2 * dmax2) / ((1 - c) * dmax2 * dmax2);
  sw->preSecDev = -1 / ((1 - c) * dmax2 * dmax2);
  sw->d = 1 / (c * dmax2 * dmax2);
  sw->c = c * dmax2;
}

/* DHEnergy::DHEnergy, colvar/DHEnergy.cpp:104-128, default units (energy kJ/mol, length nm, charge e) */
void orc_dhenergy_setup(orc_switch* sw, double I, double T, double epsilon) {
  memset(sw, 0, sizeof(*sw));
  sw->type = ORC_PAIR_DHENERGY;
  sw->lambda = 138.935458111;                              /* constant */
  sw->beta = sqrt(I / (epsilon * T)) * 502.903741125;      /* k */
  sw->ref = epsilon;
}

static size_t coordination_base_calculate(const orc_nl* nl, const orc_pbc* pbc, int do_pbc, const orc_switch* sw,
                                          const double* pos, const unsigned* abs_index, const double* charges,
                                          const unsigned* types, unsigned ntypes, const double* etas, size_t n,
                                          unsigned rank, unsigned nranks, int nthreads, double* value, double* deriv,
                                          double* virial);

size_t orc_ghbfix_calculate(const orc_nl* nl, const orc_pbc* pbc, int do_pbc, const orc_switch* sw, const double* pos,
                            const unsigned* abs_index, const unsigned* types, unsigned ntypes, const double* etas, size_t n,
                            unsigned rank, unsigned nranks, int nthreads, double* value, double* deriv, double* virial) {
  return coordination_base_calculate(nl, pbc, do_pbc, sw, pos, abs_index, NULL, types, ntypes, etas, n, rank, nranks,
                                     nthreads, value, deriv, virial);
}

size_t orc_dhenergy_calculate(const orc_nl* nl, const orc_pbc* pbc, int do_pbc, const orc_switch* sw, const double* pos,
                              const unsigned* abs_index, const double* charges, size_t n, unsigned rank, unsigned nranks,
                              int nthreads, double* value, double* deriv, double* virial) {
  return coordination_base_calculate(nl, pbc, do_pbc, sw, pos, abs_index, charges, NULL, 0, NULL, n, rank, nranks, nthreads,
                                     value, deriv, virial);
}

size_t orc_coordination_calculate(const orc_nl* nl, const orc_pbc* pbc, int do_pbc, const orc_switch* sw,
                                  const double* pos, const unsigned* abs_index, size_t n, unsigned rank,
                                  unsigned nranks, int nthreads, double* value, double* deriv, double* virial) {
  return coordination_base_calculate(nl, pbc, do_pbc, sw, pos, abs_index, NULL, NULL, 0, NULL, n, rank, nranks, nthreads,
                                     value, deriv, virial);
}

static size_t coordination_base_calculate(const orc_nl* nl, const orc_pbc* pbc, int do_pbc, const orc_switch* sw,
                                          const double* pos, const unsigned* abs_index, const double* charges,
                                          const unsigned* types, unsigned ntypes, const double* etas, size_t n,
                                          unsigned rank, unsigned nranks, int nthreads, double* value, double* deriv,
                                          double* virial) {
  double ncoord = 0.;
  memset(deriv, 0, sizeof(double) * 3 * n);
  memset(virial, 0, sizeof(double) * 9);
  const size_t nn = orc_nl_size(nl);
  unsigned stride = nranks ? nranks : 1;
  int nt = nthreads > 0 ? nthreads : 1;
  if ((size_t)nt * stride * 10 > nn) nt = 1; /* :164-166 */
  const size_t perRank = (size_t)ceil((double)nn / stride); /* :168-170 */
  const size_t start = rank * perRank;
  const size_t end = ((start + perRank) < nn) ? (start + perRank) : nn;
  const int on_the_fly = !nl->list_built;

  double* tderiv = NULL;
  double* tvir = NULL;
  if (nt > 1) {
    tderiv = (double*)calloc((size_t)nt * 3 * n, sizeof(double));
    tvir = (double*)calloc((size_t)nt * 9, sizeof(double));
  }
#pragma omp parallel num_threads(nt) reduction(+ : ncoord)
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    double* mderiv = (nt > 1) ? tderiv + (size_t)t * 3 * n : deriv;
    double* mvir = (nt > 1) ? tvir + (size_t)t * 9 : virial;
#pragma omp for schedule(static) nowait
    for (long long ii = (long long)start; ii < (long long)end; ++ii) {
      unsigned i0, i1;
      if (on_the_fly)
        orc_nl_index_pair(nl, (size_t)ii, &i0, &i1);
      else {
        i0 = nl->pairs[2 * ii];
        i1 = nl->pairs[2 * ii + 1];
      }
      if (abs_index[i0] == abs_index[i1]) continue; /* :183 */
      double distance[3];
      if (do_pbc)
        orc_pbc_distance(pbc, pos + 3 * (size_t)i0, pos + 3 * (size_t)i1, distance);
      else
        for (int k = 0; k < 3; k++) distance[k] = pos[3 * (size_t)i1 + k] - pos[3 * (size_t)i0 + k];
      double dfunc = 0.;
      if (sw->type == ORC_PAIR_GHBFIX) /* GHBFIX::pairing: scale = etas[n*t1+t2], t1 = type of i0 (:189-197) */
        ncoord += ghbfix_pairing(sw, mod2(distance), etas[(size_t)ntypes * types[i0] + types[i1]], &dfunc);
      else if (sw->type == ORC_PAIR_DHENERGY)
        ncoord += dhenergy_pairing(sw, mod2(distance), charges[i0], charges[i1], &dfunc); /* DHEnergy::pairing */
      else
        ncoord += orc_switch_calculate_sqr(sw, mod2(distance), &dfunc); /* Coordination::pairing */
      double dd[3] = {dfunc * distance[0], dfunc * distance[1], dfunc * distance[2]};
      for (int a = 0; a < 3; a++) {
        mderiv[3 * (size_t)i0 + a] -= dd[a];
        mderiv[3 * (size_t)i1 + a] += dd[a];
        for (int b = 0; b < 3; b++) mvir[3 * a + b] -= dd[a] * distance[b]; /* Tensor(dd,distance) */
      }
    }
  }
  if (nt > 1) {
    for (int t = 0; t < nt; t++) {
      const double* md = tderiv + (size_t)t * 3 * n;
      for (size_t k = 0; k < 3 * n; k++) deriv[k] += md[k];
      for (int k = 0; k < 9; k++) virial[k] += tvir[9 * t + k];
    }
    free(tderiv);
    free(tvir);
  }
  *value = ncoord;
  return end > start ? end - start : 0;
}
