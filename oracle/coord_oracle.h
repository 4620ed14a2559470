/* TEST INFRASTRUCTURE ONLY -- the CPU oracle for plumed2_b200.
 *
 * A plain-C restatement of the reference's (plumed/plumed2 v2.11.0-dev) COORDINATION hot path:
 *   src/colvar/CoordinationBase.cpp, src/colvar/Coordination.cpp, src/tools/SwitchingFunction.cpp,
 *   src/tools/Pbc.cpp, src/tools/LatticeReduction.cpp, src/tools/NeighborList.cpp,
 *   src/tools/LinkCells.cpp, src/tools/Tools.h (pbc, fastpow).
 * Every function in coord_oracle.c cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (plumed2_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against golden
 * vectors produced by the real reference (oracle/_ref, built by oracle/Makefile from
 * /root/reference) and against the reference's own regtest fixtures (rt42*, rt-make-switch,
 * rt-Neigbourlist, rt-make-CellLists) re-stated in tests/golden/.
 */
#ifndef COORD_ORACLE_H
#define COORD_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- switching functions: reference enum switchContainers::switchType, SwitchingFunction.h:36-57 */
enum {
  ORC_SW_RATIONALFIX12 = 0, ORC_SW_RATIONALFIX10, ORC_SW_RATIONALFIX8, ORC_SW_RATIONALFIX6,
  ORC_SW_RATIONALFIX4, ORC_SW_RATIONALFIX2,
  ORC_SW_RATIONAL, ORC_SW_RATIONALFAST, ORC_SW_RATIONALSIMPLE, ORC_SW_RATIONALSIMPLEFAST,
  ORC_SW_EXPONENTIAL, ORC_SW_GAUSSIAN, ORC_SW_FASTGAUSSIAN, ORC_SW_SMAP, ORC_SW_CUBIC,
  ORC_SW_TANH, ORC_SW_COSINUS, ORC_SW_NATIVEQ, ORC_SW_LEPTON, ORC_SW_NOT_INITIALIZED
};

/* mirrors switchContainers::Data, SwitchingFunction.h:58-94 */
typedef struct {
  int type;
  double d0, dmax, dmax_2, invr0, invr0_2, stretch, shift;
  int nn, mm;
  double preRes, preDfunc, preSecDev;
  int nnf, mmf;
  double preDfuncF, preSecDevF;
  int a, b;
  double c, d;
  double beta, lambda, ref;
} orc_switch;

/* SwitchingFunction::set(string), SwitchingFunction.cpp:1055-1159. returns 0 ok, else err filled */
int orc_switch_set(orc_switch* sw, const char* definition, char* err, int errlen);
/* SwitchingFunction::set(nn,mm,r0,d0), SwitchingFunction.cpp:1176-1184 */
void orc_switch_set_rational(orc_switch* sw, int nn, int mm, double r0, double d0);
double orc_switch_calculate(const orc_switch* sw, double r, double* dfunc);
double orc_switch_calculate_sqr(const orc_switch* sw, double r2, double* dfunc);

/* ---- Pbc: Pbc.h:52 type enum, Pbc.cpp:165-212 setBox */
enum { ORC_PBC_UNSET = 0, ORC_PBC_ORTHO = 1, ORC_PBC_GENERIC = 2 };
#define ORC_MAXSHIFT 6 /* Pbc.h:58 */
typedef struct {
  int type;
  double box[9], invBox[9], reduced[9], invReduced[9];
  int nshift[8];                      /* index = 4*(s0>0)+2*(s1>0)+(s2>0) */
  double shifts[8][ORC_MAXSHIFT][3];
} orc_pbc;

double orc_tools_pbc(double x);                           /* Tools.h:545-571 */
void orc_lattice_reduce(double t[9]);                     /* LatticeReduction.cpp:144-192 */
void orc_pbc_set_box(orc_pbc* p, const double box[9]);    /* Pbc.cpp:165-212 */
void orc_pbc_distance(const orc_pbc* p, const double v1[3], const double v2[3], double d[3]); /* Pbc.cpp:362-415 */
void orc_pbc_full_search(const orc_pbc* p, double d[3]);  /* Pbc.cpp:137-163 */

/* ---- LinkCells: LinkCells.cpp */
typedef struct {
  int nopbc;           /* bounding-box mode (no box set) */
  double cutoff;
  double origin[3];
  orc_pbc mypbc;
  unsigned ncells[3], nstride[3];
} orc_linkcells;
/* LinkCells::setupCells(pos,pbc), LinkCells.cpp:49-97 + createCells :99-122 */
void orc_linkcells_setup(orc_linkcells* lc, double cutoff, const double* pos, size_t n, const orc_pbc* pbc);
unsigned orc_linkcells_find_cell(const orc_linkcells* lc, const double pos[3]);      /* :277-292,:313-315 */
/* addRequiredCells, :195-239 : returns number of cells written (<=27) */
unsigned orc_linkcells_required(const orc_linkcells* lc, const unsigned celn[3], int use_pbc, unsigned* out);

/* ---- NeighborList: NeighborList.cpp */
enum { ORC_NL_PAIR = 0, ORC_NL_TWOLIST = 1, ORC_NL_SINGLELIST = 2 };
typedef struct {
  int style, do_pbc, use_cells;
  unsigned nlist0, nlist1;
  double cutoff;
  unsigned stride;
  size_t nallpairs;
  int list_built;
  unsigned* pairs;     /* 2*npairs entries: (i0,i1) */
  size_t npairs, cap;
} orc_nl;

orc_nl* orc_nl_create(int style, unsigned n0, unsigned n1, int do_pbc, int use_cells, double cutoff, unsigned stride);
void orc_nl_free(orc_nl* nl);
void orc_nl_index_pair(const orc_nl* nl, size_t ipair, unsigned* i0, unsigned* i1);  /* :147-166 (64-bit safe) */
void orc_nl_update(orc_nl* nl, const orc_pbc* pbc, const double* pos);                /* :168-315 */
/* same pair SET as the classic branch of orc_nl_update but found through a cell grid (O(N));
 * NOT a reference algorithm -- validated against orc_nl_update in tests, used for large-N parity */
void orc_nl_update_classic_cells(orc_nl* nl, const orc_pbc* pbc, const double* pos);
size_t orc_nl_size(const orc_nl* nl);                                                 /* :369-375 */
const unsigned* orc_nl_pairs(const orc_nl* nl);
/* NeighborList::prepare schedule :433-456 ; state in/out: firsttime, invalidate */
void orc_nl_prepare(const orc_nl* nl, long step, int exchange_step, int* firsttime, int* invalidate);

/* ---- CoordinationBase::calculate, CoordinationBase.cpp:142-232.
 * pos: n*3 AoS; abs_index: n absolute atom indices (self-pair skip :183); rank/nranks: MPI stride split
 * :152-170 (results are this rank's partial sums; nranks=1 -> full result); nthreads: OpenMP threads.
 * deriv: n*3, virial: 9 (row-major), value: 1.  Returns pairs iterated. */
/* DHENERGY (colvar/DHEnergy.cpp): the same CoordinationBase loop with the Debye-Hueckel pairing; k and constant as
 * computed by its constructor (:119-120), charges per requested atom (GROUPA then GROUPB) */
#define ORC_PAIR_GHBFIX 33
/* GHBFIX (colvar/GHBFIX.cpp): constants of the constructor (:98-113); types = typesTable[absolute index] per requested
 * atom, etas = ntypes x ntypes table in PLUMED energy units */
void orc_ghbfix_setup(orc_switch* sw, double dmax, double d0, double c);
size_t orc_ghbfix_calculate(const orc_nl* nl, const orc_pbc* pbc, int do_pbc, const orc_switch* sw, const double* pos,
                            const unsigned* abs_index, const unsigned* types, unsigned ntypes, const double* etas, size_t n,
                            unsigned rank, unsigned nranks, int nthreads, double* value, double* deriv, double* virial);
#define ORC_PAIR_DHENERGY 32
void orc_dhenergy_setup(orc_switch* sw, double I, double T, double epsilon); /* default PLUMED units */
size_t orc_dhenergy_calculate(const orc_nl* nl, const orc_pbc* pbc, int do_pbc, const orc_switch* sw, const double* pos,
                              const unsigned* abs_index, const double* charges, size_t n, unsigned rank, unsigned nranks,
                              int nthreads, double* value, double* deriv, double* virial);
size_t orc_coordination_calculate(const orc_nl* nl, const orc_pbc* pbc, int do_pbc, const orc_switch* sw,
                                  const double* pos, const unsigned* abs_index, size_t n,
                                  unsigned rank, unsigned nranks, int nthreads,
                                  double* value, double* deriv, double* virial);

#ifdef __cplusplus
}
#endif
#endif
