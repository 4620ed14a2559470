/*
 * b200coord.h -- C ABI of the B200-native COORDINATION + neighbour-list engine (libb200coord.so).
 *
 * This is the drop-in boundary: a PLUMED action (plumed2_b200/csrc/plugin/CoordinationB200.cpp,
 * registered with PLUMED_REGISTER_ACTION as "COORDINATION") or any other host (the ctypes mirror in
 * plumed2_b200/coordination.py, an MD engine) calls only these functions.  Plain C types, caller-owned
 * host arrays, library-owned device memory.  Every function returns 0 on success and a non-zero
 * B200COORD_ERR_* code otherwise -- it never throws and never falls back to a CPU path: if no usable
 * CUDA device / kernel image is present the call fails and b200coord_last_error() says why.
 *
 * Reference interfaces replaced (paths relative to plumed2 v2.11.0-dev, /root/reference):
 *   b200coord_switch_parse / _rational   <- SwitchingFunction::set            src/tools/SwitchingFunction.cpp:1055-1159, :1176-1184
 *   b200coord_create                     <- CoordinationBase ctor + NeighborList ctors
 *                                           src/colvar/CoordinationBase.cpp:42-131, src/tools/NeighborList.cpp:43-141
 *   b200coord_set_box                    <- Pbc::setBox                       src/tools/Pbc.cpp:165-212
 *   b200coord_prepare                    <- CoordinationBase::prepare -> NeighborList::prepare
 *                                           src/colvar/CoordinationBase.cpp:137-139, src/tools/NeighborList.cpp:433-456
 *   b200coord_update_list                <- NeighborList::update              src/tools/NeighborList.cpp:168-315
 *   b200coord_calculate                  <- CoordinationBase::calculate       src/colvar/CoordinationBase.cpp:142-232
 *   b200coord_comm_*                     <- Communicator::Sum (MPI_Allreduce) src/tools/Communicator.cpp:194-201,
 *                                           called at src/colvar/CoordinationBase.cpp:218-224
 *   b200coord_nl_size / _nl_pairs        <- NeighborList::size / getClosePair src/tools/NeighborList.cpp:369-389
 */
#ifndef B200COORD_H
#define B200COORD_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200COORD_ABI_VERSION 1

/* error codes */
enum {
  B200COORD_OK = 0,
  B200COORD_ERR_INVALID = 1,   /* bad argument / unsupported keyword combination */
  B200COORD_ERR_CUDA = 2,      /* CUDA runtime error (no device, launch failure, OOM) */
  B200COORD_ERR_PARSE = 3,     /* SWITCH string could not be parsed */
  B200COORD_ERR_UNSUPPORTED = 4, /* valid in the reference but not available on the GPU (SWITCH=CUSTOM) */
  B200COORD_ERR_NCCL = 5,
  B200COORD_ERR_STATE = 6      /* call order violated (e.g. calculate before set_box with PBC) */
};

/* pair-list style, reference NeighborList::NNStyle (src/tools/NeighborList.h:44) */
enum { B200COORD_STYLE_PAIR = 0, B200COORD_STYLE_TWOLIST = 1, B200COORD_STYLE_SINGLELIST = 2 };

/* neighbour-list mode: none (all pairs, NL off), NLIST (distance filtered), NLISTCELLS (27-cell superset) */
enum { B200COORD_NL_NONE = 0, B200COORD_NL_CLASSIC = 1, B200COORD_NL_CELLS = 2 };

/* arithmetic of the pair sweep: FP64 (default, 1e-10 parity) or the opt-in FP32 mode (1e-5 parity): coordinate
 * difference and minimum image in FP64, then r^2, switching function, dd and the sums over one row in FP32, rows
 * added to FP64 accumulators; no exact patch for pairs within rounding of D_MAX / D_0.  The neighbour list is the
 * same bit-exact list in both modes.  PAIR style always runs its FP64 kernel. */
enum { B200COORD_FP64 = 0, B200COORD_FP32 = 1 };

/* switching-function kinds, same order as switchContainers::switchType (src/tools/SwitchingFunction.h:36-57) */
enum {
  B200COORD_SW_RATIONALFIX12 = 0, B200COORD_SW_RATIONALFIX10, B200COORD_SW_RATIONALFIX8,
  B200COORD_SW_RATIONALFIX6, B200COORD_SW_RATIONALFIX4, B200COORD_SW_RATIONALFIX2,
  B200COORD_SW_RATIONAL, B200COORD_SW_RATIONALFAST, B200COORD_SW_RATIONALSIMPLE, B200COORD_SW_RATIONALSIMPLEFAST,
  B200COORD_SW_EXPONENTIAL, B200COORD_SW_GAUSSIAN, B200COORD_SW_FASTGAUSSIAN, B200COORD_SW_SMAP,
  B200COORD_SW_CUBIC, B200COORD_SW_TANH, B200COORD_SW_COSINUS, B200COORD_SW_NATIVEQ,
  B200COORD_SW_LEPTON, B200COORD_SW_NOT_INITIALIZED
};

/* pairing functions of COORDINATION's siblings on CoordinationBase, carried in the same struct:
 * DHENERGY (src/colvar/DHEnergy.cpp:130-143): s = exp(-k r)/r * constant/epsilon * q_i q_j, dfunc = -(k + 1/r) s / r;
 * beta = k, lambda = constant/epsilon, no D_MAX; needs b200coord_set_charges */
enum { B200COORD_PAIR_DHENERGY = 32, B200COORD_PAIR_GHBFIX = 33 };
/* GHBFIX (src/colvar/GHBFIX.cpp:186-220): typed piecewise polynomial, s = eta(t_i0,t_i1) * f(r), zero beyond D_MAX;
 * d0, dmax, dmax_2 as named; preRes = A, preDfunc = B, preSecDev = C, d = D, c = C_keyword*(D_MAX-D_0) (the joint of
 * the two polynomials), GHBFIX.cpp:108-113; needs b200coord_set_types */

/* mirrors switchContainers::Data (src/tools/SwitchingFunction.h:58-94) */
typedef struct b200coord_switch {
  int type;
  double d0, dmax, dmax_2, invr0, invr0_2, stretch, shift;
  int nn, mm;
  double preRes, preDfunc, preSecDev;
  int nnf, mmf;
  double preDfuncF, preSecDevF;
  int a, b;
  double c, d;
  double beta, lambda, ref;
} b200coord_switch;

typedef struct b200coord_config {
  int abi_version;       /* must be B200COORD_ABI_VERSION */
  int device;            /* CUDA device ordinal, -1 = the calling thread's current device */
  int precision;         /* B200COORD_FP64 | B200COORD_FP32 */
  int style;             /* B200COORD_STYLE_* ; SINGLELIST when GROUPB is empty, PAIR for the PAIR flag */
  unsigned n_group_a;    /* atoms in GROUPA */
  unsigned n_group_b;    /* atoms in GROUPB (0 for SINGLELIST) */
  int pbc;               /* 1 unless NOPBC */
  int nl_mode;           /* B200COORD_NL_* */
  double nl_cutoff;      /* NL_CUTOFF (>0 when nl_mode != NONE) */
  int nl_stride;         /* NL_STRIDE (>0 when nl_mode != NONE) */
  int rank;              /* this context's share of the i-atoms, like the MPI stride split ... */
  int nranks;            /* ... CoordinationBase.cpp:152-170 ; 0/1 = everything */
} b200coord_config;

typedef struct b200coord_stats {
  unsigned long long nl_size;        /* list entries the reference would iterate (nl->size()), self-pairs included */
  unsigned long long pair_evals;     /* pair evaluations this context executed in the last calculate (both directions) */
  unsigned long long kernel_launches;/* kernels launched by this context since creation */
  unsigned long long rebuilds;       /* neighbour-list rebuilds since creation */
  float last_sweep_ms;               /* CUDA-event time of the last pair-sweep kernel */
  float last_build_ms;               /* CUDA-event time of the last list rebuild (all its kernels) */
  float last_h2d_ms, last_d2h_ms;    /* last host<->device copies inside calculate */
  unsigned ncells[3];                /* cell grid in use */
  int pbc_type;                      /* 0 unset, 1 orthorhombic, 2 generic (Pbc.h:52) */
  float sweep_ms_sum;                /* CUDA-event time of all pair sweeps since the last b200coord_stream_mark(ctx,0) ... */
  unsigned sweep_count;              /* ... and how many sweeps that covers (at most the last 64 are kept) */
  float build_ms_sum;                /* same for list rebuilds */
  unsigned build_count;
  int f32_search;                    /* 1: the last rebuild used the FP32 candidate search (+ exact FP64 band) */
  unsigned long long super_builds;   /* rebuilds that (re)built the super-list (cutoff + 10 %) from the cells */
  unsigned long long filter_rebuilds;/* rebuilds that only filtered the super-list (displacement bound held) */
  float build_ms_max;                /* longest single rebuild since the last b200coord_stream_mark(ctx,0) */
} b200coord_stats;

typedef struct b200coord_ctx b200coord_ctx;

int b200coord_abi_version(void);

/* ---- switching function set-up (host only, no CUDA needed) */
/* SWITCH={...} form; err (may be NULL) receives the reference's error text */
int b200coord_switch_parse(const char* definition, b200coord_switch* out, char* err, size_t errlen);
/* R_0= NN= MM= D_0= keyword form: automatic D_MAX and stretch (SwitchingFunction.cpp:1176-1184) */
int b200coord_switch_rational(int nn, int mm, double r0, double d0, b200coord_switch* out);
/* DHENERGY constants from its keywords I, TEMP, EPSILON and the unit factors of the host code (all 1 for PLUMED's
 * default kJ/mol, nm, e), DHEnergy.cpp:104-128 */
int b200coord_pairing_dhenergy(double ionic_strength, double temp, double epsilon, double energy_unit, double length_unit,
                               double charge_unit, b200coord_switch* out);
/* GHBFIX polynomial constants from its keywords D_MAX, D_0, C (GHBFIX.cpp:98-113) */
int b200coord_pairing_ghbfix(double dmax, double d0, double c, b200coord_switch* out);
/* text like SwitchingFunction::description() for the action's log */
int b200coord_switch_describe(const b200coord_switch* sw, char* buf, size_t buflen);

/* ---- life cycle */
/* abs_index: n_group_a+n_group_b absolute atom indices in GROUPA-then-GROUPB order (self-pair skip by
 * absolute index, CoordinationBase.cpp:183) */
int b200coord_create(const b200coord_config* cfg, const b200coord_switch* sw, const unsigned* abs_index,
                     b200coord_ctx** out);
/* charges of the n_group_a+n_group_b atoms in GROUPA-then-GROUPB order (ActionAtomistic::getCharge, used by
 * DHEnergy::pairing); required before calculate when the pairing is B200COORD_PAIR_DHENERGY; may be called again
 * whenever the charges change */
int b200coord_set_charges(b200coord_ctx* ctx, const double* charges);
/* GHBFIX: interaction type of each of the n_group_a+n_group_b atoms (typesTable[absolute index], GHBFIX.cpp:117-140)
 * and the ntypes x ntypes scaling table etas (row = type of the pair's first atom, :142-163), already in PLUMED energy
 * units; required before calculate when the pairing is B200COORD_PAIR_GHBFIX */
int b200coord_set_types(b200coord_ctx* ctx, const unsigned* types, unsigned ntypes, const double* etas);
void b200coord_destroy(b200coord_ctx* ctx);
/* last error text of this context (ctx==NULL: of the last failed create/parse in this thread) */
const char* b200coord_last_error(const b200coord_ctx* ctx);

/* ---- per step */
/* box: 9 doubles row-major, box[3*i+j] = j-th component of lattice vector i; all zeros = no box */
int b200coord_set_box(b200coord_ctx* ctx, const double box[9]);
/* neighbour-list schedule; *will_rebuild = 1 when the next calculate rebuilds the list.
 * Returns B200COORD_ERR_STATE for a non-rebuild exchange step (NeighborList.cpp:447-449) */
int b200coord_prepare(b200coord_ctx* ctx, long step, int exchange_step, int* will_rebuild);
/* force a rebuild from these host positions now (NeighborList::update) */
int b200coord_update_list(b200coord_ctx* ctx, const double* pos);
/* pos: n*3 host doubles (AoS, GROUPA then GROUPB).  Outputs (host): *value, deriv n*3, virial 9 (row-major).
 * With a communicator attached the outputs are the sums over all ranks on every rank. */
int b200coord_calculate(b200coord_ctx* ctx, const double* pos, double* value, double* deriv, double* virial);
/* same with device-resident buffers: d_pos n*3 doubles, d_out 3n+10 doubles = [deriv | virial(9) | value] */
int b200coord_calculate_device(b200coord_ctx* ctx, const double* d_pos, double* d_out);

/* enqueue-only variant for callers that live on the GPU (GPU-resident MD engines, the HBM-resident leg of
 * bench.py): same work on the context's stream, no host synchronisation at the end unless a list rebuild needs
 * to size its buffers.  Results are valid after b200coord_stream_elapsed_ms / b200coord_device_synchronize. */
int b200coord_enqueue_device(b200coord_ctx* ctx, const double* d_pos, double* d_out);
/* CUDA-event stopwatch on the context's own stream (where all its kernels are launched) */
int b200coord_stream_mark(b200coord_ctx* ctx, int which /*0 = start, 1 = stop*/);
int b200coord_stream_elapsed_ms(b200coord_ctx* ctx, float* ms);   /* waits for the stop mark */
/* Distributed step (needs b200coord_comm_init): this rank uploads only its slice of the positions
 * [slot_begin, slot_begin+slot_count) (equal chunks of ceil(n/nranks) slots, rank-ordered), positions are
 * all-gathered over NVLink, every rank sweeps its own i-atoms, derivative rows are all-gathered and
 * value/virial all-reduced (the Comm::Sum of CoordinationBase.cpp:218-224), and only this rank's slice of
 * the derivatives is copied back to deriv_slice (slot_count*3 doubles). */
int b200coord_calculate_distributed(b200coord_ctx* ctx, const double* pos_slice, double* value, double* deriv_slice,
                                    double* virial);
int b200coord_my_slice(const b200coord_ctx* ctx, unsigned* slot_begin, unsigned* slot_count);
/* enqueue-only distributed step for callers whose positions already live on the GPU: d_pos = the WHOLE position array
 * on this rank's device, d_out_slice = 3*slot_count+10 doubles on this device = [derivatives of this rank's slots |
 * virial(9) | value].  No host synchronisation unless the list is rebuilt. */
int b200coord_enqueue_device_distributed(b200coord_ctx* ctx, const double* d_pos, double* d_out_slice);
/* number of CUDA devices visible to this process (cudaGetDeviceCount) */
int b200coord_device_count(int* n);
/* FP64 FMA peak of the device measured with a register-resident DFMA kernel (roofline denominator) */
int b200coord_measure_fp64_peak(int device, double* tflops);

/* ---- inspection */
int b200coord_get_stats(const b200coord_ctx* ctx, b200coord_stats* out);
/* current list as (i0,i1) index pairs into the position array, sorted by (i0,i1); self-pairs included as the
 * reference lists them.  *n receives the number of pairs; pairs may be NULL to query the size only. */
int b200coord_nl_pairs(b200coord_ctx* ctx, unsigned* pairs, unsigned long long capacity, unsigned long long* n);

/* The neighbour list as a tool for other consumers -- what the reference's NeighborList is to ContactMap
 * (src/colvar/ContactMap.cpp:190-260), isdb/NOE.cpp, PRE.cpp, piv/PIV.cpp: the list of this context's rows (NLIST,
 * not PAIR) as (i0,i1) index pairs into the position array, each pair once in the reference's orientation
 * (NeighborList::getClosePair: GROUPA atom first, resp. lower index first), written to DEVICE memory in row order.
 * d_pairs = 2*capacity unsigned on this context's device, or NULL to query the count.  Pairs of one and the same atom
 * (groups that share atoms) are not in the list.  Rebuild with b200coord_update_list or the schedule of prepare(). */
int b200coord_nl_pairs_device(b200coord_ctx* ctx, unsigned* d_pairs, unsigned long long capacity, unsigned long long* n);

/* ---- multi-GPU (one context per GPU/process; i-atoms sharded by cfg.rank/nranks) */
#define B200COORD_UNIQUE_ID_BYTES 128
int b200coord_comm_unique_id(char id[B200COORD_UNIQUE_ID_BYTES]);          /* rank 0, then broadcast by the host */
int b200coord_comm_init(b200coord_ctx* ctx, const char id[B200COORD_UNIQUE_ID_BYTES]); /* every rank */

/* Exchange over NVLink peer memory (optional, after b200coord_comm_init): every rank exports IPC handles of its
 * derivative-row buffers and of its position-slice buffers, the host gathers the handles of all ranks (rank order)
 * and hands them back.  From then on nothing is broadcast: a rank's derivative rows stay in its own buffer and every
 * rank PULLS the rows of the atoms it returns (24 bytes each); on steps that keep the neighbour list a rank uploads
 * only its slice of the positions and the gather of every rank pulls the positions around its own rows from the
 * owners.  What remains of NCCL per step is the 10-double all-reduce of virial/value (which also orders the ranks)
 * and a one-element all-reduce as a barrier after the uploads; rebuild steps still all-gather the positions. */
#define B200COORD_PEER_HANDLE_BYTES 256
int b200coord_peer_export(b200coord_ctx* ctx, char handle[B200COORD_PEER_HANDLE_BYTES]);
int b200coord_peer_attach(b200coord_ctx* ctx, const char* all_handles /* nranks * B200COORD_PEER_HANDLE_BYTES */);

/* the same for contexts that live in ONE process (no IPC): all[r] = the context of rank r, after every context has
 * called b200coord_comm_init and b200coord_peer_export */
int b200coord_peer_attach_local(b200coord_ctx* ctx, b200coord_ctx* const* all, int n);

/* ---- several GPUs inside one process (a `plumed driver` run, an MD engine without MPI): one context per device, one
 * worker thread per context, the i-atoms sharded over them exactly like MPI ranks (CoordinationBase.cpp:152-170), the
 * Comm::Sum of :218-224 done by NCCL + NVLink peer memory as above.  Same call sequence as a single context; the
 * positions / derivatives are the whole host arrays, every device uploads and returns its own slice of them.
 * cfg->device, cfg->rank and cfg->nranks are ignored.  PAIR style: one device only. */
typedef struct b200coord_group b200coord_group;
int b200coord_group_create(const b200coord_config* cfg, const b200coord_switch* sw, const unsigned* abs_index,
                           const int* devices, int ndevices, b200coord_group** out);
void b200coord_group_destroy(b200coord_group* g);
int b200coord_group_size(const b200coord_group* g);
b200coord_ctx* b200coord_group_context(b200coord_group* g, int rank);   /* for b200coord_get_stats etc. */
const char* b200coord_group_last_error(const b200coord_group* g);
int b200coord_group_set_box(b200coord_group* g, const double box[9]);
int b200coord_group_prepare(b200coord_group* g, long step, int exchange_step, int* will_rebuild);
int b200coord_group_set_charges(b200coord_group* g, const double* charges);
int b200coord_group_set_types(b200coord_group* g, const unsigned* types, unsigned ntypes, const double* etas);
int b200coord_group_calculate(b200coord_group* g, const double* pos, double* value, double* deriv, double* virial);

/* ---- frames known in advance (trajectory post-processing; plumed driver, src/cltools/Driver.cpp): two steps in flight,
   the upload of the next frame and the download of the previous result overlap the sweep of the current one.
   b200coord_prepare(step) before every submit as before a calculate.  `deriv` (page-locked for a real overlap) is
   complete, and *value / virial[9] are written, when the second submit after this one or b200coord_collect returns. */
int b200coord_submit(b200coord_ctx* ctx, const double* pos, double* value, double* deriv, double* virial);
int b200coord_collect(b200coord_ctx* ctx);

/* ---- an MD engine that keeps positions and forces on the GPU. The reference's engines hand host
   pointers to plumed_cmd (patches/gromacs-2025.0.diff/src/gromacs/applied_forces/plumed/plumedforceprovider.cpp:171-204)
   and PLUMED copies them in and out (src/core/ActionAtomistic.cpp:444-537). Here the engine publishes its device
   arrays (double[3*natoms], PLUMED's internal units) under a name; the action that carries GPU_COUPLING=<name> reads the
   positions where they are and adds (force on the CV) x dCV/dx to the force array -- Colvar::apply
   (src/core/Colvar.cpp:50-60) on the device. Value and virial are returned on the host, derivatives stay on the device.
   The positions must be complete when calculate_coupled is called; the forces are complete when apply_coupled returns. */
int b200coord_coupling_publish(const char* key, int device, const double* d_pos, double* d_force, size_t natoms);
int b200coord_coupling_withdraw(const char* key);
int b200coord_coupling_lookup(const char* key, int* device, const double** d_pos, double** d_force, size_t* natoms);
/* index[i] = where atom i of the context (GROUPA then GROUPB) sits in the engine's arrays; NULL: atom i is atom i */
int b200coord_coupled_set_index(b200coord_ctx* ctx, const unsigned* index);
int b200coord_calculate_coupled(b200coord_ctx* ctx, const double* d_pos_all, double* value, double* virial);
int b200coord_apply_coupled(b200coord_ctx* ctx, double factor, double* d_force_all);
int b200coord_coupled_derivatives(b200coord_ctx* ctx, double* deriv /* host, 3n: for tests and DUMPDERIVATIVES-like needs */);

/* ---- pinned host memory for callers that want full-speed copies */
int b200coord_host_alloc(size_t bytes, void** ptr);
int b200coord_host_free(void* ptr);
/* raw device scratch for benchmarks/tests that keep inputs resident in HBM */
int b200coord_device_alloc(size_t bytes, void** dptr);
int b200coord_device_free(void* dptr);
int b200coord_memcpy_h2d(void* dptr, const void* hptr, size_t bytes);
int b200coord_memcpy_d2h(void* hptr, const void* dptr, size_t bytes);
int b200coord_device_synchronize(void);

#ifdef __cplusplus
}
#endif
#endif /* B200COORD_H */
