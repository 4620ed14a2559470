// The PLUMED-facing side of the drop-in: a Colvar registered under the key "COORDINATION".
//
// Loaded with `LOAD FILE=libb200coord_plumed.so` at the top of plumed.dat it overrides the built-in
// COORDINATION for every later input line (src/core/RegisterBase.h:236-252, src/setup/Load.cpp:86-135),
// for `plumed driver`, `plumed benchmark` and any MD engine talking plumed_cmd.  It keeps the keyword set
// of src/colvar/CoordinationBase.cpp:30-40 and src/colvar/Coordination.cpp:115-126 (plus the additive
// top-level D_MAX of plugins/cudaCoord), the prepare()/calculate() life cycle and the error texts, and does
// every computation through the C ABI of libb200coord.so (include/b200coord.h).  No pair is ever evaluated
// on the host; if the CUDA library reports an error the action calls error().
#include "b200coord.h"
#include "core/ActionRegister.h"
#include "core/ActionSet.h"
#include "core/ActionToPutData.h"
#include "core/Colvar.h"
#include "core/PlumedMain.h"
#include "tools/Communicator.h"
#include "tools/IFile.h"
#include "tools/Units.h"
#include "tools/OpenMP.h"
#include "tools/Pbc.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace PLMD {
namespace colvar {

// Everything of CoordinationBase (src/colvar/CoordinationBase.cpp) lives here; the two registered actions only differ
// in the keywords of their pairing function, as Coordination.cpp / DHEnergy.cpp do in the reference.
class CoordinationBaseB200 : public Colvar {
protected:
  b200coord_ctx* ctx = nullptr;
  b200coord_group* group = nullptr;  // GPU_DEVICES=0,1,...: one context per device inside this process
  bool needCharges = false;
  bool pbc = true;
  bool serial = false;
  bool combineWithMpi = false;
  // page-locked when the library can give it (copies at link speed, no staging), ordinary memory otherwise
  struct HostArray {
    double* p = nullptr;
    std::size_t n = 0;
    bool pinned = false;
    std::vector<double> plain;
    void resize(std::size_t count) {
      release();
      void* q = nullptr;
      if (count >= 4096 && b200coord_host_alloc(count * sizeof(double), &q) == B200COORD_OK && q) {
        p = static_cast<double*>(q);
        pinned = true;
        std::fill(p, p + count, 0.0);
      } else {
        plain.assign(count, 0.0);
        p = plain.data();
      }
      n = count;
    }
    void release() {
      if (pinned) {
        b200coord_host_free(p);
      }
      plain.clear();
      p = nullptr;
      n = 0;
      pinned = false;
    }
    double* data() { return p; }
    ~HostArray() { release(); }
  };
  HostArray derivBuffer;
  std::vector<double> chargeBuffer, chargeSent;
  // B200COORD_PLUGIN_TIMERS=1: seconds spent in the C-ABI call vs in handing the result to PLUMED's Value
  bool timers = false;
  double tEngine = 0.0, tStore = 0.0;
  unsigned long nCalls = 0;
  void check(int rc, const char* what);
  // Host side of large systems.  PLUMED gathers the requested atoms (ActionAtomistic::retrieveAtoms) and scatters the
  // forces (ActionWithValue::checkForForces + ActionAtomistic::setForcesOnAtoms) with serial loops: 6 ms + 8 ms per step
  // at 1 M atoms, against a 1.5 ms GPU step.  When nothing else needs those arrays (no charges, no numerical derivatives,
  // no virtual atoms, no atom listed twice) the action does both itself with OpenMP: it tells PLUMED not to retrieve
  // (ActionAtomistic::doNotRetrieve, core/ActionAtomistic.h:169-176), reads the shared position values through
  // getGlobalPosition() and adds f * derivative to the force arrays of posx / posy / posz with Value::addForce.
  std::string coupling;  // GPU_COUPLING=<name>: positions and forces live in the engine's device arrays
  bool fastHost = false;
  std::vector<std::pair<std::size_t, std::size_t>> valueIdx;
  Value* posValue[3] = {nullptr, nullptr, nullptr};
  HostArray posBuffer;
  std::vector<double> forceBuffer[3];  // dense per-component force arrays for Value::addForces
  double lastVirial[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  void setupFastHost();

  // group / list keywords and engine set-up; `sw` is the pairing the derived constructor parsed
  void setup(const b200coord_switch& sw, const char* what);

public:
  explicit CoordinationBaseB200(const ActionOptions& ao) : PLUMED_COLVAR_INIT(ao) {}
  ~CoordinationBaseB200() override;
  static void registerKeywords(Keywords& keys);
  void prepare() override;
  void calculate() override;
  void apply() override;
  void clearDerivatives(const bool& force = false) override;
};

class CoordinationB200 : public CoordinationBaseB200 {
public:
  explicit CoordinationB200(const ActionOptions&);
  static void registerKeywords(Keywords& keys);
};
PLUMED_REGISTER_ACTION(CoordinationB200, "COORDINATION")

// DHENERGY (src/colvar/DHEnergy.cpp): Debye-Hueckel interaction energy between the two groups
class DHEnergyB200 : public CoordinationBaseB200 {
public:
  explicit DHEnergyB200(const ActionOptions&);
  static void registerKeywords(Keywords& keys);
};
PLUMED_REGISTER_ACTION(DHEnergyB200, "DHENERGY")

// GHBFIX (src/colvar/GHBFIX.cpp): typed piecewise-polynomial interaction energy between the two groups
class GHBFIXB200 : public CoordinationBaseB200 {
public:
  explicit GHBFIXB200(const ActionOptions&);
  static void registerKeywords(Keywords& keys);
};
PLUMED_REGISTER_ACTION(GHBFIXB200, "GHBFIX")

void CoordinationBaseB200::registerKeywords(Keywords& keys) {
  Colvar::registerKeywords(keys);
  keys.addFlag("SERIAL", false, "Perform the calculation in serial - for debug purpose");
  keys.addFlag("PAIR", false, "Pair only 1st element of the 1st group with 1st element in the second, etc");
  keys.addFlag("NLIST", false, "Use a neighbor list to speed up the calculation");
  keys.addFlag("NLISTCELLS", false, "Use a neighbor list to speed up the calculation - cell list flavour (27-cell superset)");
  keys.add("optional", "NL_CUTOFF", "The cutoff for the neighbor list");
  keys.add("optional", "NL_STRIDE", "The frequency with which we are updating the atoms in the neighbor list");
  keys.add("atoms", "GROUPA", "First list of atoms");
  keys.add("atoms", "GROUPB", "Second list of atoms (if empty, N*(N-1)/2 pairs in GROUPA are counted)");
  keys.add("optional", "GPU_DEVICE", "CUDA device ordinal to run on (default: the B200COORD_DEVICE environment variable, else the current device)");
  keys.add("optional", "GPU_DEVICES", "comma separated CUDA device ordinals: shard the i-atoms over several GPUs of this node inside this process (NCCL + NVLink peer memory); not with PAIR, not together with several MPI ranks");
  keys.add("optional", "GPU_COUPLING", "name under which the MD engine published device arrays of positions and forces (b200coord_coupling_publish): the action reads positions from the device array and adds its forces to the device array; nothing is copied through the host. The value and the virial are still returned to PLUMED; the derivatives stay on the device");
  keys.addFlag("GPU_FP32", false, "opt-in FP32 pair arithmetic (1e-5 relative instead of 1e-10; FP64 minimum image and accumulation). Also switched on by the B200COORD_FP32=1 environment variable");
}

void CoordinationB200::registerKeywords(Keywords& keys) {
  CoordinationBaseB200::registerKeywords(keys);
  keys.add("compulsory", "NN", "6", "The n parameter of the switching function ");
  keys.add("compulsory", "MM", "0", "The m parameter of the switching function; 0 implies 2*NN");
  keys.add("compulsory", "D_0", "0.0", "The d_0 parameter of the switching function");
  keys.add("compulsory", "R_0", "The r_0 parameter of the switching function");
  keys.add("optional", "D_MAX", "cut the rational switching function built from R_0/NN/MM/D_0 at this distance (stretched to zero)");
  keys.add("optional", "SWITCH", "This keyword is used if you want to employ an alternative to the continuous switching function defined above. "
           "When this keyword is present you no longer need the NN, MM, D_0 and R_0 keywords.");
  keys.setValueDescription("scalar", "the value of the coordination");
}

void DHEnergyB200::registerKeywords(Keywords& keys) {  // DHEnergy.cpp:76-83
  CoordinationBaseB200::registerKeywords(keys);
  keys.add("compulsory", "I", "1.0", "Ionic strength (M)");
  keys.add("compulsory", "TEMP", "300.0", "Simulation temperature (K)");
  keys.add("compulsory", "EPSILON", "80.0", "Dielectric constant of solvent");
  keys.setValueDescription("scalar", "the value of the DHENERGY");
}

void GHBFIXB200::registerKeywords(Keywords& keys) {  // GHBFIX.cpp:91-103
  CoordinationBaseB200::registerKeywords(keys);
  keys.add("optional", "ENERGY_UNITS", "the value of ENERGY_UNITS in the switching function");
  keys.add("compulsory", "TYPES", "the value of TYPES in the switching function");
  keys.add("compulsory", "PARAMS", "the value of PARAMS in the switching function");
  keys.add("compulsory", "D_MAX", "the value of D_MAX in the switching function");
  keys.add("compulsory", "D_0", "the value of D_0 in the switching function");
  keys.add("compulsory", "C", "the value of C in the switching function");
  keys.setValueDescription("scalar", "the GHBFIX interaction energy between the atoms in GROUPA and GROUPB");
}

void CoordinationBaseB200::check(int rc, const char* what) {
  if (rc != B200COORD_OK) {
    error(std::string(what) + " failed on the GPU: " + (group ? b200coord_group_last_error(group) : b200coord_last_error(ctx)));
  }
}

CoordinationB200::CoordinationB200(const ActionOptions& ao) : Action(ao), CoordinationBaseB200(ao) {
  b200coord_switch sw;
  std::string swDef;
  parse("SWITCH", swDef);
  if (!swDef.empty()) {
    char err[1024];
    const int rc = b200coord_switch_parse(swDef.c_str(), &sw, err, sizeof(err));
    if (rc != B200COORD_OK) {
      error("problem reading SWITCH keyword : " + std::string(err));
    }
  } else {
    int nn = 6, mm = 0;
    double d0 = 0.0, r0 = 0.0, dmax = -1.0;
    parse("R_0", r0);
    if (r0 <= 0.0) {
      error("R_0 should be explicitly specified and positive");
    }
    parse("D_0", d0);
    parse("NN", nn);
    parse("MM", mm);
    parse("D_MAX", dmax);
    if (dmax > 0.0) {
      // 17 significant digits: the text round-trips to the same doubles (std::to_string keeps 6 decimals)
      char def[512];
      std::snprintf(def, sizeof(def), "RATIONAL R_0=%.17g D_0=%.17g NN=%d MM=%d D_MAX=%.17g", r0, d0, nn, mm, dmax);
      char err[1024];
      if (b200coord_switch_parse(def, &sw, err, sizeof(err)) != B200COORD_OK) {
        error("problem building the switching function : " + std::string(err));
      }
    } else {
      b200coord_switch_rational(nn, mm, r0, d0, &sw);
    }
  }
  setup(sw, "COORDINATION");
}

DHEnergyB200::DHEnergyB200(const ActionOptions& ao) : Action(ao), CoordinationBaseB200(ao) {  // DHEnergy.cpp:104-128
  double I = 1.0, T = 300.0, epsilon = 80.0;
  parse("I", I);
  parse("TEMP", T);
  parse("EPSILON", epsilon);
  if (usingNaturalUnits()) {
    error("DHENERGY cannot be used for calculations performed with natural units");
  }
  b200coord_switch sw;
  if (b200coord_pairing_dhenergy(I, T, epsilon, getUnits().getEnergy(), getUnits().getLength(), getUnits().getCharge(), &sw) !=
      B200COORD_OK) {
    error(b200coord_last_error(nullptr));
  }
  needCharges = true;
  setup(sw, "DHENERGY");
  log << "  with solvent dielectric constant " << epsilon << "\n";
  log << "  at temperature " << T << " K\n";
  log << "  at ionic strength " << I << "M\n";
  log << "  Bibliography " << plumed.cite("Do, Carloni, Varani and Bussi, J. Chem. Theory Comput. 9, 1720 (2013)") << " \n";
}

GHBFIXB200::GHBFIXB200(const ActionOptions& ao) : Action(ao), CoordinationBaseB200(ao) {  // GHBFIX.cpp:94-172
  double dmax = 0.0, d0 = 0.0, c = 0.0;
  std::string types, params, energy_units = "plumed";
  parse("D_MAX", dmax);
  parse("D_0", d0);
  parse("C", c);
  parse("TYPES", types);
  parse("PARAMS", params);
  parse("ENERGY_UNITS", energy_units);
  b200coord_switch sw;
  if (b200coord_pairing_ghbfix(dmax, d0, c, &sw) != B200COORD_OK) {
    error(b200coord_last_error(nullptr));
  }
  // the two tables, read exactly as the reference reads them (std::map::operator[] makes a type name that only the
  // parameter file knows read as index 0)
  std::map<std::string, unsigned> table;
  std::vector<unsigned> typesTable;
  {
    IFile typesfile;
    typesfile.link(*this);
    typesfile.open(types);
    std::string itype;
    while (typesfile.scanField("itype", itype).scanField()) {
      plumed_assert(itype.empty() == false) << "itype is empty";
      if (table.empty()) {
        table.insert({itype, 0});
      } else if (table.count(itype) == 0) {
        unsigned currentMax = 0;
        for (const auto& kv : table) {
          currentMax = std::max(currentMax, kv.second);
        }
        table.insert({itype, currentMax + 1});
      }
      typesTable.push_back(table[itype]);
    }
  }
  plumed_assert(!typesTable.empty()) << "the TYPES file is empty";
  const unsigned nt = *std::max_element(typesTable.begin(), typesTable.end()) + 1;
  std::vector<double> etas(static_cast<std::size_t>(nt) * nt, 0.0);
  {
    IFile etafile;
    etafile.open(params);
    std::string it, jt;
    double eta;
    while (etafile.scanField("itype", it).scanField("jtype", jt).scanField("eta", eta).scanField()) {
      plumed_assert(it.empty() == false) << "itype is empty";
      plumed_assert(jt.empty() == false) << "jtype is empty";
      etas[nt * table[it] + table[jt]] = eta;
    }
  }
  if (energy_units != "plumed") {
    Units units;
    units.setEnergy(energy_units);
    for (auto& e : etas) {
      e *= units.getEnergy() / getUnits().getEnergy();
    }
  }
  setup(sw, "GHBFIX");
  // typesTable is indexed by absolute atom index (GHBFIX.cpp:188-196)
  const unsigned n = getNumberOfAtoms();
  std::vector<unsigned> mine(n);
  for (unsigned i = 0; i < n; ++i) {
    const auto a = getAbsoluteIndex(i).index();
    plumed_assert(a < typesTable.size()) << "your types table only covers " << typesTable.size()
                                         << " atoms, but you are trying to access atom number " << (a + 1);
    mine[i] = typesTable[a];
  }
  check(group ? b200coord_group_set_types(group, mine.data(), nt, etas.data()) : b200coord_set_types(ctx, mine.data(), nt, etas.data()),
        "set_types");
  log.printf("  %u interaction types from %s, scaling parameters from %s\n", nt, types.c_str(), params.c_str());
}

void CoordinationBaseB200::setup(const b200coord_switch& sw, const char* what) {
  parseFlag("SERIAL", serial);
  std::vector<AtomNumber> ga, gb;
  parseAtomList("GROUPA", ga);
  parseAtomList("GROUPB", gb);
  bool nopbc = !pbc;
  parseFlag("NOPBC", nopbc);
  pbc = !nopbc;
  bool dopair = false;
  parseFlag("PAIR", dopair);
  bool classic = false, cells = false;
  parseFlag("NLIST", classic);
  parseFlag("NLISTCELLS", cells);
  plumed_assert(!(cells && classic)) << "Please activate only one of the two version of the NL";
  plumed_assert(!(cells && dopair)) << "Pair is not compatible with the CELLS implementation of the NL";
  double nlCut = 0.0;
  int nlStride = 0;
  if (classic || cells) {
    parse("NL_CUTOFF", nlCut);
    if (nlCut <= 0.0) {
      error("NL_CUTOFF should be explicitly specified and positive");
    }
    parse("NL_STRIDE", nlStride);
    if (nlStride <= 0) {
      error("NL_STRIDE should be explicitly specified and positive");
    }
  }
  if (dopair && ga.size() != gb.size()) {
    error("when using PAIR option, the two groups should have the same number of elements");
  }

  // device: GPU_DEVICE keyword > B200COORD_DEVICE > (several MPI ranks: the launcher's local rank modulo the number
  // of visible devices, so that ranks of one node do not pile up on device 0) > the calling thread's current device
  int device = -1;
  bool deviceFromRank = false;
  if (const char* env = std::getenv("B200COORD_DEVICE")) {
    device = std::atoi(env);
  } else if (comm.Get_size() > 1) {
    const char* names[] = {"OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID", "LOCAL_RANK"};
    int local = -1;
    for (const char* nm : names) {
      if (const char* env = std::getenv(nm)) {
        local = std::atoi(env);
        break;
      }
    }
    if (local < 0) {
      local = comm.Get_rank();
    }
    int ndev = 0;
    if (b200coord_device_count(&ndev) == B200COORD_OK && ndev > 0) {
      device = local % ndev;
      deviceFromRank = true;
    }
  }
  parse("GPU_DEVICE", device);
  std::vector<int> devices;
  {
    std::string list;
    parse("GPU_DEVICES", list);
    std::size_t at = 0;
    while (at < list.size()) {
      std::size_t comma = list.find(',', at);
      if (comma == std::string::npos) {
        comma = list.size();
      }
      if (comma > at) {
        devices.push_back(std::atoi(list.substr(at, comma - at).c_str()));
      }
      at = comma + 1;
    }
  }
  parse("GPU_COUPLING", coupling);
  if (!coupling.empty()) {
    int cdev = 0;
    std::size_t cn = 0;
    if (b200coord_coupling_lookup(coupling.c_str(), &cdev, nullptr, nullptr, &cn) != B200COORD_OK) {
      error("GPU_COUPLING=" + coupling + ": the engine has not published device arrays under this name (b200coord_coupling_publish)");
    }
    if (devices.size() > 1) {
      error("GPU_COUPLING works on the engine's device; it cannot be combined with GPU_DEVICES");
    }
    device = cdev;  // the arrays decide where the action runs
    devices.clear();
    for (const AtomNumber& a : ga) {
      if (a.index() >= cn) {
        error("GPU_COUPLING=" + coupling + ": an atom of GROUPA is beyond the published arrays");
      }
    }
    for (const AtomNumber& a : gb) {
      if (a.index() >= cn) {
        error("GPU_COUPLING=" + coupling + ": an atom of GROUPB is beyond the published arrays");
      }
    }
  }
  bool fp32 = false;
  if (const char* env = std::getenv("B200COORD_FP32")) {
    fp32 = std::atoi(env) != 0;
  }
  {
    bool flag = false;
    parseFlag("GPU_FP32", flag);
    fp32 = fp32 || flag;
  }
  checkRead();

  addValueWithDerivatives();
  setNotPeriodic();

  std::vector<AtomNumber> all(ga);
  all.insert(all.end(), gb.begin(), gb.end());
  std::vector<unsigned> absIndex(all.size());
  for (unsigned i = 0; i < all.size(); ++i) {
    absIndex[i] = all[i].index();
  }

  b200coord_config cfg;
  cfg.abi_version = B200COORD_ABI_VERSION;
  cfg.device = device;
  cfg.precision = fp32 ? B200COORD_FP32 : B200COORD_FP64;
  cfg.style = gb.empty() ? B200COORD_STYLE_SINGLELIST : (dopair ? B200COORD_STYLE_PAIR : B200COORD_STYLE_TWOLIST);
  cfg.n_group_a = ga.size();
  cfg.n_group_b = gb.size();
  cfg.pbc = pbc ? 1 : 0;
  cfg.nl_mode = cells ? B200COORD_NL_CELLS : (classic ? B200COORD_NL_CLASSIC : B200COORD_NL_NONE);
  cfg.nl_cutoff = nlCut;
  cfg.nl_stride = nlStride;
  // the reference splits the pair range over MPI ranks and sums with Comm::Sum (CoordinationBase.cpp:152-170,
  // :218-224); here every rank's context takes the same share of the i-atoms and the partial results are
  // summed over the PLUMED communicator in exactly the same way.
  combineWithMpi = !serial && comm.Get_size() > 1;
  cfg.rank = combineWithMpi ? comm.Get_rank() : 0;
  cfg.nranks = combineWithMpi ? comm.Get_size() : 1;
  if (devices.size() > 1) {
    if (combineWithMpi) {
      error("GPU_DEVICES shards the atoms over the GPUs of one process; with several MPI ranks give each rank one device (GPU_DEVICE)");
    }
    if (dopair) {
      error("GPU_DEVICES is not available with PAIR");
    }
    const int rc = b200coord_group_create(&cfg, &sw, absIndex.data(), devices.data(), static_cast<int>(devices.size()), &group);
    if (rc != B200COORD_OK) {
      error(std::string("cannot set up the B200 ") + what + " engine on the given GPU_DEVICES: " + b200coord_last_error(nullptr));
    }
  } else {
    if (devices.size() == 1) {
      cfg.device = device = devices[0];
    }
    const int rc = b200coord_create(&cfg, &sw, absIndex.data(), &ctx);
    if (rc != B200COORD_OK) {
      error(std::string("cannot set up the B200 ") + what + " engine: " + b200coord_last_error(nullptr));
    }
  }
  requestAtoms(all);
  if (!coupling.empty()) {
    if (combineWithMpi || needCharges || checkNumericalDerivatives()) {
      error("GPU_COUPLING is not available with several MPI ranks, with charges from PLUMED or with NUMERICAL_DERIVATIVES");
    }
    bool identity = true;
    for (unsigned i = 0; i < absIndex.size(); ++i) {
      identity = identity && absIndex[i] == i;
    }
    if (b200coord_coupled_set_index(ctx, identity ? nullptr : absIndex.data()) != B200COORD_OK) {
      error(b200coord_last_error(ctx));
    }
    doNotRetrieve();  // PLUMED's host copy of the positions is not looked at
  } else {
    derivBuffer.resize(3 * all.size());
    setupFastHost();
  }
  if (const char* env = std::getenv("B200COORD_PLUGIN_TIMERS")) {
    timers = std::atoi(env) != 0;
  }

  char desc[512];
  b200coord_switch_describe(&sw, desc, sizeof(desc));
  log.printf("  B200-native %s (libb200coord, sm_100a kernels)\n", what);
  if (group) {
    log.printf("  i-atoms sharded over %d CUDA devices of this process (GPU_DEVICES), derivative rows and positions exchanged over NVLink peer memory\n",
               b200coord_group_size(group));
  } else if (device >= 0) {
    log.printf("  on CUDA device %d%s\n", device, deviceFromRank ? " (local MPI rank modulo visible devices)" : "");
  } else {
    log.printf("  on the current CUDA device\n");
  }
  if (!coupling.empty()) {
    log.printf("  positions read from and forces added to the device arrays published as '%s' (GPU_COUPLING); derivatives stay on the device\n",
               coupling.c_str());
  }
  if (combineWithMpi) {
    log.printf("  %d MPI ranks: each computes its share of the i-atoms on its own device, partial results are summed with Comm::Sum on the host\n",
               comm.Get_size());
  }
  if (fp32) {
    log.printf("  FP32 pair arithmetic (opt-in): results within 1e-5 of the FP64 path\n");
  }
  log.printf("  between two groups of %u and %u atoms\n", static_cast<unsigned>(ga.size()), static_cast<unsigned>(gb.size()));
  log.printf(pbc ? "  using periodic boundary conditions\n" : "  without periodic boundary conditions\n");
  if (dopair) {
    log.printf("  with PAIR option\n");
  }
  if (classic || cells) {
    log.printf("  using neighbor lists (%s) with\n", cells ? "cell superset" : "distance filtered");
    log.printf("  update every %d steps and cutoff %f\n", nlStride, nlCut);
  }
  log << "  contacts are counted with cutoff " << desc << "\n";
}

void CoordinationBaseB200::setupFastHost() {
  fastHost = false;
  bool always = false;  // B200COORD_HOST_FAST: 0 = never, 1 (default) = from 50000 atoms on, 2 = always (tests)
  if (const char* env = std::getenv("B200COORD_HOST_FAST")) {
    if (std::atoi(env) == 0) {
      return;
    }
    always = std::atoi(env) >= 2;
  }
  const unsigned n = getNumberOfAtoms();
  if (needCharges || checkNumericalDerivatives() || n == 0 || (n < 50000 && !always)) {
    return;  // small systems: PLUMED's own loops cost nothing
  }
  valueIdx.resize(n);
  std::vector<std::size_t> seen(n);
  for (unsigned i = 0; i < n; ++i) {
    valueIdx[i] = getValueIndices(getAbsoluteIndex(i));
    if (valueIdx[i].first != 0) {
      return;  // a virtual atom: positions live in another action's values
    }
    seen[i] = valueIdx[i].second;
  }
  std::sort(seen.begin(), seen.end());
  if (std::adjacent_find(seen.begin(), seen.end()) != seen.end()) {
    return;  // an atom listed twice: the parallel scatter would race
  }
  const char* names[3] = {"posx", "posy", "posz"};
  for (int k = 0; k < 3; ++k) {
    ActionToPutData* a = plumed.getActionSet().selectWithLabel<ActionToPutData*>(names[k]);
    if (!a || a->getNumberOfComponents() != 1) {
      return;
    }
    posValue[k] = a->copyOutput(0);
    if (!posValue[k] || posValue[k]->getRank() != 1) {
      return;
    }
  }
  posBuffer.resize(3 * static_cast<std::size_t>(n));
  doNotRetrieve();
  fastHost = true;
  log.printf("  host side: atoms gathered and forces scattered by the action itself (OpenMP) instead of PLUMED's serial loops\n");
}

void CoordinationBaseB200::clearDerivatives(const bool& force) {
  // PlumedMain clears the derivatives of every active action before calculate() (PlumedMain.cpp:1312-1318): one
  // thread filling 3N+10 doubles. calculate() below overwrites every one of them (value, 3N atom derivatives,
  // 9 box derivatives), so on the parallel host path the fill is skipped.
  if ((!fastHost && coupling.empty()) || force) {
    Colvar::clearDerivatives(force);
  }
}

void CoordinationBaseB200::apply() {
  if (!coupling.empty()) {
    Value* v = getPntrToValue();
    if (!v->forcesWereAdded()) {
      return;
    }
    const double ff = v->getForce(0);
    double* dForce = nullptr;
    if (b200coord_coupling_lookup(coupling.c_str(), nullptr, nullptr, &dForce, nullptr) != B200COORD_OK) {
      error(b200coord_last_error(nullptr));
    }
    check(b200coord_apply_coupled(ctx, ff, dForce), "apply_coupled");
    double f9[9];
    for (int j = 0; j < 9; ++j) {
      f9[j] = ff * lastVirial[j];
    }
    unsigned ind = 0;
    setForcesOnCell(f9, 9, ind);
    return;
  }
  if (!fastHost) {
    Colvar::apply();
    return;
  }
  Value* v = getPntrToValue();
  if (!v->forcesWereAdded()) {
    return;  // ActionWithValue::checkForForces: nobody pushed a force on the value
  }
  const double ff = v->getForce(0);
  const long n = static_cast<long>(getNumberOfAtoms());
  const double* d = derivBuffer.data();
  Value* vx = posValue[0];
  Value* vy = posValue[1];
  Value* vz = posValue[2];
  // Value::addForce sets a flag of the Value on every call, so calling it from many threads makes them fight over
  // one cache line (measured: 46 ms for 3 M calls on 16 threads). Instead: f * derivative goes into dense
  // per-component arrays in parallel, then one Value::addForces (a plain sum over the array) per component, the
  // three of them side by side.
  Value* pv[3] = {vx, vy, vz};
  for (int k = 0; k < 3; ++k) {
    const std::size_t nv = pv[k]->getNumberOfStoredValues();
    if (forceBuffer[k].size() != nv) {
      forceBuffer[k].assign(nv, 0.0);  // entries of atoms this action does not own stay zero
    }
  }
  double* fx = forceBuffer[0].data();
  double* fy = forceBuffer[1].data();
  double* fz = forceBuffer[2].data();
  const unsigned nt = OpenMP::getNumThreads();
  #pragma omp parallel for num_threads(nt) schedule(static)
  for (long i = 0; i < n; ++i) {  // distinct atoms -> distinct elements
    const std::size_t kk = valueIdx[i].second;
    fx[kk] = ff * d[3 * i];
    fy[kk] = ff * d[3 * i + 1];
    fz[kk] = ff * d[3 * i + 2];
  }
  #pragma omp parallel for num_threads(3) schedule(static, 1)
  for (int k = 0; k < 3; ++k) {
    pv[k]->addForces(View<const double>(forceBuffer[k].data(), forceBuffer[k].size()));
  }
  double f9[9];
  for (int j = 0; j < 9; ++j) {
    f9[j] = ff * lastVirial[j];
  }
  unsigned ind = 0;
  setForcesOnCell(f9, 9, ind);
}

CoordinationBaseB200::~CoordinationBaseB200() {
  if (timers && nCalls) {
    std::fprintf(stderr, "B200COORD plugin timers: %lu calls, engine %.3f ms/call, store-to-Value %.3f ms/call\n", nCalls,
                 1e3 * tEngine / nCalls, 1e3 * tStore / nCalls);
  }
  b200coord_group_destroy(group);
  b200coord_destroy(ctx);
}

void CoordinationBaseB200::prepare() {
  // NeighborList::prepare (src/tools/NeighborList.cpp:433-456); the full atom list stays requested on
  // every step (legal: the reduced list is only a communication optimisation of the CPU code)
  int willRebuild = 0;
  const int rc = group ? b200coord_group_prepare(group, static_cast<long>(getStep()), getExchangeStep() ? 1 : 0, &willRebuild)
                       : b200coord_prepare(ctx, static_cast<long>(getStep()), getExchangeStep() ? 1 : 0, &willRebuild);
  if (rc != B200COORD_OK) {
    error(group ? b200coord_group_last_error(group) : b200coord_last_error(ctx));
  }
}

void CoordinationBaseB200::calculate() {
  const unsigned n = getNumberOfAtoms();
  if (needCharges) {  // ActionAtomistic::getCharge, as DHEnergy::pairing reads them every step
    chargeBuffer.resize(n);
    bool changed = chargeBuffer.size() != chargeSent.size();
    for (unsigned i = 0; i < n; ++i) {
      chargeBuffer[i] = getCharge(i);
      changed = changed || chargeBuffer[i] != chargeSent[i];
    }
    if (changed) {
      check(group ? b200coord_group_set_charges(group, chargeBuffer.data()) : b200coord_set_charges(ctx, chargeBuffer.data()),
            "set_charges");
      chargeSent = chargeBuffer;
    }
  }
  double box[9];
  const Tensor& b = getBox();
  for (unsigned i = 0; i < 3; ++i)
    for (unsigned j = 0; j < 3; ++j) {
      box[3 * i + j] = b[i][j];
    }
  check(group ? b200coord_group_set_box(group, box) : b200coord_set_box(ctx, box), "set_box");
  double value = 0.0;
  double virial[9];
  if (!coupling.empty()) {
    const double* dPos = nullptr;
    if (b200coord_coupling_lookup(coupling.c_str(), nullptr, &dPos, nullptr, nullptr) != B200COORD_OK) {
      error(b200coord_last_error(nullptr));
    }
    const auto t0 = std::chrono::steady_clock::now();
    check(b200coord_calculate_coupled(ctx, dPos, &value, virial), "calculate_coupled");
    if (timers) {
      tEngine += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      nCalls++;
    }
    for (int j = 0; j < 9; ++j) {
      lastVirial[j] = virial[j];
    }
    setValue(value);
    setBoxDerivatives(Tensor(virial[0], virial[1], virial[2], virial[3], virial[4], virial[5], virial[6], virial[7], virial[8]));
    return;
  }
  const double* pos = n ? &getPositions()[0][0] : nullptr;
  if (fastHost) {  // the shared position values, read in parallel (retrieveAtoms was told not to)
    double* pb = posBuffer.data();
    const long nn = static_cast<long>(n);
    const unsigned ntg = OpenMP::getNumThreads();
    #pragma omp parallel for num_threads(ntg) schedule(static)
    for (long i = 0; i < nn; ++i) {
      const Vector p = getGlobalPosition(valueIdx[i]);
      pb[3 * i] = p[0];
      pb[3 * i + 1] = p[1];
      pb[3 * i + 2] = p[2];
    }
    pos = pb;
  }
  const auto t0 = std::chrono::steady_clock::now();
  check(group ? b200coord_group_calculate(group, pos, &value, derivBuffer.data(), virial)
              : b200coord_calculate(ctx, pos, &value, derivBuffer.data(), virial),
        "calculate");
  if (combineWithMpi) {
    comm.Sum(value);
    comm.Sum(derivBuffer.data(), derivBuffer.n);
    comm.Sum(&virial[0], 9);
  }
  const auto t1 = std::chrono::steady_clock::now();
  // Value::data was cleared by clearDerivatives() just before calculate() (PlumedMain.cpp:1312-1318), so
  // setting equals the reference's adding; distinct indices -> safe to spread over PLUMED's OpenMP threads
  Value* v = getPntrToValue();
  const double* d = derivBuffer.data();
  const long n3 = 3L * static_cast<long>(n);
  const unsigned nt = OpenMP::getNumThreads();
  #pragma omp parallel for num_threads(nt) schedule(static) if (n3 > 30000)
  for (long i = 0; i < n3; ++i) {
    v->setDerivative(static_cast<unsigned>(i), d[i]);
  }
  if (timers) {
    const auto t2 = std::chrono::steady_clock::now();
    tEngine += std::chrono::duration<double>(t1 - t0).count();
    tStore += std::chrono::duration<double>(t2 - t1).count();
    nCalls++;
  }
  for (int j = 0; j < 9; ++j) {
    lastVirial[j] = virial[j];
  }
  setValue(value);
  setBoxDerivatives(Tensor(virial[0], virial[1], virial[2], virial[3], virial[4], virial[5], virial[6], virial[7], virial[8]));
}

}  // namespace colvar
}  // namespace PLMD
