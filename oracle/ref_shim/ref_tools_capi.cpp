// TEST INFRASTRUCTURE ONLY -- part of the oracle, never linked into the product.
//
// extern "C" entry points over the REAL reference tool classes (PLMD::SwitchingFunction, PLMD::Pbc,
// PLMD::LatticeReduction, PLMD::LinkCells, PLMD::NeighborList) as compiled into
// oracle/_ref/lib/libplumedKernel.so by oracle/Makefile.  Tests use these (through ctypes) to pin
// oracle/coord_oracle.c and to generate tests/golden/ fixtures (oracle/gen_golden.py).
// This file contains no algorithm of its own: each function forwards to one reference method.
#include "tools/AtomNumber.h"
#include "tools/Communicator.h"
#include "tools/LatticeReduction.h"
#include "tools/LinkCells.h"
#include "tools/NeighborList.h"
#include "tools/Pbc.h"
#include "tools/SwitchingFunction.h"
#include "tools/Tensor.h"
#include "tools/Tools.h"
#include "tools/Vector.h"
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace PLMD;

static Tensor to_tensor(const double b[9]) { return Tensor(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8]); }

extern "C" {

// ---- Tools::pbc (tools/Tools.h:545-571)
double ref_tools_pbc(double x) { return Tools::pbc(x); }

// ---- SwitchingFunction (tools/SwitchingFunction.cpp:1055-1184)
void* ref_switch_create(const char* definition, char* err, int errlen) {
  auto* sf = new SwitchingFunction();
  std::string errors;
  try {
    sf->set(definition, errors);
  } catch (std::exception& e) {
    errors = e.what();
  }
  if (err && errlen > 0) std::snprintf(err, errlen, "%s", errors.c_str());
  if (!errors.empty()) {
    delete sf;
    return nullptr;
  }
  return sf;
}
void* ref_switch_create_rational(int nn, int mm, double r0, double d0) {
  auto* sf = new SwitchingFunction();
  sf->set(nn, mm, r0, d0);
  return sf;
}
void ref_switch_free(void* h) { delete static_cast<SwitchingFunction*>(h); }
double ref_switch_calculate(void* h, double r, double* df) { return static_cast<SwitchingFunction*>(h)->calculate(r, *df); }
double ref_switch_calculate_sqr(void* h, double r2, double* df) {
  return static_cast<SwitchingFunction*>(h)->calculateSqr(r2, *df);
}
// out: d0 dmax dmax_2 invr0 invr0_2 stretch shift
void ref_switch_data(void* h, double out[7]) {
  const auto& d = static_cast<SwitchingFunction*>(h)->get_data();
  out[0] = d.d0;
  out[1] = d.dmax;
  out[2] = d.dmax_2;
  out[3] = d.invr0;
  out[4] = d.invr0_2;
  out[5] = d.stretch;
  out[6] = d.shift;
}

// ---- LatticeReduction::reduce (tools/LatticeReduction.cpp:140-192)
void ref_lattice_reduce(double t[9]) {
  Tensor tt = to_tensor(t);
  LatticeReduction::reduce(tt);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = tt[i][j];
}

// ---- Pbc (tools/Pbc.cpp:165-212, :362-415)
void* ref_pbc_create(const double box[9]) {
  auto* p = new Pbc();
  p->setBox(to_tensor(box));
  return p;
}
void ref_pbc_free(void* h) { delete static_cast<Pbc*>(h); }
int ref_pbc_is_orthorombic(void* h) { return static_cast<Pbc*>(h)->isOrthorombic() ? 1 : 0; }
void ref_pbc_distance(void* h, const double v1[3], const double v2[3], double d[3]) {
  Vector r = static_cast<Pbc*>(h)->distance(Vector(v1[0], v1[1], v1[2]), Vector(v2[0], v2[1], v2[2]));
  d[0] = r[0];
  d[1] = r[1];
  d[2] = r[2];
}
// n pairs at once: v1,v2,d are n*3
void ref_pbc_distance_many(void* h, const double* v1, const double* v2, double* d, size_t n) {
  auto* p = static_cast<Pbc*>(h);
  for (size_t i = 0; i < n; i++) {
    Vector r = p->distance(Vector(v1[3 * i], v1[3 * i + 1], v1[3 * i + 2]), Vector(v2[3 * i], v2[3 * i + 1], v2[3 * i + 2]));
    d[3 * i] = r[0];
    d[3 * i + 1] = r[1];
    d[3 * i + 2] = r[2];
  }
}
void ref_pbc_full_search(void* h, double d[3]) {
  Vector v(d[0], d[1], d[2]);
  static_cast<Pbc*>(h)->fullSearch(v);
  d[0] = v[0];
  d[1] = v[1];
  d[2] = v[2];
}

// ---- LinkCells (tools/LinkCells.cpp:85-122, :277-315, :195-239)
struct RefLinkCells {
  Communicator comm;
  LinkCells lc;
  Pbc pbc;
  RefLinkCells() : lc(comm) {}
};
void* ref_linkcells_create(double cutoff, const double* pos, size_t n, const double box[9]) {
  auto* h = new RefLinkCells();
  h->pbc.setBox(to_tensor(box));
  std::vector<Vector> p(n);
  for (size_t i = 0; i < n; i++) p[i] = Vector(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
  h->lc.setCutoff(cutoff);
  h->lc.setupCells(p, h->pbc);
  return h;
}
void ref_linkcells_free(void* h) { delete static_cast<RefLinkCells*>(h); }
void ref_linkcells_ncells(void* h, unsigned out[3]) {
  const auto& l = static_cast<RefLinkCells*>(h)->lc.getCellLimits();
  out[0] = l[0];
  out[1] = l[1];
  out[2] = l[2];
}
unsigned ref_linkcells_find_cell(void* h, const double pos[3]) {
  return static_cast<RefLinkCells*>(h)->lc.findCell(Vector(pos[0], pos[1], pos[2]));
}
unsigned ref_linkcells_required(void* h, const unsigned celn[3], int use_pbc, unsigned* out) {
  auto& lc = static_cast<RefLinkCells*>(h)->lc;
  std::vector<unsigned> req(27);
  unsigned n = 0;
  lc.addRequiredCells({celn[0], celn[1], celn[2]}, n, req, use_pbc != 0);
  for (unsigned i = 0; i < n; i++) out[i] = req[i];
  return n;
}

// ---- NeighborList (tools/NeighborList.cpp:43-101 ctors, :168-315 update)
// style: 0 Pair, 1 TwoList, 2 SingleList
struct RefNL {
  Communicator comm;
  Pbc pbc;
  std::unique_ptr<NeighborList> nl;
  size_t n = 0;
};
void* ref_nl_create(int style, unsigned n0, unsigned n1, int do_pbc, int use_cells, double cutoff, unsigned stride,
                    const double box[9]) {
  auto* h = new RefNL();
  h->pbc.setBox(to_tensor(box));
  std::vector<AtomNumber> l0(n0), l1(n1);
  for (unsigned i = 0; i < n0; i++) l0[i].setIndex(i);
  for (unsigned i = 0; i < n1; i++) l1[i].setIndex(n0 + i);
  try {
    if (style == 2) {
      if (stride > 0)
        h->nl = std::make_unique<NeighborList>(l0, true, do_pbc != 0, h->pbc, h->comm, cutoff, stride, use_cells != 0);
      else
        h->nl = std::make_unique<NeighborList>(l0, true, do_pbc != 0, h->pbc, h->comm);
      h->n = n0;
    } else {
      if (stride > 0)
        h->nl = std::make_unique<NeighborList>(l0, l1, true, style == 0, do_pbc != 0, h->pbc, h->comm, cutoff, stride,
                                               use_cells != 0);
      else
        h->nl = std::make_unique<NeighborList>(l0, l1, true, style == 0, do_pbc != 0, h->pbc, h->comm);
      h->n = (size_t)n0 + n1;
    }
  } catch (std::exception& e) {
    std::fprintf(stderr, "ref_nl_create: %s\n", e.what());
    delete h;
    return nullptr;
  }
  return h;
}
void ref_nl_free(void* h) { delete static_cast<RefNL*>(h); }
void ref_nl_update(void* hh, const double* pos) {
  auto* h = static_cast<RefNL*>(hh);
  std::vector<Vector> p(h->n);
  for (size_t i = 0; i < h->n; i++) p[i] = Vector(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
  h->nl->update(p);
}
size_t ref_nl_size(void* hh) { return static_cast<RefNL*>(hh)->nl->size(); }
void ref_nl_pairs(void* hh, unsigned* out) {
  auto* h = static_cast<RefNL*>(hh);
  size_t n = h->nl->size();
  for (size_t i = 0; i < n; i++) {
    auto p = h->nl->getClosePair(i);
    out[2 * i] = p.first;
    out[2 * i + 1] = p.second;
  }
}

// ---- plumed_cmd with C++ exceptions turned into return codes (core/PlumedMainInitializer.cpp:75-110, :499)
extern "C" void* plumed_plumedmain_create();
extern "C" void plumed_plumedmain_cmd(void* plumed, const char* key, const void* val);
extern "C" void plumed_plumedmain_finalize(void* plumed);

void* ref_plumed_create() {
  try {
    return plumed_plumedmain_create();
  } catch (...) {
    return nullptr;
  }
}
int ref_plumed_cmd(void* p, const char* key, const void* val, char* err, int errlen) {
  try {
    plumed_plumedmain_cmd(p, key, val);
    return 0;
  } catch (std::exception& e) {
    if (err && errlen > 0) std::snprintf(err, errlen, "%s", e.what());
    return 1;
  } catch (...) {
    if (err && errlen > 0) std::snprintf(err, errlen, "unknown exception");
    return 2;
  }
}
void ref_plumed_finalize(void* p) {
  try {
    plumed_plumedmain_finalize(p);
  } catch (...) {
  }
}

}  // extern "C"
