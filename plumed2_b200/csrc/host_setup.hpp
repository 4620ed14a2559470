// Host-side set-up shared by the C ABI (capi.cu) and usable without CUDA:
//   * switching-function definition parsing, automatic D_MAX, stretch/shift
//   * periodic box classification, inverse, lattice reduction, octant shift lists
//   * cell-grid dimensions
// These run once per action / per box change (SURVEY 8(a) rows a12, a15.5 "keep on host").
#pragma once
#include "../../include/b200coord.h"
#include <string>

namespace b200 {

constexpr int kMaxShift = 6;  // Pbc.h:58

struct HostPbc {
  int type = 0;  // 0 unset, 1 orthorhombic, 2 generic  (Pbc.h:52)
  double box[9] = {0}, inv_box[9] = {0}, reduced[9] = {0}, inv_reduced[9] = {0};
  int nshift[8] = {0};               // octant = 4*(s0>0)+2*(s1>0)+(s2>0)
  double shifts[8][kMaxShift][3] = {{{0}}};
};

// Pbc::setBox (src/tools/Pbc.cpp:165-212)
void setup_pbc(const double box[9], HostPbc& out);
// LatticeReduction::reduceFast (src/tools/LatticeReduction.cpp:144-192)
void reduce_lattice(double rows[9]);

// SwitchingFunction::set(string) (src/tools/SwitchingFunction.cpp:1055-1159)
int parse_switch(const std::string& definition, b200coord_switch& out, std::string& err);
// SwitchingFunction::set(nn,mm,r0,d0) (:1176-1184)
void rational_switch(int nn, int mm, double r0, double d0, b200coord_switch& out);
// GHBFIX ctor (src/colvar/GHBFIX.cpp:98-113)
void ghbfix_pairing(double dmax, double d0, double c, b200coord_switch& out);
// DHEnergy ctor (src/colvar/DHEnergy.cpp:104-128): k and constant/epsilon
void dhenergy_pairing(double I, double T, double epsilon, double energy_unit, double length_unit, double charge_unit,
                      b200coord_switch& out);
std::string describe_switch(const b200coord_switch& sw);
// s(r) on the host, used only for the two evaluations of setupStretch (:63-72)
double switch_value_host(const b200coord_switch& sw, double r);

// LinkCells::createCells (src/tools/LinkCells.cpp:99-122): cells per lattice direction for `cutoff`
void cell_grid(const double inv_box[9], double cutoff, unsigned ncells[3]);

}  // namespace b200
