#!/bin/bash
# The reference's own A/B timing harness (src/cltools/Benchmark.cpp:399-871): `plumed benchmark --plumed ref.dat:gpu.dat`
# runs both inputs on the same generated trajectory and prints comparative timings.  ref.dat = the CPU action,
# gpu.dat = the same line behind LOAD of the plugin.  20000 atoms so that the reference can use NLIST (it segfaults above 32768).
# usage: scripts/plumed_benchmark_ab.sh [natoms] [nsteps]   (needs oracle/_ref and a GPU)
set -e
cd "$(dirname "$0")/.."
N=${1:-20000}
STEPS=${2:-200}
T=$(mktemp -d)
L=$(python -c "print((${N}/100.0)**(1/3.))")
BODY="GROUPA=1-${N} SWITCH={RATIONAL R_0=0.3 NN=6 MM=12 D_MAX=0.8} NLIST NL_CUTOFF=1.0 NL_STRIDE=10"
printf 'c: COORDINATION %s\nRESTRAINT ARG=c AT=0 KAPPA=0 SLOPE=1\n' "$BODY" > $T/ref.dat
printf 'LOAD FILE=%s/plumed2_b200/lib/libb200coord_plumed.so\nc: COORDINATION %s\nRESTRAINT ARG=c AT=0 KAPPA=0 SLOPE=1\n' "$PWD" "$BODY" > $T/gpu.dat
export PLUMED_NUM_THREADS=$(nproc) PLUMED_IGNORE_NL_MEMORY_ERROR=1
# cube|scale: N uniform points in a cube of edge cbrt(N), scaled to 100 atoms/nm^3; the cube distribution
# redraws all positions every frame, which keeps both arms on their rebuild-independent worst case
oracle/_ref/bin/plumed benchmark --plumed "$T/ref.dat:$T/gpu.dat" --natoms $N --nsteps $STEPS --atom-distribution "cube|scale 0.2154" 2>&1 | tail -40
rm -rf $T
