// TEST INFRASTRUCTURE ONLY -- part of the oracle, never linked into the product.
//
// Stand-in for the reference's build-generated src/config/Config.cpp (that file
// includes Config.inc / version.h / Makefile.conf.xxd, all of which only exist
// after ./configure).  oracle/Makefile compiles the reference's hot-path sources
// directly with g++ (no configure, no reference Makefiles), so the handful of
// PLMD::config:: accessors declared in /root/reference/src/config/Config.h:25-115
// are provided here with fixed answers describing *this* mini build.
#include "config/Config.h"
#include <cstdlib>
#include <dlfcn.h>
#include <string>

namespace PLMD {
namespace config {

static std::string env_or(const char* key, const std::string& fallback) {
  const char* v = std::getenv(key);
  return v ? std::string(v) : fallback;
}

std::string getSoExt() { return "so"; }
bool isInstalled() { return false; }
// The CLI lists <root>/scripts and <root>/patches at start-up (core/CLToolMain.cpp:205-223).  The mini build
// ships no scripts; oracle/Makefile creates an empty <_ref>/root/{scripts,patches} next to lib/ so that the
// executable also starts on machines where the reference tree is not mounted (the GPU box).
std::string getPlumedRoot() {
  const char* v = std::getenv("PLUMED_ROOT");
  if (v) return std::string(v);
  std::string lib = getLibraryPath();              // .../_ref/lib/libplumedKernel.so
  const size_t a = lib.rfind('/');
  if (a == std::string::npos) return B200_REF_ROOT;
  const size_t b = lib.rfind('/', a - 1);
  if (b == std::string::npos) return B200_REF_ROOT;
  return lib.substr(0, b) + "/root";
}
std::string getPlumedHtmldir() { return getPlumedRoot(); }
std::string getPlumedIncludedir() { return getPlumedRoot() + "/src/include"; }
std::string getPlumedProgramName() { return "plumed"; }
std::string getEnvCommand() {
  return "env PLUMED_ROOT='" + getPlumedRoot() + "' PLUMED_VERSION='" + getVersionLong() +
         "' PLUMED_HTMLDIR='" + getPlumedHtmldir() + "' PLUMED_INCLUDEDIR='" + getPlumedIncludedir() +
         "' PLUMED_PROGRAM_NAME='plumed' PLUMED_IS_INSTALLED='no'";
}
std::string getMakefile() {
  return "# mini oracle build (oracle/Makefile): g++ -O3 -fopenmp, no MPI, internal BLAS/LAPACK\n";
}
std::string getVersion() { return "2.11"; }
std::string getVersionLong() { return "2.11.0-dev"; }
std::string getVersionGit() { return "oracle-mini-build"; }
std::string getCompilationDate() { return __DATE__; }
std::string getCompilationTime() { return __TIME__; }
bool hasMatheval() { return false; }
bool hasDlopen() {
#ifdef __PLUMED_HAS_DLOPEN
  return true;
#else
  return false;
#endif
}
bool hasMolfile() { return false; }
bool hasExternalMolfile() { return false; }
bool hasZlib() { return false; }

std::string getLibraryPath() {
  Dl_info info;
  if (dladdr((void*)&getLibraryPath, &info) && info.dli_fname) {
    return std::string(info.dli_fname);
  }
  return "";
}

}  // namespace config
}  // namespace PLMD
