// poisson_harness.cpp — TEST INFRASTRUCTURE.  Flat C entry into the reference's own screened-Poisson solver
// (src/integrators/poisson_solver/{Solver,Backend,BackendOpenMP,Defs}.cpp, compiled from where they lie by
// `make -C oracle poisson_ref`, nothing copied): the call sequence of gvpm.cpp:631-636 / 667-676.
// This marshals arrays and parameters only; all arithmetic is the reference's.
#include <cstring>

#include "Solver.hpp"

extern "C" int gvpm_ref_poisson_solve(int w, int h, const float *throughput, const float *dx, const float *dy,
                                      const float *direct, float alpha, int irlsIterMax, float irlsRegInit,
                                      float irlsRegIter, int cgIterMax, int cgIterCheck, int cgPrecond,
                                      float cgTolerance, const char *backend, float *reconstruction) {
  poisson::Solver::Params p;
  p.setConfigPreset("L2D");
  p.alpha = alpha;
  p.irlsIterMax = irlsIterMax;
  p.irlsRegInit = irlsRegInit;
  p.irlsRegIter = irlsRegIter;
  p.cgIterMax = cgIterMax;
  p.cgIterCheck = cgIterCheck;
  p.cgPrecond = cgPrecond != 0;
  p.cgTolerance = cgTolerance;
  p.backend = backend;   // "Naive" (single thread, Backend.cpp) or "OpenMP"
  p.setLogFunction(poisson::Solver::Params::LogFunction([](const std::string &) {}));
  poisson::Solver s(p);
  s.importImagesMTS(const_cast<float *>(dx), const_cast<float *>(dy), const_cast<float *>(throughput),
                    const_cast<float *>(direct), w, h);
  s.setupBackend();
  s.solveIndirect();
  s.exportImagesMTS(reconstruction);
  return 0;
}

// Solver::Params::setConfigPreset (Solver.cpp:90-158) as data: {irlsIterMax, irlsRegInit, irlsRegIter, cgIterMax,
// cgIterCheck, cgPrecond, cgTolerance}
extern "C" int gvpm_ref_poisson_preset(const char *preset, int *irlsIterMax, float *irlsRegInit, float *irlsRegIter,
                                       int *cgIterMax, int *cgIterCheck, int *cgPrecond, float *cgTolerance,
                                       float *alpha) {
  poisson::Solver::Params p;
  if (!p.setConfigPreset(preset)) return -1;
  *irlsIterMax = p.irlsIterMax; *irlsRegInit = p.irlsRegInit; *irlsRegIter = p.irlsRegIter;
  *cgIterMax = p.cgIterMax; *cgIterCheck = p.cgIterCheck; *cgPrecond = p.cgPrecond ? 1 : 0;
  *cgTolerance = p.cgTolerance; *alpha = p.alpha;
  return 0;
}
