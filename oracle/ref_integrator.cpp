// ref_integrator.cpp — TEST INFRASTRUCTURE.  Calls two member functions of the REFERENCE'S OWN integrator class
//   GPMIntegrator::scaleVolumeAPA     gvpm/gvpm.cpp:181-215    per-iteration kernel reduction (row a18)
//   GPMIntegrator::computeGradient    gvpm/gvpm.cpp:1205-1304  gradient images from the per-pixel accumulators (row a19)
// The class has no header: gvpm.cpp is compiled as part of this translation unit from where it lies under /root/reference
// (nothing is copied; oracle/Makefile, target integrator_ref -> _ref/libgvpm_integrator_ref.so).  The integrator object is
// raw zeroed storage with the handful of members these two functions read poked in (its constructor needs the whole
// renderer); GatherPoints and Bitmaps likewise.  Everything else of the class is dropped by the linker (--gc-sections).
#include <algorithm>
#include <array>
#include <atomic>
#include <bitset>
#include <cassert>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <stack>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <typeinfo>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>
#include <condition_variable>
#include <future>

// the members below are private
#define private public
#define protected public
#include "gvpm/gvpm.cpp"
#undef private
#undef protected

using namespace mitsuba;

namespace {
template <class T> T *rawZeroed() { return reinterpret_cast<T *>(std::calloc(1, sizeof(T) + 64)); }
}  // namespace

extern "C" {

int ref_int_version() { return 1; }

// scales[k] = globalScaleVolume after scaleVolumeAPA(it) for it = 1 .. n, starting from scale0 (= initialScaleVolume,
// gvpm.cpp:291).  technique: EVolumeTechnique value; force_apa: "" | "1D" | "2D" | "3D".
void ref_int_scale_volume_apa(float scale0, int n, float alpha, int technique, const char *force_apa,
                              int use_3d_kernel_reduction, float *scales) {
  GPMIntegrator *I = rawZeroed<GPMIntegrator>();
  new (&I->m_config) GPMConfig();
  I->m_config.alpha = alpha;
  I->m_config.volTechnique = (EVolumeTechnique)technique;
  I->m_config.forceAPA = force_apa ? force_apa : "";
  I->m_config.use3DKernelReduction = use_3d_kernel_reduction != 0;
  I->m_config.volumePhotonCount = 0;
  I->m_independentScale = false;               // gvpm.cpp:107
  I->globalScaleVolume = scale0;
  for (int it = 1; it <= n; ++it) {
    I->scaleVolumeAPA(it);
    scales[it - 1] = (float)I->globalScaleVolume;
  }
  I->m_config.~GPMConfig();
  std::free(I);
}

// acc: [w * h * 27] per pixel (row-major, y * w + x) = mediumFlux, shiftedMediumFlux[4], weightedMediumFlux[4];
// gx, gy: [w * h * 3].  Volume-only rendering (m_totalEmittedSurface = 0, directTracing off).
void ref_int_compute_gradient(const float *acc, int w, int h, int use_abs, int technique, size_t total_emitted_volume,
                              float *gx, float *gy) {
  GPMIntegrator *I = rawZeroed<GPMIntegrator>();
  new (&I->m_config) GPMConfig();
  I->m_config.volTechnique = (EVolumeTechnique)technique;
  I->m_config.directTracing = false;
  I->m_totalEmittedSurface = 0;
  I->m_totalEmittedVolume = total_emitted_volume;
  std::vector<GatherPoint> gps((size_t)w * h);
  new (&I->m_imgGP) std::vector<std::vector<GatherPoint *>>();
  I->m_imgGP.resize(w);
  for (int x = 0; x < w; ++x) {
    I->m_imgGP[x].resize(h);
    for (int y = 0; y < h; ++y) {
      GatherPoint &g = gps[(size_t)y * w + x];
      const float *a = acc + 27 * ((size_t)y * w + x);
      for (int c = 0; c < 3; ++c) {
        g.mediumFlux[c] = a[c];
        for (int k = 0; k < 4; ++k) {
          g.shiftedMediumFlux[k][c] = a[3 * (1 + k) + c];
          g.weightedMediumFlux[k][c] = a[3 * (5 + k) + c];
        }
      }
      I->m_imgGP[x][y] = &g;
    }
  }
  // Bitmaps: raw storage with the size and data pointer the function reads
  Bitmap *bx = rawZeroed<Bitmap>(), *by = rawZeroed<Bitmap>();
  bx->m_size = by->m_size = Vector2i(w, h);
  bx->m_data = reinterpret_cast<uint8_t *>(gx);
  by->m_data = reinterpret_cast<uint8_t *>(gy);
  I->computeGradient(1, Vector2i(w, h), bx, by, use_abs != 0);
  I->m_imgGP.~vector();
  I->m_config.~GPMConfig();
  std::free(bx);
  std::free(by);
  std::free(I);
}

}  // extern "C"
