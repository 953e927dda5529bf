// ref_integrator_sppm.cpp — TEST INFRASTRUCTURE.  SPPMIntegrator::scaleVolumeAPA (photonmapper/sppm.cpp:255-290) of the
// REFERENCE'S OWN primal integrator class, called on raw storage; sppm.cpp is compiled as part of this translation unit
// from where it lies under /root/reference (see ref_integrator.cpp for the arrangement).
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <future>
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

#define private public
#define protected public
#include "sppm.cpp"
#undef private
#undef protected

using namespace mitsuba;

extern "C" {

// scales[k] = globalScaleVolume after scaleVolumeAPA(it), it = 1 .. n.  technique: EVolumeTechnique value.
void ref_int_sppm_scale_volume_apa(float scale0, int n, float alpha, int technique, const char *force_apa, float *scales) {
  SPPMIntegrator *I = reinterpret_cast<SPPMIntegrator *>(std::calloc(1, sizeof(SPPMIntegrator) + 64));
  new (&I->m_forceAPA) std::string(force_apa ? force_apa : "");
  I->m_alpha = alpha;
  I->m_volTechnique = (EVolumeTechnique)technique;
  I->m_independentScale = false;               // sppm.cpp:238
  I->globalScaleVolume = scale0;
  for (int it = 1; it <= n; ++it) {
    I->scaleVolumeAPA(it);
    scales[it - 1] = (float)I->globalScaleVolume;
  }
  I->m_forceAPA.~basic_string();
  std::free(I);
}

}  // extern "C"
