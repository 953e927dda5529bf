// physics_stubs.cpp — TEST INFRASTRUCTURE.  Link-time stand-ins for the NON-ARITHMETIC plumbing that the reference's
// medium / phase / BSDF / emitter plugins and shift_diffuse.cpp pull in: the Properties container (the reference's
// src/libcore/properties.cpp needs boost::variant), the ConfigurableObject / NetworkedObject bookkeeping that lives in the
// same file, a constant texture (src/librender/texture.cpp needs boost::filesystem through mipmap.h), and abort-stubs
// for the plugin manager, the hardware renderer and the stream accessors.  Nothing here evaluates a radiometric
// quantity: every number the pin tests compare comes out of the reference's own sources compiled by oracle/Makefile
// (target physics_ref):
//   src/medium/homogeneous.cpp, src/phase/{isotropic,hg}.cpp, src/bsdfs/diffuse.cpp, src/emitters/area.cpp,
//   src/librender/{medium,phase,bsdf,emitter,shader,shape}.cpp, src/libcore/{warp,util,...}.cpp, src/libbidir/vertex.cpp,
//   src/integrators/photonmapper/gvpm/shift/operation/shift_diffuse.cpp
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/properties.h>
#include <mitsuba/core/cobject.h>
#include <mitsuba/core/netobject.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/statistics.h>
#include <mitsuba/core/sched.h>
#include <mitsuba/render/texture.h>
#include <mitsuba/render/shape.h>
#include <mitsuba/hw/renderer.h>
#include <mitsuba/hw/gpuprogram.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/track.h>
#include <mitsuba/render/sensor.h>
#include <cstdio>
#include <cstdlib>
#include <map>

MTS_NAMESPACE_BEGIN

void phys_unreachable(const char *what) {
  std::fprintf(stderr, "gvpm physics ref harness: unexpected call to %s\n", what);
  std::abort();
}

// ---- Properties: a typed key/value container ------------------------------------------------------------------------
struct PropertyElement {
  int kind = 0;  // 0 bool, 1 int, 2 float, 3 string, 4 spectrum
  bool b = false;
  int64_t i = 0;
  Float f = 0;
  std::string s;
  Spectrum spec;
};
Properties::Properties() : m_elements(new std::map<std::string, PropertyElement>()) {}
Properties::Properties(const std::string &pluginName)
    : m_elements(new std::map<std::string, PropertyElement>()), m_pluginName(pluginName) {}
Properties::Properties(const Properties &p)
    : m_elements(new std::map<std::string, PropertyElement>(*p.m_elements)), m_pluginName(p.m_pluginName), m_id(p.m_id) {}
Properties::~Properties() { delete m_elements; }
void Properties::operator=(const Properties &p) { *m_elements = *p.m_elements; m_pluginName = p.m_pluginName; m_id = p.m_id; }
bool Properties::hasProperty(const std::string &name) const { return m_elements->count(name) != 0; }
void Properties::setBoolean(const std::string &n, const bool &v, bool) { PropertyElement e; e.kind = 0; e.b = v; (*m_elements)[n] = e; }
void Properties::setInteger(const std::string &n, const int &v, bool) { PropertyElement e; e.kind = 1; e.i = v; (*m_elements)[n] = e; }
void Properties::setFloat(const std::string &n, const Float &v, bool) { PropertyElement e; e.kind = 2; e.f = v; (*m_elements)[n] = e; }
void Properties::setString(const std::string &n, const std::string &v, bool) { PropertyElement e; e.kind = 3; e.s = v; (*m_elements)[n] = e; }
void Properties::setSpectrum(const std::string &n, const Spectrum &v, bool) { PropertyElement e; e.kind = 4; e.spec = v; (*m_elements)[n] = e; }
#define PHYS_GET(T, fn, field)                                                                   \
  T Properties::fn(const std::string &n) const {                                                 \
    auto it = m_elements->find(n);                                                               \
    if (it == m_elements->end()) phys_unreachable(("Properties: missing " + n).c_str());         \
    return it->second.field;                                                                     \
  }                                                                                              \
  T Properties::fn(const std::string &n, const T &d) const {                                     \
    auto it = m_elements->find(n);                                                               \
    return it == m_elements->end() ? d : (T)it->second.field;                                    \
  }
PHYS_GET(bool, getBoolean, b)
PHYS_GET(int, getInteger, i)
PHYS_GET(Float, getFloat, f)
PHYS_GET(std::string, getString, s)
PHYS_GET(Spectrum, getSpectrum, spec)
std::string Properties::toString() const { return "Properties[harness]"; }

// ---- ConfigurableObject / NetworkedObject bookkeeping (same file as Properties in the reference) ----------------------
void ConfigurableObject::setParent(ConfigurableObject *) {}
void ConfigurableObject::addChild(const std::string &, ConfigurableObject *) { phys_unreachable("ConfigurableObject::addChild"); }
void ConfigurableObject::configure() {}
void ConfigurableObject::serialize(Stream *, InstanceManager *) const { phys_unreachable("ConfigurableObject::serialize"); }
ConfigurableObject::ConfigurableObject(Stream *s, InstanceManager *m) : SerializableObject(s, m) {}
MTS_IMPLEMENT_CLASS(ConfigurableObject, true, SerializableObject)
void NetworkedObject::bindUsedResources(ParallelProcess *) const {}
void NetworkedObject::wakeup(ConfigurableObject *, std::map<std::string, SerializableObject *> &) {}
void NetworkedObject::serialize(Stream *, InstanceManager *) const { phys_unreachable("NetworkedObject::serialize"); }
MTS_IMPLEMENT_CLASS(NetworkedObject, true, ConfigurableObject)

// ---- plugin manager / instance manager / renderer: never reached ------------------------------------------------------
ref<PluginManager> PluginManager::m_instance;
ConfigurableObject *PluginManager::createObject(const Class *, const Properties &) { phys_unreachable("PluginManager::createObject"); return NULL; }
ConfigurableObject *PluginManager::createObject(const Properties &) { phys_unreachable("PluginManager::createObject"); return NULL; }
SerializableObject *InstanceManager::getInstance(Stream *) { phys_unreachable("InstanceManager::getInstance"); return NULL; }
void InstanceManager::serialize(Stream *, const SerializableObject *) { phys_unreachable("InstanceManager::serialize"); }
Shader *Renderer::registerShaderForResource(const HWResource *) { phys_unreachable("Renderer::registerShaderForResource"); return NULL; }
void Renderer::unregisterShaderForResource(const HWResource *) { phys_unreachable("Renderer::unregisterShaderForResource"); }
Float Stream::readSingle() { phys_unreachable("Stream::readSingle"); return 0; }
void Stream::writeSingle(float) { phys_unreachable("Stream::writeSingle"); }
void Stream::readSingleArray(float *, size_t) { phys_unreachable("Stream::readSingleArray"); }
void Stream::writeSingleArray(const float *, size_t) { phys_unreachable("Stream::writeSingleArray"); }

void Stream::writeUChar(unsigned char) { phys_unreachable("Stream::writeUChar"); }
unsigned char Stream::readUChar() { phys_unreachable("Stream::readUChar"); return 0; }
void Stream::writeString(const std::string &) { phys_unreachable("Stream::writeString"); }
Properties::EPropertyType Properties::getType(const std::string &) const { phys_unreachable("Properties::getType"); return EBoolean; }
// AbstractEmitter's constructor asks for its "toWorld" transform (src/librender/emitter.cpp:28); AnimatedTransform lives in
// src/libcore/track.cpp, which needs Eigen.  The harness' area light is never transformed (evalDirection / pdfDirection
// only read the position record's normal): no transform object.
ref<const AnimatedTransform> Properties::getAnimatedTransform(const std::string &, const Transform &) const { return NULL; }
// statistics counters of the plugins (src/libcore/statistics.cpp registers them with a global singleton): inert here
StatsCounter::StatsCounter(const std::string &, const std::string &, EStatsType, uint64_t, uint64_t) : m_value(NULL), m_base(NULL) {}
StatsCounter::~StatsCounter() {}
std::string Spectrum::toString() const { return "Spectrum[harness]"; }
Spectrum Spectrum::CIE_D65(1.0f);     // set by Spectrum::staticInitialization in src/libcore/spectrum.cpp (needs boost::filesystem); the harness always passes a radiance
Class *Sensor::m_theClass = NULL;          // only referenced by a derivesFrom() check in Shape::addChild, never reached

MTS_NAMESPACE_END

// Symbols that are referenced from code the harness never reaches (bitmap export of a constant texture, stream
// constructors of animated transforms): satisfied by name only, so that no Bitmap / AnimatedTransform class code (and
// the image libraries behind it) has to be compiled.
extern "C" {
void _ZN7mitsuba17AnimatedTransformC1EPNS_6StreamE() { mitsuba::phys_unreachable("AnimatedTransform(Stream*)"); }
void _ZNK7mitsuba17AnimatedTransform9serializeEPNS_6StreamE() { mitsuba::phys_unreachable("AnimatedTransform::serialize"); }
void _ZN7mitsuba6BitmapC1ENS0_12EPixelFormatENS0_16EComponentFormatERKNS_8TVector2IiEEhPh() { mitsuba::phys_unreachable("Bitmap"); }
void _ZN7mitsuba6Bitmap19arithmeticOperationENS0_20EArithmeticOperationEPKS0_S3_() { mitsuba::phys_unreachable("Bitmap::arithmeticOperation"); }
}
