// platform_stubs.cpp — TEST INFRASTRUCTURE.  Link-time stand-ins for the handful of NON-ARITHMETIC
// libcore symbols (logging, threads, streams, serialization registry) that the reference's hot-path
// headers reference but that cannot be compiled here because their implementation files need Boost.Thread /
// Boost.Filesystem / Xerces.  Everything that computes a number on the pinned path is the
// reference's own source, compiled from /root/reference by oracle/Makefile:
//   include/mitsuba/core/{kdtree,aabb,triangle,ray,point,vector,frame,transform,math}.h   (header code)
//   src/libcore/{util,transform,object,class,timer,quad,random}.cpp
//   src/integrators/photonmapper/gvpm/gvpm_accel.{h,cpp}, beams_accel.{h,cpp}, plane_accel.h,
//   beams_struct.h, beams_3d_intersections.h, plane_struct.h
// The stubs below carry no arithmetic: a logger that is never installed (Thread::getLogger() returns NULL,
// so the Log()/SLog() macros of include/mitsuba/core/logger.h:33-56 skip the call), unreachable stream
// accessors, the serialization base-class constructor, and the data-only PhotonSubBeam constructor whose
// home file (photonmapper/beams.cpp:336-342) drags in the particle-tracing process classes.
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/serialization.h>
#include <mitsuba/core/stream.h>
#include <mitsuba/core/thread.h>
#include <mitsuba/core/logger.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "beams_accel.h"

MTS_NAMESPACE_BEGIN

static char g_fakeThread[64];
Thread *Thread::getThread() { return reinterpret_cast<Thread *>(g_fakeThread); }
Logger *Thread::getLogger() { return NULL; }

void Logger::log(ELogLevel level, const Class *, const char *file, int line, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::fprintf(stderr, "[ref %d] %s:%d: ", (int)level, file, line);
  std::vfprintf(stderr, fmt, ap);
  std::fprintf(stderr, "\n");
  va_end(ap);
  if (level >= EError) std::abort();
}

static void unreachable(const char *what) {
  std::fprintf(stderr, "gvpm ref harness: unexpected call to %s\n", what);
  std::abort();
}
int Stream::readInt() { unreachable("Stream::readInt"); return 0; }
unsigned int Stream::readUInt() { unreachable("Stream::readUInt"); return 0; }
void Stream::writeInt(int) { unreachable("Stream::writeInt"); }
void Stream::writeUInt(unsigned int) { unreachable("Stream::writeUInt"); }
void Stream::readULongArray(uint64_t *, size_t) { unreachable("Stream::readULongArray"); }
void Stream::writeULongArray(const uint64_t *, size_t) { unreachable("Stream::writeULongArray"); }

SerializableObject::SerializableObject(Stream *, InstanceManager *) { unreachable("SerializableObject(Stream*)"); }
MTS_IMPLEMENT_CLASS(SerializableObject, true, Object)

// photonmapper/beams.cpp:336-342 (field assignments only)
PhotonSubBeam::PhotonSubBeam(const Point &pos, const PhotonBeam *beam, Float t1, Float t2) {
  position = pos;
  data.beam = beam;
  data.t1 = t1;
  data.t2 = t2;
  flags = 0;
}

MTS_NAMESPACE_END
