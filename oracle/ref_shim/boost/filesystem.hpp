// Shim: just enough of boost::filesystem::path for the declarations in mitsuba/core/plugin.h,
// fresolver.h etc. to parse.  Nothing on the harness' path touches the file system.
#pragma once
#include <string>
namespace boost { namespace filesystem {
class path {
public:
  path() {}
  path(const std::string &s) : m(s) {}
  path(const char *s) : m(s) {}
  std::string string() const { return m; }
  const char *c_str() const { return m.c_str(); }
  bool empty() const { return m.empty(); }
  path filename() const { return *this; }
  path extension() const { return *this; }
  path parent_path() const { return *this; }
  path operator/(const path &o) const { return path(m + "/" + o.m); }
  bool operator==(const path &o) const { return m == o.m; }
private:
  std::string m;
};
// never reached by the harness (declarations that mipmap.h's disk cache needs to parse)
inline unsigned long long file_size(const path &) { return 0; }
inline bool exists(const path &) { return false; }
inline bool remove(const path &) { return false; }
inline long last_write_time(const path &) { return 0; }
} }
