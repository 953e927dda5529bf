// TEST INFRASTRUCTURE: stand-in for <boost/filesystem/fstream.hpp> (file streams keyed by the path stand-in)
#pragma once
#include <fstream>
#include <boost/filesystem.hpp>
namespace boost { namespace filesystem {
struct ifstream : std::ifstream {
  ifstream() {}
  explicit ifstream(const path &p, std::ios_base::openmode m = std::ios_base::in) : std::ifstream(p.string(), m) {}
  explicit ifstream(const std::string &p, std::ios_base::openmode m = std::ios_base::in) : std::ifstream(p, m) {}
};
struct ofstream : std::ofstream {
  ofstream() {}
  explicit ofstream(const path &p, std::ios_base::openmode m = std::ios_base::out) : std::ofstream(p.string(), m) {}
  explicit ofstream(const std::string &p, std::ios_base::openmode m = std::ios_base::out) : std::ofstream(p, m) {}
};
typedef std::fstream fstream;
} }
