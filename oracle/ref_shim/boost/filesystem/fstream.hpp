// TEST INFRASTRUCTURE: stand-in for <boost/filesystem/fstream.hpp> (file streams keyed by the path stand-in)
#pragma once
#include <fstream>
#include <boost/filesystem.hpp>
namespace boost { namespace filesystem {
typedef std::ifstream ifstream;
typedef std::ofstream ofstream;
typedef std::fstream fstream;
} }
