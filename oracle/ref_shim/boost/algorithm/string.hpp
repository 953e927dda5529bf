// TEST INFRASTRUCTURE: stand-in for <boost/algorithm/string.hpp> (the reference uses to_lower_copy / to_lower only)
#pragma once
#include <algorithm>
#include <cctype>
#include <string>
namespace boost {
inline void to_lower(std::string &s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); }); }
inline std::string to_lower_copy(std::string s) { to_lower(s); return s; }
inline void to_upper(std::string &s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::toupper(c); }); }
inline std::string to_upper_copy(std::string s) { to_upper(s); return s; }
}
