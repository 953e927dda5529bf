// Shim: boost::function -> std::function.
#pragma once
#include <functional>
namespace boost { template <class S> using function = std::function<S>; }
