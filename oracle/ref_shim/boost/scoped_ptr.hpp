// Shim: boost::scoped_ptr -> std::unique_ptr (only used for pimpl members of platform classes).
#pragma once
#include <memory>
namespace boost { template <class T> using scoped_ptr = std::unique_ptr<T>; }
