// Shim: BOOST_STATIC_ASSERT -> static_assert (Boost is absent from this image).
#pragma once
#define BOOST_STATIC_ASSERT(x) static_assert(x, #x)
