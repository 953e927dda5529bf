// Shim for the one Boost header include/mitsuba/mitsuba.h:24 pulls in (Boost is absent from this image).
#pragma once
#define BOOST_VERSION 105400
