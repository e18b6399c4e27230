// Minimal stand-in for a Boost header (Boost is not in this image): std:: equivalents, only what the
// reference's hider sources need to compile in place.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <boost/noncopyable.hpp>
#include <boost/utility/enable_if.hpp>
#include <iterator>
namespace boost {
template<class T> T next(T x) { return ++x; }
template<class T> T prior(T x) { return --x; }
}
