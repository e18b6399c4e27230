// Minimal stand-in for <boost/utility/enable_if.hpp> (Boost is not in this image).
// Only what include/aqsis/math/math.h of the reference needs. Test infrastructure only.
#pragma once
#include <type_traits>
namespace boost {
template<class C, class T = void> struct enable_if : std::enable_if<C::value, T> {};
template<class C, class T = void> struct disable_if : std::enable_if<!C::value, T> {};
}
