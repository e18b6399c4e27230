// Minimal stand-in for <boost/type_traits/arithmetic_traits.hpp>. Test infrastructure only.
#pragma once
#include <type_traits>
namespace boost {
template<class T> struct is_arithmetic : std::is_arithmetic<T> {};
template<class T> struct is_integral : std::is_integral<T> {};
template<class T> struct is_float : std::is_floating_point<T> {};
}
