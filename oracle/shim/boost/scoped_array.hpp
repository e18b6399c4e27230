// Minimal stand-in for <boost/scoped_array.hpp> (used by include/aqsis/util/autobuffer.h).
#pragma once
#include <memory>
#include <cstddef>
namespace boost {
template<class T> class scoped_array {
	std::unique_ptr<T[]> m_p;
public:
	explicit scoped_array(T* p = 0) : m_p(p) {}
	void reset(T* p = 0) { m_p.reset(p); }
	T& operator[](std::ptrdiff_t i) const { return m_p[i]; }
	T* get() const { return m_p.get(); }
	void swap(scoped_array& o) { m_p.swap(o.m_p); }
};
}
