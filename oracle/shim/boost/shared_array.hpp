// Minimal stand-in for a Boost header (Boost is not in this image): std:: equivalents, only what the
// reference's hider sources need to compile in place.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <memory>
#include <cstddef>
namespace boost {
template<class T> class shared_array {
	std::shared_ptr<T> m_p;
public:
	shared_array() {}
	explicit shared_array(T* p) : m_p(p, std::default_delete<T[]>()) {}
	void reset(T* p = 0) { if(p) m_p.reset(p, std::default_delete<T[]>()); else m_p.reset(); }
	T& operator[](std::ptrdiff_t i) const { return m_p.get()[i]; }
	T* get() const { return m_p.get(); }
	explicit operator bool() const { return bool(m_p); }
	bool operator!() const { return !m_p; }
};
}
