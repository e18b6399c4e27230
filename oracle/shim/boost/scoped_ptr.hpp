// Minimal stand-in for a Boost header (Boost is not in this image): std:: equivalents, only what the
// reference's hider sources need to compile in place.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <memory>
namespace boost {
template<class T> class scoped_ptr {
	std::unique_ptr<T> m_p;
public:
	explicit scoped_ptr(T* p = 0) : m_p(p) {}
	void reset(T* p = 0) { m_p.reset(p); }
	T& operator*() const { return *m_p; }
	T* operator->() const { return m_p.get(); }
	T* get() const { return m_p.get(); }
	explicit operator bool() const { return bool(m_p); }
	bool operator!() const { return !m_p; }
	void swap(scoped_ptr& o) { m_p.swap(o.m_p); }
};
}
