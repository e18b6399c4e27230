// Minimal stand-in for a Boost header (Boost is not in this image): std:: equivalents, only what the
// reference's hider sources need to compile in place.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
namespace boost {
// Calls the unqualified intrusive_ptr_add_ref / intrusive_ptr_release found by ADL, like Boost's.
template<class T> class intrusive_ptr {
	T* m_p;
public:
	typedef T element_type;
	intrusive_ptr() : m_p(0) {}
	intrusive_ptr(T* p, bool add_ref = true) : m_p(p) { if(m_p && add_ref) intrusive_ptr_add_ref(m_p); }
	intrusive_ptr(const intrusive_ptr& r) : m_p(r.m_p) { if(m_p) intrusive_ptr_add_ref(m_p); }
	template<class U> intrusive_ptr(const intrusive_ptr<U>& r) : m_p(r.get()) { if(m_p) intrusive_ptr_add_ref(m_p); }
	~intrusive_ptr() { if(m_p) intrusive_ptr_release(m_p); }
	intrusive_ptr& operator=(const intrusive_ptr& r) { intrusive_ptr(r).swap(*this); return *this; }
	intrusive_ptr& operator=(T* r) { intrusive_ptr(r).swap(*this); return *this; }
	void reset() { intrusive_ptr().swap(*this); }
	void reset(T* r) { intrusive_ptr(r).swap(*this); }
	T* get() const { return m_p; }
	T& operator*() const { return *m_p; }
	T* operator->() const { return m_p; }
	explicit operator bool() const { return m_p != 0; }
	bool operator!() const { return m_p == 0; }
	void swap(intrusive_ptr& r) { T* t = m_p; m_p = r.m_p; r.m_p = t; }
};
template<class T, class U> bool operator==(const intrusive_ptr<T>& a, const intrusive_ptr<U>& b) { return a.get() == b.get(); }
template<class T, class U> bool operator!=(const intrusive_ptr<T>& a, const intrusive_ptr<U>& b) { return a.get() != b.get(); }
template<class T> bool operator==(const intrusive_ptr<T>& a, T* b) { return a.get() == b; }
template<class T> bool operator!=(const intrusive_ptr<T>& a, T* b) { return a.get() != b; }
template<class T> bool operator<(const intrusive_ptr<T>& a, const intrusive_ptr<T>& b) { return a.get() < b.get(); }
template<class T> T* get_pointer(const intrusive_ptr<T>& p) { return p.get(); }
}
