// Minimal stand-in for a Boost header (Boost is not in this image): std:: equivalents, only what the
// reference's hider sources need to compile in place.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
namespace boost {
class noncopyable {
protected:
	noncopyable() {}
	~noncopyable() {}
private:
	noncopyable(const noncopyable&);
	const noncopyable& operator=(const noncopyable&);
};
}
