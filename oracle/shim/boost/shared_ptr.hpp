// Minimal stand-in for a Boost header (Boost is not in this image): std:: equivalents, only what the
// reference's hider sources need to compile in place.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr; using std::weak_ptr; using std::enable_shared_from_this; using std::make_shared;
using std::static_pointer_cast; using std::dynamic_pointer_cast; using std::const_pointer_cast;
}
