// Minimal stand-in for a Boost header (Boost is not in this image): std:: equivalents, only what the
// reference's hider sources need to compile in place.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <functional>
namespace boost { using std::bind; using std::ref; using std::cref; }
using namespace std::placeholders;
