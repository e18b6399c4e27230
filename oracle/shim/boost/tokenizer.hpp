// Minimal stand-in for a Boost header (Boost is not in this image): std:: equivalents, only what the
// reference's hider sources need to compile in place.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#pragma once
#include <string>
namespace boost {
// Declaration-level stand-in: aqsis/util/file.h only names the type in a typedef that the hider never uses.
template<class F, class It = std::string::const_iterator, class T = std::string> class tokenizer;
}
