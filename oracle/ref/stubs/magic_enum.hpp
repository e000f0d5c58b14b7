// Stub of magic_enum (only enum_name is used, inside log messages).
#pragma once
#include <string>
namespace magic_enum { template <class E> inline std::string enum_name(E e) { return std::to_string((long long)e); } }
