// diffuclass.cpp:63-64 creates an output directory through boost::filesystem; the parity pin writes no files.
#pragma once
#include <string>
namespace boost { namespace filesystem {
struct path { std::string s; path() {} path(const std::string &x) : s(x) {} path(const char *x) : s(x) {} };
inline bool create_directory(const path &) { return true; }
} }
