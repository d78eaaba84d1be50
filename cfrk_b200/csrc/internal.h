// internal.h -- shared between the translation units of libcfrk_b200.so.
#pragma once
#include <string>

namespace cfrk {
// text returned by cfrk_last_error() for the calling thread
void set_last_error(const std::string& msg);
}
