// internal.h -- shared between the translation units of libcfrk_b200.so.
#pragma once
#include <cstddef>
#include <string>

namespace cfrk {
// text returned by cfrk_last_error() for the calling thread
void set_last_error(const std::string& msg);

// pinned-memory arena behind rd->Freq / cfrk_free_host (api.cu)
void* pinned_alloc(size_t bytes);
void pinned_free(void* p);
void pinned_release_cached();

// per-stream launch scratch of the dense kernels (kernels.cu)
void release_stream_scratch();
}
