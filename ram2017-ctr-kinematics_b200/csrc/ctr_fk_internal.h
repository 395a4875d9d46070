// ctr_fk_internal.h — glue shared by the host-only and the CUDA translation units of libctrfk.
#pragma once
#include <stdint.h>

namespace ctrk {
// printf-style; stores a thread-local message returned by ctr_fk_last_error().
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
}  // namespace ctrk
