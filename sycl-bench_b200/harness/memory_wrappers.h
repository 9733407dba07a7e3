// Same include name as /root/reference/include/memory_wrappers.h so that benchmark sources keep their
// #include lines; the SYCL-free implementation lives in sb_harness.h.
#pragma once
#include "sb_harness.h"
