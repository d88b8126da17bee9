// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <string>

#include "ops.h"

namespace dm {
std::string& last_error_ref();
template <class F>
int abi_guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& ex) {
    last_error_ref() = ex.what();
    return -1;
  } catch (...) {
    last_error_ref() = "unknown error";
    return -1;
  }
}
int device_sm_count();
}  // namespace dm
