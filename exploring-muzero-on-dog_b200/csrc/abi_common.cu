// abi_common.cu — version / error-string entry points of libdogstep.so.
#include <cstring>
#include "common.cuh"

namespace dogstep {
static thread_local char g_last_error[256] = "";
void set_last_error(const char* msg) {
  std::strncpy(g_last_error, msg ? msg : "", sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}
}  // namespace dogstep

extern "C" {
int dogstep_version(void) { return 100; }
const char* dogstep_last_error(void) { return dogstep::g_last_error; }
}
