// core.cu -- error plumbing, ABI version and launch accounting for libvpf_b200.so
#include "common.cuh"
#include <string.h>

namespace vpf {

std::atomic<int64_t> g_launch_count{0};

char *last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace vpf

extern "C" {

int vpf_abi_version(void) { return VPF_ABI_VERSION; }
const char *vpf_last_error_string(void) { return vpf::last_error_buf(); }
int64_t vpf_launch_count(void) { return vpf::g_launch_count.load(); }
void vpf_launch_count_reset(void) { vpf::g_launch_count.store(0); }

}
