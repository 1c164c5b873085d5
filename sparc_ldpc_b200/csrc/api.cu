// api.cu -- library-wide state of libsparc_b200 (error string, launch counter, version).
#include "common.cuh"

namespace sb {
thread_local char g_err[512] = "";
std::atomic<long> g_launches{0};
}  // namespace sb

extern "C" const char *sb_last_error(void) { return sb::g_err; }
extern "C" int sb_version(void) { return 100; }
extern "C" long sb_launch_count(void) { return sb::g_launches.load(); }
extern "C" void sb_launch_count_reset(void) { sb::g_launches.store(0); }
