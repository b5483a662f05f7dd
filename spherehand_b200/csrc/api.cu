// Library-wide C-ABI helpers: version, last error text.
#include "common.cuh"

char g_sh_last_error[512] = "";
unsigned long long g_sh_launches = 0;



SH_EXPORT const char* sh_last_error(void) { return g_sh_last_error; }
SH_EXPORT int sh_abi_version(void) { return 3; }
SH_EXPORT const char* sh_build_arch(void) { return "sm_100a"; }
SH_EXPORT long sh_launch_count(void) { return (long)g_sh_launches; }
