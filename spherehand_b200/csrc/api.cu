// Library-wide C-ABI helpers: version, last error text.
#include "common.cuh"

char g_sh_last_error[512] = "";



SH_EXPORT const char* sh_last_error(void) { return g_sh_last_error; }
SH_EXPORT int sh_abi_version(void) { return 1; }
SH_EXPORT const char* sh_build_arch(void) { return "sm_100a"; }
