/* TEST INFRASTRUCTURE — default (no-op) body for the f/p/v observation hook that oracle/Makefile
 * splices into the reference's chain.c stream; oracle/dump_shim.c overrides it in minimap2-sw. */
#include <stdint.h>
#ifdef __cplusplus
extern "C"
#endif
__attribute__((weak)) void mm2_ref_hook_fpv(int64_t n, const int32_t *f, const int32_t *p, const int32_t *v)
{
	(void)n; (void)f; (void)p; (void)v;
}

/* libmm2ref.so links chain.c without options.c, which is where the reference defines the learned
 * HW/SW split constants (options.c:6).  All zero => software DP always (chain.c:101). */
float K1_HW, K2_HW, C_HW, K_SW, C_SW;
