/* Replacement for the reference's chain_hardware.h (/root/reference/chain_hardware.h:1-75) when minimap2 is built against
 * the B200 backend instead of the OpenCL/FPGA one.  The three host files that include it keep compiling unchanged:
 *
 *   main.c:367   if (!hardware_init(BUFFER_N, XCLBIN_FILE)) return -1;     -> mm2b_init_async() (CUDA devices, streams, workers)
 *   main.c:430   cleanup();                                                -> mm2b_shutdown_at_exit()
 *   options.c:95-99,118-122   K1_HW = ONT_K1_HW; ...                       -> the learned HW/SW split is gone: constants are 0
 *   chain.c      is NOT compiled; mm_chain_dp comes from libmm2chain_b200 (same signature, mmpriv.h:65)
 *
 * Like the original, this header is C++ (the reference compiles every .c with $(CXX), Makefile:184-185).
 */
#ifndef MM2B_COMPAT_CHAIN_HARDWARE_H
#define MM2B_COMPAT_CHAIN_HARDWARE_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "minimap.h"
#define MM2B_HOST_DECLARES_MM_CHAIN_DP     /* mmpriv.h:65 declares it with mm128_t; same C symbol */
#include "mm2chain_b200.h"

/* learned HW/SW split (chain_hardware.h:18-30): removed — every read is chained on the GPU */
#define ONT_K1_HW 0.0f
#define ONT_K2_HW 0.0f
#define ONT_C_HW 0.0f
#define ONT_K_SW 0.0f
#define ONT_C_SW 0.0f
#define PBCCS_K1_HW 0.0f
#define PBCCS_K2_HW 0.0f
#define PBCCS_C_HW 0.0f
#define PBCCS_K_SW 0.0f
#define PBCCS_C_SW 0.0f

#define XCLBIN_FILE ((char*)"")    /* no bitstream: kernels are compiled into the library */
#define BUFFER_N (500L << 20)      /* not a device buffer any more (workspaces grow with the batches seen): the bytes of read sequence per
                                      mini-batch (-K, 500M by default, options.c:53) whose pinned staging is reserved while the index loads */

static inline bool hardware_init(long buf_size, char *binary_name)
{
	(void)binary_name;
	if (mm2b_init_async(0, 0) != MM2B_OK) {      /* devices come up while main() loads the index (main.c:371) */
		fprintf(stderr, "[ERROR] B200 chaining backend: %s\n", mm2b_last_error());
		return false;
	}
	if (buf_size > 0 && !(getenv("MM2B_RESERVE") && atoi(getenv("MM2B_RESERVE")) == 0)) mm2b_reserve_for_mapping((size_t)buf_size);
	return true;
}
static inline void cleanup(void) { mm2b_shutdown_at_exit(); }      /* main.c:430: the process ends right after */

#endif
