/* mm2chain_b200 — C ABI of the B200-native chaining backend for minimap2 (drop-in for the FPGA offload of
 * kisarur/minimap2-fpga).  Plain C: pointers and sizes only, no CUDA or torch types in any signature.
 *
 * What each entry point replaces in the reference (/root/reference):
 *
 *   mm_chain_dp            chain.c:29 / mmpriv.h:65     same name, same 16 positional arguments, same ownership rules.
 *                                                       Called per read from mm_map_frag (map.c:316, map.c:338).  Every
 *                                                       read is chained on the GPU; there is no HW/SW predictor and no
 *                                                       CPU fallback (chain.c:51-101 is gone).
 *   mm2b_init              chain_hardware.h:70 / chain_hardware.cpp:278  hardware_init(long, char*): OpenCL platform,
 *                                                       xclbin load, device buffers  ->  CUDA devices, streams, pinned rings.
 *   mm2b_shutdown          chain_hardware.h:71 / chain_hardware.cpp:403  cleanup()
 *   mm2b_chain_batch       chain_hardware.h:68 / chain_hardware.cpp:27   run_chaining_on_hw(): one blocking offload call
 *                                                       with host buffers — but for MANY reads at once and with the full
 *                                                       software semantics (max_skip, max_iter, gap_scale, is_cdna, n_segs),
 *                                                       returning final chains (u[], b[]) instead of raw f[]/p[].
 *   mm2b_chain_batch_device  (no equivalent)            the same computation on buffers already resident in HBM, on a
 *                                                       caller-supplied stream: the building block the two above use.
 *   mm2b_last_error        chain_hardware.h:72 / chain_hardware.cpp:208  checkError(): the reference prints and exit()s;
 *                                                       mm_chain_dp keeps that behaviour, the mm2b_* calls return codes.
 *
 * C++ hosts that still call hardware_init()/cleanup() by those names (main.c:367, main.c:430) include
 * minimap2-fpga_b200/host/compat/chain_hardware.h, which maps them onto mm2b_init/mm2b_shutdown.
 */
#ifndef MM2CHAIN_B200_H
#define MM2CHAIN_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM2B_ABI_VERSION 4

/* == mm128_t (minimap.h:53).  x = rev<<63 | rid<<32 | ref_pos;  y = seg_id<<48 | flags(40..43) | q_span<<32 | q_pos */
typedef struct { uint64_t x, y; } mm2b_anchor_t;

/* The chaining arguments of mm_chain_dp (chain.c:29) in call order. */
typedef struct {
	int32_t max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs;
	float gap_scale;
} mm2b_params_t;

/* Per-read outcome, mirroring the three ways mm_chain_dp returns:
 *   EMPTY     n == 0                       -> returns NULL, *_u = NULL, *n_u_ = 0      (chain.c:38-41)
 *   NO_CHAIN  no chain end reaches min_sc  -> returns NULL, *_u = NULL, *n_u_ = 0      (chain.c:355-358)
 *   OK        otherwise                    -> *_u non-NULL (even when n_u == 0), b = kmalloc(n_v*16) */
enum { MM2B_READ_EMPTY = 0, MM2B_READ_NO_CHAIN = 1, MM2B_READ_OK = 2 };

/* Return codes */
enum { MM2B_OK = 0, MM2B_ERR_CUDA = -1, MM2B_ERR_ARG = -2, MM2B_ERR_CAPACITY = -3, MM2B_ERR_NOT_INIT = -4 };

typedef struct {
	int64_t n_reads, n_anchors;
	int64_t n_chains, n_chained;     /* totals: sum n_u, sum n_v */
	int64_t cells_issued;            /* GPU lanes evaluated (32 per chunk); 0 unless counting is on */
	int64_t cells_ref;               /* iterations of the reference's inner j loop (chain.c:197) these reads take, `continue`d ones
	                                    included, up to the max_skip break: the "cell" of the GCUPS metric; 0 unless counting is on */
	int64_t window_cells;            /* sum over anchors of the window size i - st after the max_iter clamp (chain.c:192-193); 0 unless counting is on */
	int64_t n_general_reads;         /* reads that took the general (multi-segment / cDNA / gap_scale != 1) scoring path */
	double  h2d_ms, kernel_ms, d2h_ms; /* device-side timings of the last host-buffer call (CUDA events), 0 for device calls */
	int64_t n_heavy_reads;           /* reads chained by the heavy-read kernel (one CTA per read: long windows, e.g. tandem repeats) */
	/* host-buffer calls only (ABI v3): what crossed PCIe and what the host threads did around it */
	int64_t h2d_bytes, d2h_bytes;    /* bytes copied host->device / device->host by this call */
	int64_t n_packed_subs, n_raw_subs; /* sub-batches whose anchors went over as 8-byte words + runs / as 16-byte mm128_t */
	double  pack_ms, gather_ms;      /* summed wall time of the host-side packing / b[] gathering tasks (over all helper threads) */
	int64_t n_cut_reads;             /* long reads cut at x-gap cut points (chain.c:192) into pieces filled by warps of their own */
} mm2b_stats_t;

/* ---- lifecycle ------------------------------------------------------------------------------------------------ */

/* Bind the backend to `n_devices` CUDA devices (ids in `devices`, or NULL for 0..n_devices-1; n_devices <= 0 means
 * "all visible", or the MM2B_DEVICES environment variable if set).  Creates per-device streams and worker threads. */
int mm2b_init(int n_devices, const int *devices);
/* Same, but returns at once and brings the devices up on a background thread (only "no GPU at all" is reported immediately);
 * the first call that needs the backend waits for it.  Lets a host overlap CUDA start-up with its own (e.g. index loading). */
int mm2b_init_async(int n_devices, const int *devices);
void mm2b_shutdown(void);
/* For a host that calls its clean-up right before the process ends (main.c:430): stops and joins the library's threads but leaves device and
 * pinned memory to the operating system (un-pinning gigabytes of staging costs as much as mapping a mini-batch).  The library must not be
 * used afterwards. */
void mm2b_shutdown_at_exit(void);
int mm2b_num_devices(void);                 /* devices bound by mm2b_init (0 before) */
void mm2b_set_counting(int on);             /* the same statistics switch for mm2b_chain_batch / mm_chain_dp (all internal workspaces) */
int mm2b_cuda_device_count(void);           /* devices visible to the CUDA runtime; <= 0 when there is no usable GPU */
const char *mm2b_last_error(void);          /* thread-local text of the last failure */
int mm2b_abi_version(void);

/* Pinned host memory for batch inputs/outputs (H2D/D2H run at PCIe speed only from pinned pages). */
void *mm2b_host_alloc(size_t bytes);
void mm2b_host_free(void *p);               /* the block goes back to the library's pool (released by mm2b_shutdown) */
/* Fill that pool ahead of time with n_blocks blocks of `bytes` (pinning memory is slow: ~0.5 ms per MB); mm2b_host_pool_trim empties it. */
void mm2b_host_reserve(size_t bytes, int n_blocks);
void mm2b_host_pool_trim(void);
/* After the devices are up (mm2b_init_async), reserve on the same background thread the pinned staging that mapping mini-batches of
 * about `seq_bytes` of read sequence will ask for (mm2seed_b200.h), so that the first mini-batch does not pay for it. */
void mm2b_reserve_for_mapping(size_t seq_bytes);

/* ---- batch chaining, host buffers (the end-to-end path) ------------------------------------------------------- */

/* Chain `n_reads` independent reads.  Read r owns anchors a[off[r] .. off[r+1]) (sorted by x, as map.c:245 leaves them).
 * Reads are sharded over the bound devices by per-device worker threads; per-read results come back in input order:
 *   n_u[r], n_v[r], status[r]                 per read
 *   u[u_off[r] .. u_off[r]+n_u[r])            == the reference's final u[] for read r   (chain.c:419)
 *   b[b_off[r] .. b_off[r]+n_v[r])            == the reference's final b[] for read r   (chain.c:420)
 * u_off[r] / b_off[r] say where read r's results are; the packing order inside u / b is NOT the read order (every read's warp
 * reserves its own share of the output).  u_off/b_off have n_reads+1 entries (the last one is off[n_reads], the capacity
 * actually needed).  u_cap/b_cap are the capacities of u/b in elements; off[n_reads] always suffices.
 * `stats` may be NULL.  Blocking.  Thread-safe; concurrent calls share the devices. */
int mm2b_chain_batch(const mm2b_params_t *par, int64_t n_reads, const int64_t *off, const mm2b_anchor_t *a,
                     int32_t *n_u, int32_t *n_v, int32_t *status, int64_t *u_off, int64_t *b_off,
                     uint64_t *u, int64_t u_cap, mm2b_anchor_t *b, int64_t b_cap, mm2b_stats_t *stats);

/* The same call with the two PCIe diets spelled out.
 *   bi      when non-NULL receives, at the same offsets as b, the INDEX of every chained anchor inside its read
 *           (b[b_off[r]+k] == a[off[r] + bi[b_off[r]+k]]): 4 bytes instead of 16 come back over PCIe, and a caller that still
 *           holds a[] (every caller of mm_chain_dp does: map.c:316 passes it in) gathers b itself or uses the indices directly.
 *           `b` may then be NULL.  With both b and bi the library gathers b on its helper threads.
 *   flags   MM2B_F_RAW_INPUT      send anchors as 16-byte mm128_t.  Default: helper threads pack sub-batches into 8-byte
 *                                 {x_lo, y_lo} words plus run-length lists of the high words (strand/rid; flags/q_span/segment) on
 *                                 their way into the pinned staging buffer and the device restores mm128_t in HBM.  Packing is a
 *                                 pass over host memory; MM2B_PACK_INFLIGHT (default 99 = every sub-batch) bounds how many sub-batches are packed
 *                                 at a time, the others — and any sub-batch whose high words change too often (e.g. a
 *                                 homopolymer-compressed index) — go over raw by themselves.
 *           MM2B_F_DEVICE_GATHER  b[] comes back from the device as 16-byte anchors (the default when only b is asked for).
 *           MM2B_F_HOST_GATHER    b[] is gathered on the host's helper threads from 4-byte indices (pays off only where host
 *                                 memory bandwidth is plentiful compared with the PCIe link).
 * mm2b_chain_batch itself (b[] out as anchors) sends the input raw and gathers on the device: with 16 B per chained anchor coming back the
 * host's memory system is busy enough, and packing next to it measured slower.  Environment overrides (tuning): MM2B_PACK=1 (pack there
 * too), MM2B_PACK_INFLIGHT=n, MM2B_GATHER=host|device, MM2B_HOST_THREADS=n, MM2B_ONE_STREAM=1 (copies and kernels of a pipeline slot on
 * one stream instead of three: slower, kept for comparison). */
enum { MM2B_F_RAW_INPUT = 1, MM2B_F_DEVICE_GATHER = 2, MM2B_F_HOST_GATHER = 4 };
int mm2b_chain_batch_ex(const mm2b_params_t *par, int64_t n_reads, const int64_t *off, const mm2b_anchor_t *a,
                        int32_t *n_u, int32_t *n_v, int32_t *status, int64_t *u_off, int64_t *b_off,
                        uint64_t *u, int64_t u_cap, mm2b_anchor_t *b, int32_t *bi, int64_t b_cap, unsigned flags, mm2b_stats_t *stats);

/* ---- batch chaining, device buffers (inputs already in HBM) --------------------------------------------------- */

typedef struct mm2b_workspace mm2b_workspace_t;

/* Scratch for batches of up to max_anchors anchors / max_reads reads on `device` (32 B per anchor + 9 B per read). */
mm2b_workspace_t *mm2b_ws_create(int device, int64_t max_anchors, int64_t max_reads);
void mm2b_ws_destroy(mm2b_workspace_t *ws);
size_t mm2b_ws_bytes(const mm2b_workspace_t *ws);
/* Statistics switch: when on, batches run on this workspace also tally cells_ref / cells_issued (mm2b_stats_t).  Off by default
 * (the tally costs kernel time and is not part of the result); also switched on by the environment variable MM2B_COUNT_CELLS=1. */
void mm2b_ws_set_counting(mm2b_workspace_t *ws, int on);
/* Optional hint for the NEXT batch on this workspace: the anchor count of its longest read (the host knows the offsets, the
 * device call does not).  When no read can be long enough for the heavy-read kernel, that kernel and its classification pass
 * are not launched at all.  -1 (the default, restored after every batch) = unknown: classify on the device. */
void mm2b_ws_set_longest_read(mm2b_workspace_t *ws, int64_t n_anchors);

/* All d_* pointers are device memory on the workspace's device; `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  Asynchronous: work is enqueued on `stream` and the call returns.  `n_anchors` == off[n_reads] (known to the host).
 * d_u_off/d_b_off: n_reads+1 entries (entry n_reads = total entries written); d_u/d_b: capacity n_anchors elements each is
 * always enough.  Offsets are per read; the packing order is not the read order (see mm2b_chain_batch). */
int mm2b_chain_batch_device(mm2b_workspace_t *ws, const mm2b_params_t *par, int64_t n_reads, int64_t n_anchors,
                            const int64_t *d_off, const mm2b_anchor_t *d_a,
                            int32_t *d_n_u, int32_t *d_n_v, int32_t *d_status, int64_t *d_u_off, int64_t *d_b_off,
                            uint64_t *d_u, mm2b_anchor_t *d_b, void *stream);

/* The same with the chained anchors returned as int32 indices inside their read (d_bi, capacity n_anchors) instead of copies. */
int mm2b_chain_batch_device_idx(mm2b_workspace_t *ws, const mm2b_params_t *par, int64_t n_reads, int64_t n_anchors,
                                const int64_t *d_off, const mm2b_anchor_t *d_a,
                                int32_t *d_n_u, int32_t *d_n_v, int32_t *d_status, int64_t *d_u_off, int64_t *d_b_off,
                                uint64_t *d_u, int32_t *d_bi, void *stream);
/* Restore 16-byte anchors in HBM from the packed transfer format: d_lo[n_anchors] = {x_lo, y_lo} (uint32 pairs) and two run
 * lists of {first anchor of the run, high word} (uint32 pairs, sorted, run 0 starts at anchor 0) for x and y. */
int mm2b_unpack_anchors_device(int device, int64_t n_anchors, const void *d_lo, const void *d_xruns, int32_t n_xruns,
                               const void *d_yruns, int32_t n_yruns, mm2b_anchor_t *d_a, void *stream);

/* Host side of that format (pure CPU, what the helper threads of mm2b_chain_batch run per chunk): lo receives n uint32 pairs,
 * xruns / yruns up to cap_runs uint32 pairs each.  MM2B_ERR_CAPACITY when the high words change more often than cap_runs. */
int mm2b_pack_anchors(const mm2b_anchor_t *a, int64_t n, void *lo, void *xruns, int32_t *n_xruns, void *yruns, int32_t *n_yruns,
                      int32_t cap_runs);

/* Host memory bandwidth as n_threads concurrent memcpy's see it (GB/s, bytes read + bytes written): the ceiling of any host-side
 * pass over the anchors (packing, gathering), reported by bench.py next to the PCIe copy rates. */
double mm2b_measure_host_copy(int n_threads, size_t bytes_per_thread);

/* Counters of the last batch run on this workspace (synchronises the given stream). */
int mm2b_ws_stats(mm2b_workspace_t *ws, void *stream, mm2b_stats_t *stats);
/* Device time of the dominant kernel (the warp-per-read chaining kernel) in the last batch run on this workspace, in ms,
 * from CUDA events recorded on the launching stream around that one launch (synchronises on the second event). */
double mm2b_ws_chain_kernel_ms(mm2b_workspace_t *ws);
/* Kernels launched by this library since load (for bench.py's gpu_launches). */
int64_t mm2b_launch_count(void);
/* Debug / test access to the per-anchor DP state of the last batch: f[], p[], v[] as the reference has them at chain.c:238.
 * Only valid when the workspace was created with the environment variable MM2B_KEEP_FPV=1 (costs 12 B/anchor extra). */
int mm2b_ws_copy_fpv(mm2b_workspace_t *ws, void *stream, int64_t n_anchors, int32_t *h_f, int32_t *h_p, int32_t *h_v);

/* Range-check violations recorded by the debug build of the library (libmm2chain_b200_dbg.so, -DMM2B_DEBUG_CHECKS): bit 31 set means
 * "this is the checking build", the low bits are violation codes.  Always 0 in the release build. */
unsigned mm2b_debug_flags(void);

/* Measured INT32 issue peak of `device` in G lane-instructions/s (LOP3 + VIADDMNMX chains on all SMs; 2 SASS instructions per
 * source statement of 3 integer operations, so the operation-counted peak is 1.5 x this), for the roofline. */
double mm2b_measure_int32_peak(int device);

/* ---- the reference's own per-read boundary -------------------------------------------------------------------- */

/* Drop-in for chain.c:29.  Takes ownership of `a` (kfree(km, a) on every path); returns b and *_u allocated with
 * kmalloc(km, ...).  Re-entrant; called concurrently from the kt_for worker threads (map.c:561).  Initialises the backend
 * on first use if mm2b_init was not called.  Fatal CUDA errors print to stderr and exit(1), as checkError does. */
#ifndef MM2B_HOST_DECLARES_MM_CHAIN_DP   /* hosts that include mmpriv.h already have the prototype (with mm128_t) */
mm2b_anchor_t *mm_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                           float gap_scale, int is_cdna, int n_segs, int64_t n, mm2b_anchor_t *a, int *n_u_, uint64_t **_u,
                           void *km, int tid);
#endif

#ifdef __cplusplus
}
#endif
#endif
