/*
 * qsvc_b200.h -- C ABI of the B200-native MCTF hot path (libqsvc_b200.so).
 *
 * The reference (claudio382/QSVC) has no in-process plugin API for this path:
 * its interface is one process per tool, argv flags, and headerless files in
 * the current directory.  Each entry point below is the
 * in-process equivalent of one reference tool `main()`; the flag-compatible
 * command-line tools in qsvc_b200/tools/ and bin/mctf are thin wrappers that
 * read the files, call one of these functions, and write the files.
 *
 * Conventions
 *   - plain C, no CUDA or torch types; all pointers are HOST pointers to
 *     caller-owned contiguous buffers unless the name ends in `_resident`;
 *   - frames are I420 u8: Y (Y*X), U, V ((Y/2)*(X/2)); `even` holds
 *     n_pairs+1 frames, `odd`/`high` hold n_pairs frames;
 *   - motion fields: per pair 4 planes PREV.X, PREV.Y, NEXT.X, NEXT.Y of
 *     (Y/block_size)*(X/block_size) int16 (reference motion.cpp:9-15,93-101);
 *   - frame types: one byte per pair, 'I' or 'B';
 *   - return 0 on success, a negative QSVC_E* code otherwise; the message is
 *     available from qsvc_last_error() (per thread);
 *   - one context per GPU; a context is not thread-safe; contexts are
 *     independent of each other.  There is no CPU fallback: without a CUDA
 *     device every call fails with QSVC_ECUDA.
 */
#ifndef QSVC_B200_H
#define QSVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSVC_OK 0
#define QSVC_EINVAL (-1)  /* bad argument / unsupported geometry            */
#define QSVC_ECUDA (-2)   /* CUDA runtime error (no device, launch failure) */
#define QSVC_ENOMEM (-3)  /* device or host allocation failed               */
#define QSVC_EDOMAIN (-4) /* input outside the reference's defined domain   */

typedef struct qsvc_ctx qsvc_ctx;

/* Library / device management. */
int qsvc_version(void);
int qsvc_device_count(void);
qsvc_ctx *qsvc_create(int device);
void qsvc_destroy(qsvc_ctx *ctx);
const char *qsvc_last_error(void);
/* Number of kernels this context has launched so far (bench.py gpu_launches). */
long long qsvc_launch_count(const qsvc_ctx *ctx);
/* Device-side timing on the context's own stream (CUDA events). */
int qsvc_timer_start(qsvc_ctx *ctx);
int qsvc_timer_stop(qsvc_ctx *ctx, float *elapsed_ms);
int qsvc_synchronize(qsvc_ctx *ctx);
/* Per-kernel-class device timers (CUDA event pairs around every launch while
 * enabled).  Classes: 0 image load/store/border, 1 DWT rows, 2 DWT columns,
 * 3 block search, 4 predict, 5 residue/reconstruct, 6 update, 7 sub-pixel search
 * exact path.  read() waits for
 * the stream, sums what was recorded since the previous read and resets. */
#define QSVC_KERNEL_CLASSES 8
int qsvc_profile_enable(qsvc_ctx *ctx, int on);
int qsvc_profile_read(qsvc_ctx *ctx, float *ms_per_class, long long *launches_per_class,
                      int n_classes);
/* Measured SAD-instruction issue rate of this GPU (the ME roofline denominator):
 * SAD operations per second with packed bytes (__vsadu4) and 32-bit lanes (__sad). */
int qsvc_int_peak(qsvc_ctx *ctx, double *u8_sad_ops_per_s, double *i32_sad_ops_per_s);
/* Motion-estimation implementation: 0 automatic (fused sub-pixel path when the
 * geometry allows it, default; also env QSVC_ME_MODE), 1 literal path that
 * materialises the up-sampled images like the reference, 2 fused or fail.  All
 * modes produce identical motion fields. */
int qsvc_set_me_mode(qsvc_ctx *ctx, int mode);
/* Resident analysis with update_factor == 0: the motion estimation of level t+1 runs on a second
 * stream beside the decorrelate of level t (every level's inputs are frames of the resident clip).
 * 1 (default; env QSVC_OVERLAP): on, 0: one stream, strictly level by level.  Same results. */
int qsvc_set_overlap(qsvc_ctx *ctx, int on);
/* Same for decorrelate / correlate (env QSVC_MC_MODE): 1 literal path with materialised
 * int16 planes, 2 byte-plane fused path or fail, 0 automatic. */
int qsvc_set_mc_mode(qsvc_ctx *ctx, int mode);

/* GOP shards of a picture whose height is not a multiple of the block size (1080 lines, block 16)
 * are coupled through the prediction buffer of decorrelate / correlate: its rows below the last
 * whole block are never rewritten by predict() and carry the previous pair's in-place analysis
 * (decorrelate.cpp:562-567,852-861).  The state is
 * 3 components x (Y<<a - (Y/bs)*(bs<<a)) rows x (X<<a) bytes.  When a callback is installed,
 * every decorrelate / correlate level calls it twice on the calling thread:
 *   phase 0: `state` is zero-filled (what a fresh process starts from); a shard that is not the
 *            first writes the state received from its left neighbour and returns 1 (0: keep zeros);
 *   phase 1: `state` holds the state after this shard's last pair, to be passed to the right.
 * `level` is the temporal level of a resident analysis / synthesis (0 for the per-tool calls),
 * `synthesis` is 1 for correlate.  A negative return value aborts the call (QSVC_EINVAL). */
typedef int (*qsvc_tail_fn)(void *user, int level, int synthesis, int phase, uint8_t *state,
                            long long bytes);
int qsvc_set_tail_exchange(qsvc_ctx *ctx, qsvc_tail_fn fn, void *user);
/* Same exchange with `state` a DEVICE pointer (memory of the context's GPU, valid for the duration
 * of the call): the shards hand the state over GPU to GPU (NCCL point-to-point, peer copies)
 * without the host hop.  The context's stream is idle when the callback runs; the callback's own
 * transfers must have completed (phase 0) / may still read the buffer only until it returns
 * (phase 1). */
int qsvc_set_tail_exchange_device(qsvc_ctx *ctx, qsvc_tail_fn fn, void *user);

/* GOP shards with update_factor != 0: the frame two neighbouring shards
 * share receives the left shard's NEXT update first and the right shard's PREV update second
 * (update.cpp:109-140 clamps and truncates every contribution, so the order matters), in
 * update and in un_update.  When a callback is installed, every update / un_update level calls
 * it on the calling thread (level / inverse as for the tail exchange):
 *   phase 0 (right side, before its first frame is updated): `data` = 3 planes of
 *            pixels_in_y x align8(pixels_in_x) int16 to be filled with what the left neighbour
 *            passed in its phase 1; return 1 when filled, 0 for the first shard;
 *   phase 1 (left side, after its last frame received the NEXT update): `data` holds those
 *            planes, to be passed to the right; return 1 if a right neighbour exists, else 0;
 *   phase 2 (right side): `data` = the finished shared frame (I420 u8), to be passed to the left;
 *   phase 3 (left side, only after phase 1 returned 1): fill `data` with that frame, return 1.
 * The last frame of a shard is processed first, so a chain of shards does not serialise.  The
 * shards must run concurrently (one context each).  A negative return aborts the call. */
typedef int (*qsvc_boundary_fn)(void *user, int level, int inverse, int phase, void *data,
                                long long bytes);
int qsvc_set_boundary_exchange(qsvc_ctx *ctx, qsvc_boundary_fn fn, void *user);

/* Replaces `motion_estimate` main(), reference motion_estimate.cpp:490-912
 * (search: :70-184, pyramid driver: :260-413).
 * first_pair_is_global_first: 1 when even[0] is the first frame the reference
 * process would have read; GOP shards other than the first
 * pass 0. */
int qsvc_motion_estimate(qsvc_ctx *ctx, const uint8_t *even, const uint8_t *odd, int n_pairs,
                         int pixels_in_x, int pixels_in_y, int block_size, int border_size,
                         int search_range, int subpixel_accuracy,
                         int first_pair_is_global_first, int16_t *motion_out);

/* Replaces `decorrelate` main() (-D ANALYZE), reference decorrelate.cpp:199-1078
 * (predict: :69-189, I/B decision: :934-1027, entropy.cpp:20-34).
 * prediction_out may be NULL (the reference always writes prediction_<even_fn>). */
int qsvc_decorrelate(qsvc_ctx *ctx, const uint8_t *even, const uint8_t *odd,
                     const int16_t *motion_in, int n_pairs, int pixels_in_x, int pixels_in_y,
                     int block_size, int block_overlaping, int search_range,
                     int subpixel_accuracy, int always_B, uint8_t *high_out,
                     char *frame_types_out, int16_t *motion_out, uint8_t *prediction_out);

/* Replaces `correlate` main() (no ANALYZE), reference decorrelate.cpp:705-721,1029-1066. */
int qsvc_correlate(qsvc_ctx *ctx, const uint8_t *even, const uint8_t *high,
                   const int16_t *motion_in, const char *frame_types, int n_pairs,
                   int pixels_in_x, int pixels_in_y, int block_size, int block_overlaping,
                   int search_range, int subpixel_accuracy, uint8_t *odd_out,
                   uint8_t *prediction_out);

/* Replaces `update` (inverse=0: even_t -> low_t) and `un_update` (inverse=1:
 * low_t -> even_t) main(), reference update.cpp:158-684 (scatter: :71-148).
 * frames_in / frames_out hold n_pairs+1 frames. */
int qsvc_update(qsvc_ctx *ctx, int inverse, const uint8_t *frames_in, const uint8_t *high,
                const int16_t *motion, const char *frame_types, int n_pairs, int pixels_in_x,
                int pixels_in_y, int block_size, float update_factor, uint8_t *frames_out);

/* Motion-field (de)correlation: the step after the analysis / before the synthesis on the motion
 * side (motion_compress.py:141-182, motion_expand.py:147-179).  Fields are n_fields x
 * [PREV.X, PREV.Y, NEXT.X, NEXT.Y][blocks_in_y][blocks_in_x] int16.
 *
 * Replaces `bidirectional_motion_decorrelate` (inverse=0: NEXT -= PREV) and
 * `bidirectional_motion_correlate` (inverse=1: NEXT += PREV) main(), reference
 * bidirectional_motion_decorrelate.cpp:25-52,177-215. */
int qsvc_bidirectional_motion_decorrelate(qsvc_ctx *ctx, int inverse, const int16_t *fields_in,
                                          int n_fields, int blocks_in_y, int blocks_in_x,
                                          int16_t *fields_out);
/* Replaces `interlevel_motion_decorrelate` (inverse=0: residue = predicted - reference/2) and
 * `interlevel_motion_correlate` (inverse=1: predicted = residue + reference/2) main(), reference
 * interlevel_motion_decorrelate.cpp:32-69,250-297: field k pairs with reference field k/2
 * (the level above); `reference` may be NULL with n_reference = 0 (the tool falls back to
 * /dev/zero when the file is missing); a reference shorter than (n_fields+1)/2 fields repeats
 * its last field (what the reader's buffer still holds after a short fread). */
int qsvc_interlevel_motion_decorrelate(qsvc_ctx *ctx, int inverse, const int16_t *fields_in,
                                       int n_fields, const int16_t *reference, int n_reference,
                                       int blocks_in_y, int blocks_in_x, int16_t *fields_out);

/* The step after the hot path on the texture side: the distortion between
 * two frame files that psnr.py:78-90 obtains from the external program `snr --type=uchar --peak=255
 * --block_size=<bytes per picture>` (not part of the reference tree: parity unpinned).  Sum of squared
 * byte differences per block of block_bytes bytes, exact 64-bit integers;
 * PSNR = 10 log10(peak^2 * block_bytes / sse) is left to the caller. */
int qsvc_sse(qsvc_ctx *ctx, const uint8_t *a, const uint8_t *b, long long block_bytes, int n_blocks,
             unsigned long long *sse_out);

/* Whole-sequence temporal analysis with the frames kept resident in HBM between
 * levels: the device-side equivalent of analyze.py:107-153 driving
 * analyze_step.py:115-232 (split -> motion_estimate -> decorrelate -> update per
 * temporal level; split is index arithmetic on the resident frames).
 *
 *   load     : host low_0 (n_frames I420 frames) -> HBM
 *   analyze  : runs TRLs-1 levels on the resident frames, results stay in HBM
 *   fetch_*  : copies one level's results back to host buffers
 * n_frames must be GOPs * 2^(TRLs-1) + 1.  block_size_min follows analyze.py:118-151.
 */
typedef struct qsvc_analyze_params {
  int pixels_in_x, pixels_in_y;
  int TRLs;
  int block_size, block_size_min, border_size, block_overlaping;
  int search_range, subpixel_accuracy;
  int always_B;
  float update_factor;
  int first_gop_is_global_first; /* 1 unless this is a later GOP shard */
} qsvc_analyze_params;

int qsvc_resident_load(qsvc_ctx *ctx, const uint8_t *low0, int n_frames, int pixels_in_x,
                       int pixels_in_y);
int qsvc_resident_analyze(qsvc_ctx *ctx, const qsvc_analyze_params *params);
/* level t in [1, TRLs-1]; any output pointer may be NULL.  Sizes: high n_pairs(t)
 * frames, motion/motion_filtered n_pairs(t) fields, frame_types n_pairs(t) bytes,
 * low n_pairs(t)+1 frames. */
int qsvc_resident_fetch(qsvc_ctx *ctx, int level, uint8_t *high, int16_t *motion,
                        int16_t *motion_filtered, char *frame_types, uint8_t *low);
/* motion_residue_<level> of the resident analysis, computed on the device from the resident
 * motion_filtered fields the way motion_compress.py:141-182 chains the two tools: interlevel
 * against level+1 for level < TRLs-1, bidirectional for the top level. */
int qsvc_resident_fetch_motion_residue(qsvc_ctx *ctx, int level, int16_t *motion_residue);
/* Work done by the last qsvc_resident_analyze: SAD operations issued by the
 * search kernels and device milliseconds spent in them (summed over levels). */
int qsvc_resident_stats(qsvc_ctx *ctx, double *sad_ops, float *search_ms, float *total_ms);

/* analyze.py in one call with HOST buffers: uploads low_0, runs every level and
 * copies each level's results back on a second stream while the next level
 * computes.  outs has TRLs entries (outs[0] unused); NULL members are skipped.
 * Use qsvc_host_alloc (pinned memory) for the buffers so that the copies overlap. */
typedef struct qsvc_level_out {
  uint8_t *high;
  int16_t *motion;
  int16_t *motion_filtered;
  char *frame_types;
  uint8_t *low;
} qsvc_level_out;
int qsvc_analyze(qsvc_ctx *ctx, const qsvc_analyze_params *params, const uint8_t *low0,
                 int n_frames, const qsvc_level_out *outs);
void *qsvc_host_alloc(size_t bytes);
void qsvc_host_free(void *ptr);
/* Page-locks caller-owned host memory (e.g. a shared mapping several GOP shards write their
 * slices of the gathered sub-band files into) so that qsvc_analyze's copies land there directly. */
int qsvc_host_register(void *ptr, size_t bytes);
int qsvc_host_unregister(void *ptr);

/* Inverse: un_update -> correlate -> merge per level from TRLs-1 down to 1
 * (synthesize.py:95-153, synthesize_step.py:84-143), frames resident in HBM.
 * Inputs are pushed per level with qsvc_resident_push, then synthesize runs and
 * the reconstructed low_0 (GOPs*2^(TRLs-1)+1 frames) is read with fetch_low0. */
int qsvc_resident_push(qsvc_ctx *ctx, int level, int n_pairs, const uint8_t *high,
                       const int16_t *motion, const char *frame_types,
                       const uint8_t *low_top /* only for level == TRLs-1, else NULL */,
                       int pixels_in_x, int pixels_in_y, int block_size);
int qsvc_resident_synthesize(qsvc_ctx *ctx, const qsvc_analyze_params *params);
int qsvc_resident_fetch_low0(qsvc_ctx *ctx, uint8_t *low0, int n_frames);

#ifdef __cplusplus
}
#endif
#endif /* QSVC_B200_H */
