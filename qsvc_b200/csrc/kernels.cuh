// kernels.cuh -- launchers of the MCTF device kernels (defined in kernels_*.cu).
#pragma once
#include "common.cuh"

// Kernel classes for the per-class device timers (qsvc_profile_*).
enum {
  KC_IMG = 0,     // frame load/store, border fill, size fields, copies
  KC_DWT_ROWS,    // 5/3 row passes
  KC_DWT_COLS,    // 5/3 column passes
  KC_SEARCH,      // block search (SAD)
  KC_PREDICT,     // motion-compensated prediction (+ clip)
  KC_RESIDUE,     // residue / reconstruction / histograms
  KC_UPDATE,      // update lifting step
  KC_SEARCH_EXACT,  // sub-pixel search, exact int16 path for polluted / edge blocks
  KC_COUNT
};

struct Profiler {
  bool enabled = false;
  struct Rec { cudaEvent_t a, b; int cls; };
  Rec *recs = nullptr;
  int n = 0, cap = 0;
};

struct Launch {
  cudaStream_t stream;
  long long *counter;  // number of kernels launched (host side)
  Profiler *prof;
};

// RAII: brackets one kernel launch with two events when profiling is enabled.
struct ProfScope {
  const Launch &L;
  int idx;
  ProfScope(const Launch &l, int cls);
  ~ProfScope();
};

// ---- image preparation (kernels_img.cu) ----
// dst.row(slot0 + s, y)[x] = src[(f0 + s*fstep)*frame_stride + comp_off + y*w + x], y<h, x<w
void launch_load_u8(const Launch &L, Plane dst, int slot0, int nslots, const uint8_t *src,
                    long long frame_stride, long long comp_off, int f0, int fstep, int h, int w);
// glibc size fields in front of every physical row (heap-emulating planes only)
void launch_size_fields(const Launch &L, Plane p, int slot0, int nslots, int rows);
// texture::fill_border(data, y_dim, x_dim, b), the 8 regions in source order
void launch_fill_border(const Launch &L, Plane p, int slot0, int nslots, int y_dim, int x_dim,
                        int b);
// store the top-left h x w of a plane as u8 with the reference's truncating cast
void launch_store_u8(const Launch &L, Plane src, int slot0, int nslots, uint8_t *dst,
                     long long frame_stride, long long comp_off, int f0, int fstep, int h, int w);

// fill_border ring of a plain bordered plane straight from the frames' luma (one launch)
// fill_border of a compact plane (no row-pointer alias) from its own interior, one launch
void launch_ring_s16(const Launch &L, Plane p, int slot0, int nslots, int Y, int X, int b);
void launch_ring_u8(const Launch &L, Plane p, int slot0, int nslots, const uint8_t *src, long long frame_stride,
                    int f0, int Y, int X, int b);

void launch_region_copy(const Launch &L, Plane p, int slot0, int nslots, int h, int w, short *snap,
                        long long snap_slot_stride, int pitch, bool to_snapshot);

// ---- 5/3 transforms (kernels_dwt.cu), in place, reference Mallat layout ----
int dwt_init_attributes();
// one level on the top-left ny x nx of each slot
void launch_dwt_level(const Launch &L, Plane p, int slot0, int nslots, int ny, int nx,
                      bool synth);
// first analysis level of even-sized pictures straight from the frames' luma (load + rows + columns fused)
// one analysis level of the region whose int16 snapshot is `snap` (launch_region_copy), written in place: the
// four sub-bands from one pass over the snapshot (even sizes, even pitch)
bool dwt_snap_supported(int ny, int nx, int pitch, const short *snap, long long snap_slot_stride);
void launch_dwt_snap(const Launch &L, Plane p, int slot0, int nslots, const short *snap, long long snap_slot_stride,
                     int pitch, int ny, int nx);
// one synthesis level of the region whose copy is `snap`, written in place (same conditions as launch_dwt_snap)
void launch_syn_snap(const Launch &L, Plane p, int slot0, int nslots, const short *snap, long long snap_slot_stride,
                     int pitch, int ny, int nx);
void launch_dwt0_u8(const Launch &L, Plane p, int slot0, int nslots, const uint8_t *src, long long frame_stride,
                    int f0, int Y, int X);
// dwt2d::analyze(sig, y, x, levels) / dwt2d::synthesize(sig, y, x, levels)
void dwt_analyze(const Launch &L, Plane p, int slot0, int nslots, int y, int x, int levels);
void dwt_synthesize(const Launch &L, Plane p, int slot0, int nslots, int y, int x, int levels);

// ---- motion search (kernels_me.cu) ----
enum { ME_INIT = 0, ME_DESCEND = 1, ME_SUBPEL = 2 };
struct SearchParams {
  Plane img;
  const int *slots;  // per pair: R0 slot, R1 slot, P slot
  const short *mv_in;
  short *mv_out;
  int BY, BX;    // full field dimensions (row stride of a motion plane = BX)
  int nby, nbx;  // blocks searched at this level
  int bs, bd;    // block size and block border at this level
  int mode, lim;
  // optional: the same level as dense byte planes (level 0 of an invertible pyramid: the frames' luma).
  // Blocks whose two windows lie inside the picture are searched on bytes (VABSDIFF4); nullptr: never.
  const uint8_t *v0 = nullptr;
  long long v0_slot_stride = 0;
  int v0_pitch = 0, v0_Y = 0, v0_X = 0;
};
void launch_search(const Launch &L, const SearchParams &q, int npairs);
// sub-pixel search levels without materialised up-sampled images (kernels_subpel.cu)
// level-0 tiles of the "holds a non-byte sample" map are (1 << SUBPEL_TILE_SHIFT) pixels square: small
// tiles keep isolated out-of-range pixels from sending whole neighbourhoods to the exact generator
static constexpr int SUBPEL_TILE_SHIFT = 2;
struct SubpelParams {
  Plane b0;            // compact level-0 buffers after the over-pixel descent
  const int *slots;    // per pair: R0 slot, R1 slot, P slot
  const uint8_t *v;    // V_l planes (u8, one per slot) of this level
  long long v_slot_stride;
  int v_pitch;
  const uint8_t *tile_bad;  // per slot, per level-0 tile: holds a sample outside [0,255]
  int tiles_x, tiles_per_slot;
  const short *mv_in;
  short *mv_out;
  int BY, BX;
  int l;               // sub-pixel level, 1 or 2
  int Y, X;            // level-0 picture size
  int B, Bc;           // fill border (search_range + border_size), compact plane border
  int Ya, Ba;          // rows / border of the reference's allocation (size-field rows)
  unsigned long long size_field;
  int lim;             // vector clamp, search_range << a
  int *slow_count;       // blocks for the strip path (windows touch polluted strips / leave the picture)
  int *slow_list;
  int *bad_count;        // blocks over a non-byte tile: full exact generator
  int *bad_list;
  const short *strip_top;   // per slot [clean][X << l]: level-l samples of the polluted top rows
  const short *strip_left;  // per slot [(Y << l) - clean][clean]: polluted left columns below them
  long long strip_top_stride, strip_left_stride;
  int clean;             // min((2B + 2) << (l - 1), Y << l, X << l)
  int v_rows_per_slot;   // v_slot_stride / v_pitch
  int use_tma;           // tm_p / tm_r hold valid CUtensorMap objects; 1: windows at 16-byte aligned columns, 2: at their own column
  int debug;             // env QSVC_SUBPEL_DEBUG: 1 = fast blocks go to the strip kernel, 2 = strip blocks go to the exact generator
  int npairs, pair_group;  // TMA kernel: block order (launch_subpel fills these)
  int check_tiles;       // 0: every tile of every slot is known to hold bytes (tile_bad not consulted)
  unsigned kmul[3];      // 2^24, 2^16, 2^8: funnel shifts as multiplies on the FMA pipe (filled by launch_subpel)
  alignas(64) unsigned char tm_p[128];
  alignas(64) unsigned char tm_r[128];
};
bool subpel_make_tensor_maps(const uint8_t *v, int pitch, long long total_rows, int W, int box_w, void *tm_p,
                             void *tm_r);
bool subpel_supported(int W);
// int16 strips of the level-l images where they differ from the byte planes (first
// `clean` rows and columns); level 2 is derived from the level-1 strips and V_1.
void launch_strips(const Launch &L, const SubpelParams &q, int level, int slot0, int nslots, short *top, short *left,
                   const short *top1, const short *left1, long long top1_stride, long long left1_stride,
                   int clean1, const uint8_t *v1, long long v1_slot_stride, int v1_pitch);
void launch_luma_to_plane(const Launch &L, const uint8_t *src, long long frame_stride, int nframes, int Y, int X,
                          uint8_t *dst, long long dst_slot_stride, int pitch);
int subpel_tma_timeouts();
void launch_subpel(const Launch &L, const SubpelParams &q, int W, int npairs);
void launch_plane_to_u8(const Launch &L, Plane src, int slot0, int nslots, int Y, int X,
                        uint8_t *dst, long long dst_slot_stride, int pitch, uint8_t *tile_bad,
                        int tiles_x, int tiles_per_slot);
void launch_upsample2x(const Launch &L, const uint8_t *in, int n, int m, int pitch_in,
                       long long in_slot_stride, uint8_t *out, int pitch_out,
                       long long out_slot_stride, int nslots);
// nst (2 or 3) chained x2 up-samplings in one kernel; outs[k] (k < nst) = plane after k + 1 stages
// or nullptr when nobody reads it (the last one is mandatory); (m << 1) % 8 == 0
void launch_upsample_chain(const Launch &L, const uint8_t *in, int n, int m, int pitch_in, long long in_slot_stride,
                           int nst, uint8_t *const outs[3], const int pitches[3], const long long strides[3],
                           int nslots);
int run_int_peak(cudaStream_t stream, unsigned *d_out, int blocks, int iters, bool packed);

// ---- motion compensation (kernels_mc.cu) ----
struct PredictParams {
  Plane ref;          // dense up-sampled references: slot = ref_slot*3 + c
  Plane pred;         // dense prediction planes: slot = c
  const short *mv;    // one field (4 planes of BY*BX)
  int r0_slot, r1_slot;
  int BY, BX, bsa;    // block size << a
  int Ya, Xa, ba;     // up-sampled size and border << a
  int padh;           // (malloc chunk - row bytes)/2 in shorts (heap alias rule)
};
void launch_predict(const Launch &L, const PredictParams &q);
// block_overlaping > 0: per-block analysis + sub-band scatter (the caller synthesises the picture)
bool predict_obmc_supported(int bsa, int ova);
void launch_predict_obmc(const Launch &L, const PredictParams &q, int ova, int levels);
// clip to [0,255] everything outside the covered area [0,cy) x [0,cx)
void launch_clip_uncovered(const Launch &L, Plane pred, int Ya, int Xa, int cy, int cx);

struct ResidueParams {
  Plane pred;            // LL in the top-left of slot c
  const uint8_t *odd;    // analysis: odd frame; synthesis: high frame
  uint8_t *out;          // analysis: high frame ('B' variant); synthesis: odd frame
  uint8_t *prediction;   // nullable: prediction_<even> side output
  int *hist;             // nullable: [0,256) predicted luma, [256,512) residue luma + 128
  int X, Y;
  int synth;             // 0: decorrelate, 1: correlate
  int is_I;              // synthesis only: frame type of this pair
};
void launch_residue(const Launch &L, const ResidueParams &q);
// histogram of one motion field (+128), out-of-range components counted in hist[256]
void launch_mv_hist(const Launch &L, const short *mv, int n, int *hist);
void launch_mv_bidirectional(const Launch &L, const short *in, short *out, int n_fields, int plane, int inverse);
void launch_mv_interlevel(const Launch &L, const short *in, const short *ref, short *out, int n_fields, int n_ref,
                          int plane, int inverse);
// out[k] += sum of squared byte differences of block k (out zeroed by the caller)
void launch_sse_u8(const Launch &L, const uint8_t *a, const uint8_t *b, long long block, int nblocks,
                   unsigned long long *out);
void launch_copy_bytes(const Launch &L, void *dst, const void *src, size_t n);
void launch_copy_strided(const Launch &L, void *dst, long long dst_stride, const void *src,
                         long long src_stride, size_t n, int count);

struct UpdateParams {
  Plane ref;            // dense, luma-sized planes: slot = c (the frame being updated)
  Plane res;            // residue planes: slot = c (high - 128, chroma top-left only)
  const short *mv;      // the pair's field
  int dir;              // MV_PREV_X / MV_NEXT_X: which vectors displace the residue
  int BY, BX, bs, Y, X;
  float uf;
  int inverse;
  const int *reach;     // device: max |vector component| of this direction (launch_mv_reach)
};
void launch_update(const Launch &L, const UpdateParams &q);
// all frames [frame0, frame0 + nframes) of a level in one launch (frames are independent of each other)
struct UpdateBatchParams {
  Plane ref;                 // dense luma-sized planes: slot = c * slots_per_comp + (frame - frame0)
  int slots_per_comp, frame0;
  const uint8_t *high;       // high frames of the level's pairs (I420); residue = high - 128
  long long high_stride;
  const short *mv;           // the level's fields
  const char *types;         // device: frame types of the pairs
  int *cnt, *list, *reach;   // per (pair, direction): per tile block count and ids; largest |vector component|
  int4 *geo = nullptr;       // optional, parallel to list: (displaced origin y, x, source origin y, x) of the listed blocks
  int cap;                   // list capacity per tile
  int n_pairs, BY, BX, bs, Y, X, tiles_x, tiles_y;
  float uf;
  int inverse;
  // dyadic kernel only: the luma targets are the bytes of the frames themselves (frame f at luma_in + f *
  // luma_in_stride, result to luma_out likewise) instead of int16 planes; nullptr: planes for all components
  const uint8_t *luma_in = nullptr;
  uint8_t *luma_out = nullptr;
  long long luma_in_stride = 0, luma_out_stride = 0;
};
bool update_is_dyadic(float uf);  // launch_update_batch takes the integer kernel (which honours luma_in / luma_out)
// chroma component of frames f0 .. f0 + n - 1 (bytes, (Y/2) x (X/2) at src + f * frame_stride + comp_off) -> planes
// slot0 .. : the one-level zero-high-band 5/3 synthesis to Y x X (update.cpp's 4:2:0 -> 4:4:4), columns, then rows
void launch_chroma_up_s16(const Launch &L, Plane dst, int slot0, int n, const uint8_t *src, long long frame_stride,
                          long long comp_off, int f0, int Y, int X);
// ... and back: the LL band of the one-level 5/3 analysis (rows, then columns) of planes slot0 .. stored as bytes
// (truncating store, texture.cpp:139-141)
void launch_ll1_store_u8(const Launch &L, Plane src, int slot0, int n, uint8_t *dst, long long frame_stride,
                         long long comp_off, int f0, int Y, int X);
void launch_update_bin(const Launch &L, const UpdateBatchParams &q);  // cnt and reach zeroed by the caller
void launch_update_batch(const Launch &L, const UpdateBatchParams &q, int nframes);
void launch_mv_reach(const Launch &L, const short *mv, int n, int *out);
// residue plane: top-left h x w = high - 128 (rest untouched)
void launch_load_residue(const Launch &L, Plane dst, int slot, const uint8_t *src, int h, int w);

// ---- byte-plane motion compensation (kernels_mcfused.cu) ----
struct PredU8Params {
  const uint8_t *v;          // V_a planes, [even frame][component], (Ya + 2) x v_pitch each
  long long v_plane_stride;  // bytes between consecutive planes
  int v_pitch;
  uint8_t *p;                // P_a planes out, [pair][component]
  long long p_plane_stride;
  int p_pitch;
  const short *mv;           // fields of the pairs
  int f0;                    // V frame index of pair 0's reference[0]
  int BY, BX, bsa, Ya, Xa, ba, padh;
  int by0, nby;              // block rows [by0, by0 + nby) are predicted (nby == 0: all)
};
void launch_predict_u8(const Launch &L, const PredU8Params &q, int npairs);

// ---- line-based decorrelate / correlate (kernels_mcmarch.cu) ----
struct MarchParams {
  const uint8_t *v;          // V_a planes, [even frame][component], (Ya + 2) x v_pitch each
  long long v_plane_stride;
  int v_pitch;
  int f0;                    // V frame index of pair 0's reference[0]
  const uint8_t *tail;       // nullable: planes [pair][component] holding the prediction rows >= cy,
  long long tail_plane_stride;  //   addressed by absolute row (pointer pre-offset), pitch v_pitch
  const short *mv;
  const uint8_t *in;         // analysis: odd frames; synthesis: high frames (I420, per pair)
  long long in_stride;
  uint8_t *out;              // analysis: high frames ('B' variant); synthesis: odd frames
  long long out_stride;
  uint8_t *prediction;       // nullable: prediction_<even> side output
  long long pred_stride;
  int *hist;                 // nullable: per pair [0,256) predicted luma, [256,512) residue + 128
  int hist_stride;
  const char *types;         // synthesis: frame types (device)
  int X, Y, a, synth;
  int BY, BX, bsa, Ya, Xa, ba, padh, cy;
  int ring;                  // samples of materialised border around every V plane (launch_fill_ring), 0: none
  int bs_shift, nstrips, nsegs, seg_p;  // filled by the launcher
};
void launch_mc_march(const Launch &L, MarchParams q, int npairs);
// same contract, banded shared-memory pipeline (kernels_mctile.cu); a in {1, 2}, (bs << a) % 16 == 0
bool mc_tile_supported(int a, int bsa, int X);
void launch_mc_tile(const Launch &L, MarchParams q, int npairs);

// border ring of the reference's border rule around the interior of nplanes byte planes (U = interior origin)
void launch_fill_ring(const Launch &L, uint8_t *U, long long plane_stride, int pitch, int nplanes, int Yd, int Xd,
                      int ring, int b, int padh);
void launch_tail_state(const Launch &L, const uint8_t *P, long long plane_stride, int pitch,
                       uint8_t *Pnext, int Ya, int Xa, int cy, int first_comp, int ncomp);
void launch_copy_rows(const Launch &L, const uint8_t *src, uint8_t *dst, long long plane_stride,
                      int pitch, int row0, int rows, int width);
