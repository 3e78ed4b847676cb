// api.cu -- context, per-level orchestration and the C ABI of libqsvc_b200.so.
//
// Host-side control flow mirrors the reference tools' main() functions
// (motion_estimate.cpp:714-907, decorrelate.cpp:508-1075, update.cpp:439-679)
// and the Python drivers (analyze.py:107-153, synthesize.py:95-153); all pixel
// work runs in the kernels of kernels_*.cu on the context's CUDA stream.
#include <math.h>

#include <memory>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/qsvc_b200.h"
#include "kernels.cuh"

// ------------------------------------------------------------------ errors

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(QSVC_ECUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call,               \
                  cudaGetErrorString(e_));                                              \
  } while (0)
#define TRY(call)            \
  do {                       \
    int rc_ = (call);        \
    if (rc_ != QSVC_OK) return rc_; \
  } while (0)

// ----------------------------------------------------------------- context

struct PoolBlock {
  void *ptr;
  size_t size;
  bool used;
  int lane;  // stream lane whose work last used the block: a free block is only handed back to that lane
};

struct LevelResult {
  int n_pairs = 0, block_size = 0, search_range = 0;
  uint8_t *high = nullptr;      // n_pairs frames
  short *motion = nullptr;      // n_pairs fields
  short *motion_filtered = nullptr;
  uint8_t *low = nullptr;       // n_pairs+1 frames
  std::string types;
};

struct qsvc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  // second compute lane: with update_factor == 0 every level's inputs are frames of the resident clip,
  // so the motion estimation of level t+1 runs beside the decorrelate of level t (env QSVC_OVERLAP=0: off)
  cudaStream_t me_stream = nullptr;
  std::vector<cudaEvent_t> me_events;
  int lane = 0;      // lane of the pool blocks handed out right now (0: `stream`, 1: `me_stream`)
  cudaEvent_t mv_ready = nullptr;  // decorrelate: wait for this event before the first use of the motion field
  int overlap = 1;
  std::vector<cudaEvent_t> level_events;
  std::vector<cudaEvent_t> upload_events;  // qsvc_analyze: one per GOP of the clip being uploaded
  int upload_gops = 0, upload_gop_frames = 0;  // > 0: level 1 may start segment by segment behind the upload
  std::vector<int> upload_segs;                // end pair (exclusive) of every upload segment at level 1
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  long long launches = 0;
  Profiler prof;
  std::vector<PoolBlock> pool;
  qsvc_tail_fn tail_fn = nullptr;  // GOP-shard exchange of the prediction tail state (A.2.6)
  void *tail_user = nullptr;
  int tail_on_device = 0;  // the callback's `state` is a device pointer (qsvc_set_tail_exchange_device)
  qsvc_boundary_fn boundary_fn = nullptr;  // GOP-shard exchange of the boundary frame (update_factor != 0)
  void *boundary_user = nullptr;
  int cur_level = 0;  // temporal level of the running resident analysis / synthesis
  int tma_mode = 1;  // 0: plain loads in the sub-pixel fast path (env QSVC_TMA=0)
  int mc_mode = 0;  // same three values for the decorrelate / correlate path
  int me_fuse0 = 1;  // fused ME path: first pyramid level straight from the frames (env QSVC_ME_FUSE0=0: off)
  int me_bytes0 = 1;  // fused ME path: level-0 search of invertible pyramids on byte planes (env QSVC_ME_BYTES0=0: off)
  int mc_ring = 1;  // byte-plane path: materialised border ring around the reference planes (env QSVC_MC_RING=0: off)
  int mc_kernel = 0;  // byte-plane path: 0 register march (k_mc_march, default: faster), 1 banded shared-memory pipeline (k_mc_tile); env QSVC_MC_KERNEL
  int me_mode = 0;  // 0: automatic, 1: literal (materialised) path only, 2: fused path required
  size_t me_budget = (size_t)40 << 30;  // bytes of HBM for the ME image planes of one chunk
  // resident sequence
  uint8_t *low0 = nullptr;
  int n_frames = 0, X = 0, Y = 0;
  std::vector<LevelResult> levels;  // index t (0 unused)
  double sad_ops = 0;
  float search_ms = 0, total_ms = 0;
  Launch L() { return Launch{stream, &launches, &prof}; }
};

static int pool_alloc(qsvc_ctx *c, size_t bytes, void **out) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (bytes == 0) bytes = 256;
  int best = -1;
  for (size_t i = 0; i < c->pool.size(); i++)
    if (!c->pool[i].used && c->pool[i].lane == c->lane && c->pool[i].size >= bytes &&
        (best < 0 || c->pool[i].size < c->pool[best].size))
      best = (int)i;
  if (best >= 0 && c->pool[best].size <= bytes * 2 + (1 << 20)) {
    c->pool[best].used = true;
    *out = c->pool[best].ptr;
    return QSVC_OK;
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    // release cached blocks and retry once
    for (auto &b : c->pool)
      if (!b.used && b.ptr) {
        cudaFree(b.ptr);
        b.ptr = nullptr;
        b.size = 0;
      }
    cudaGetLastError();
    e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess)
      return fail(QSVC_ENOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
  }
  c->pool.push_back(PoolBlock{p, bytes, true, c->lane});
  *out = p;
  return QSVC_OK;
}

static void pool_free(qsvc_ctx *c, void *p) {
  if (!p) return;
  for (auto &b : c->pool)
    if (b.ptr == p) {
      b.used = false;
      return;
    }
}

// Scoped set of pool allocations released on scope exit.
struct Scratch {
  qsvc_ctx *c;
  std::vector<void *> ptrs;
  explicit Scratch(qsvc_ctx *ctx) : c(ctx) {}
  ~Scratch() {
    for (void *p : ptrs) pool_free(c, p);
  }
  int get(size_t bytes, void **out) {
    int rc = pool_alloc(c, bytes, out);
    if (rc == QSVC_OK) ptrs.push_back(*out);
    return rc;
  }
  // hands the block over to the caller (it outlives the scope)
  void detach(void *p) { ptrs.erase(std::remove(ptrs.begin(), ptrs.end(), p), ptrs.end()); }
};

// ---------------------------------------------------------------- geometry

static inline long long frame_bytes(int X, int Y) {
  return (long long)X * Y + 2LL * (X / 2) * (Y / 2);
}
static inline long long comp_offset(int X, int Y, int c) {
  return c == 0 ? 0 : (long long)X * Y + (long long)(c - 1) * (X / 2) * (Y / 2);
}
static inline int desp(int x, int l) {  // motion_estimate.cpp:232-236
  for (int i = 0; i < l; i++) x = (x + 1) / 2;
  return x;
}
static inline int me_levels(int sr) {  // motion_estimate.cpp:277
  return (int)rint(log((double)sr) / log(2.0)) - 1;
}

// glibc chunk geometry of one texture row (see common.cuh Plane)
static inline int heap_row_shorts(int x_dim, int b) {
  long long W = (long long)x_dim + 2LL * b;
  long long chunk = (2 * W + 8 + 15) & ~15LL;
  if (chunk < 32) chunk = 32;
  return (int)(chunk / 2);
}

struct PlaneAlloc {
  Plane p;
  size_t bytes;
  short *raw;
};

// heap-emulating plane set (ME)
static int alloc_heap_planes(Scratch &s, int nslots, int y_dim, int x_dim, int b, PlaneAlloc *out) {
  int S = heap_row_shorts(x_dim, b);
  long long rows = (long long)y_dim + 2LL * b;
  long long slot_shorts = (rows + 2) * S;
  size_t bytes = (size_t)slot_shorts * nslots * sizeof(short);
  void *raw;
  TRY(s.get(bytes, &raw));
  out->raw = (short *)raw;
  out->bytes = bytes;
  out->p.base = (short *)raw + S;
  out->p.slot_stride = slot_shorts;
  out->p.S = S;
  out->p.y_dim = y_dim;
  out->p.x_dim = x_dim;
  out->p.b = b;
  return QSVC_OK;
}

// dense plane set (b = 0: no heap emulation)
static int alloc_dense_planes(Scratch &s, int nslots, int y_dim, int x_dim, PlaneAlloc *out) {
  int S = (x_dim + 7) & ~7;
  long long slot_shorts = (long long)y_dim * S;
  size_t bytes = (size_t)slot_shorts * nslots * sizeof(short);
  void *raw;
  TRY(s.get(bytes, &raw));
  out->raw = (short *)raw;
  out->bytes = bytes;
  out->p.base = (short *)raw;
  out->p.slot_stride = slot_shorts;
  out->p.S = S;
  out->p.y_dim = y_dim;
  out->p.x_dim = x_dim;
  out->p.b = 0;
  return QSVC_OK;
}

static int check_geometry(int X, int Y, int bs, int a) {
  if (X < 2 || Y < 2 || X > 16384) return fail(QSVC_EINVAL, "unsupported picture size %dx%d", X, Y);
  if (bs < 1 || bs > X || bs > Y) return fail(QSVC_EINVAL, "unsupported block_size %d", bs);
  if (a < 0 || a > 4 || ((long long)X << a) > 16384)
    return fail(QSVC_EINVAL, "unsupported subpixel_accuracy %d for width %d (line limit 16384, texture.cpp:9)", a, X);
  return QSVC_OK;
}

// ------------------------------------------------------- motion estimation

// SAD operations of one pair
static double me_sad_ops(int BY, int BX, int bs, int bd, int L, int a) {
  double ops = 0;
  double w = bs + 2.0 * bd;
  for (int l = 0; l <= L; l++) ops += w * w * desp(BY, l) * desp(BX, l);
  for (int l = 1; l <= a; l++) {
    double wl = (double)(bs << l) + 2.0 * (bd >> l);
    ops += wl * wl * BY * BX;
  }
  return 18.0 * ops;
}

// ---- fused ME path (a in {1,2}): compact pyramid planes + u8 interpolated planes
// + packed-byte sub-pixel search (kernels_subpel.cu).  Same results as the literal
// path below; chosen automatically when the geometry allows it.
static bool me_fused_ok(int X, int Y, int bs, int bd, int sr, int a) {
  if (a < 1 || a > 2 || bd != 0) return false;
  for (int l = 1; l <= a; l++)
    if (!subpel_supported(bs << l)) return false;
  const int B = sr + bd;
  if (B > X || B > Y) return false;                 // second synthesis must see zero high bands
  if (Y + B + 2 > ((Y - B) << a)) return false;     // over-pixel search must stay clear of the
  if (X % 4 != 0) return false;                     //   un-shifted rows; word-aligned block columns
  return true;
}

// dwt_analyze (dwt2d.cpp:76-119) with a scratch copy: a level whose region has even sizes is copied out and
// transformed back in one pass (k_dwt_snap: rows and columns, no in-place hazard); other levels run the generic
// row and column passes.  tmp holds one region of y x x per slot (slot stride tmp_stride shorts, pitch tmp_pitch).
static void dwt_analyze_via_copy(const Launch &Lh, Plane img, int slot0, int nslots, int y, int x, int levels,
                                 short *tmp, long long tmp_stride, int tmp_pitch) {
  for (int lv = 0; lv < levels; lv++) {
    const int nx = x, ny = y;
    x >>= 1;
    y >>= 1;
    if (y == 0) y = 1;
    if (x == 0) x = 1;
    if (tmp && dwt_snap_supported(ny, nx, tmp_pitch, tmp, tmp_stride)) {
      launch_region_copy(Lh, img, slot0, nslots, ny, nx, tmp + (long long)slot0 * tmp_stride, tmp_stride, tmp_pitch, true);
      launch_dwt_snap(Lh, img, slot0, nslots, tmp, tmp_stride, tmp_pitch, ny, nx);
    } else {
      launch_dwt_level(Lh, img, slot0, nslots, ny, nx, false);
    }
  }
}

// one synthesis level (dwt_synthesize with levels = 1 on a ny x nx region) through the same scratch copy
static void dwt_synthesize1_via_copy(const Launch &Lh, Plane img, int slot0, int nslots, int ny, int nx, short *tmp,
                                     long long tmp_stride, int tmp_pitch) {
  if (nx == 0) nx = 1;
  if (ny == 0) ny = 1;
  if (tmp && dwt_snap_supported(ny, nx, tmp_pitch, tmp, tmp_stride)) {
    launch_region_copy(Lh, img, slot0, nslots, ny, nx, tmp + (long long)slot0 * tmp_stride, tmp_stride, tmp_pitch, true);
    launch_syn_snap(Lh, img, slot0, nslots, tmp, tmp_stride, tmp_pitch, ny, nx);
  } else {
    launch_dwt_level(Lh, img, slot0, nslots, ny, nx, true);
  }
}

static int me_level_fused(qsvc_ctx *c, const uint8_t *even, long long even_stride,
                          const uint8_t *odd, long long odd_stride, int n_pairs, int X, int Y,
                          int bs, int sr, int a, int L, bool pr, int first_global, short *mv_out) {
  const int BY = Y / bs, BX = X / bs;
  const int B = sr, Bc = (B + 2 + 7) & ~7;  // fill ring + >= 2 zero cells; multiple of 8: 16-byte aligned rows
  const long long field = 4LL * BY * BX;
  const int n_search = 1 + L + a;
  Launch Lh = c->L();
  // per-slot footprints
  const int S = (X + 2 * Bc + 7) & ~7;
  const size_t slot_shorts = (size_t)(Y + 2 * Bc + 1) * S;
  int pitch[3];
  size_t vbytes[3], vtotal = 0;
  for (int l = 0; l <= a; l++) {
    pitch[l] = (((X << l) + 15) & ~15) + 16;
    vbytes[l] = (size_t)pitch[l] * ((Y << l) + 2);
    vtotal += vbytes[l];
  }
  const size_t per_slot = slot_shorts * sizeof(short) * 3 + vtotal;  // planes + LL snapshots (< 1.34x)
  const int per_pair_slots = pr ? 2 : 3;
  long long max_pairs = (long long)(c->me_budget / per_slot - 1) / per_pair_slots;
  if (max_pairs < 1) max_pairs = 1;

  for (int i0 = 0; i0 < n_pairs; i0 += (int)max_pairs) {
    const int m = (int)std::min<long long>(max_pairs, n_pairs - i0);
    Scratch s(c);
    const int n_copy = pr ? 0 : m - 1;
    const int nslots = 2 * m + 1 + n_copy;
    short *raw;
    TRY(s.get(slot_shorts * nslots * sizeof(short), (void **)&raw));
    Plane img;
    img.base = raw;
    img.slot_stride = (long long)slot_shorts;
    img.S = S;
    img.y_dim = Y + 2 * Bc;  // every row pointer is "shifted": plain bordered image
    img.x_dim = X;
    img.b = Bc;
    uint8_t *v[3] = {nullptr, nullptr, nullptr};
    for (int l = 0; l <= a; l++) TRY(s.get(vbytes[l] * nslots, (void **)&v[l]));
    short *mv_tmp;
    int *d_slots, *d_slow;
    uint8_t *d_flags;
    const int TSZ = 1 << SUBPEL_TILE_SHIFT;
    const int tiles_x = (X + TSZ - 1) / TSZ, tiles_per_slot = tiles_x * ((Y + TSZ - 1) / TSZ);
    TRY(s.get((size_t)m * field * sizeof(short), (void **)&mv_tmp));
    TRY(s.get((size_t)m * 3 * sizeof(int), (void **)&d_slots));
    TRY(s.get((size_t)nslots * tiles_per_slot + 16, (void **)&d_flags));
    TRY(s.get(((size_t)m * BY * BX + 1) * sizeof(int), (void **)&d_slow));
    int *d_bad;
    TRY(s.get(((size_t)m * BY * BX + 1) * sizeof(int), (void **)&d_bad));
    // int16 strips of the sub-pixel images where fill_border's replicas reach (top rows, left columns)
    int cleanl[3] = {0, 0, 0};
    size_t top_sz[3] = {0, 0, 0}, left_sz[3] = {0, 0, 0};
    short *stop[3] = {nullptr, nullptr, nullptr}, *sleft[3] = {nullptr, nullptr, nullptr};
    for (int l = 1; l <= a; l++) {
      cleanl[l] = std::min((2 * B + 2) << (l - 1), std::min(Y << l, X << l));
      top_sz[l] = (size_t)cleanl[l] * (X << l);
      left_sz[l] = std::max<size_t>((size_t)((Y << l) - cleanl[l]) * cleanl[l], 8);
      TRY(s.get(top_sz[l] * nslots * sizeof(short), (void **)&stop[l]));
      TRY(s.get(left_sz[l] * nslots * sizeof(short), (void **)&sleft[l]));
    }
    std::vector<int> slots(3 * m);
    for (int i = 0; i < m; i++) {
      slots[3 * i] = (!pr && i >= 1) ? 2 * m + i : i;
      slots[3 * i + 1] = i + 1;
      slots[3 * i + 2] = m + 1 + i;
    }
    CU(cudaMemcpyAsync(d_slots, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(raw, 0, slot_shorts * nslots * sizeof(short), c->stream));
    CU(cudaMemsetAsync(d_flags, 0, (size_t)nslots * tiles_per_slot + 16, c->stream));
    // invertible pyramid of an even-sized picture: the first analysis level is computed straight
    // from the frames (load, row pass and column pass fused), the border ring likewise, and the
    // descent restores level 0 by loading the frames again -- no level-0 snapshot
    const bool fuse0 = pr && L > 0 && (Y % 2 == 0) && c->me_fuse0 != 0;
    if (fuse0) {
      launch_ring_u8(Lh, img, 0, m + 1, even, even_stride, i0, Y, X, B);
      launch_dwt0_u8(Lh, img, 0, m + 1, even, even_stride, i0, Y, X);
      launch_dwt0_u8(Lh, img, m + 1, m, odd, odd_stride, i0, Y, X);
    } else {
      launch_load_u8(Lh, img, 0, m + 1, even, even_stride, 0, i0, 1, Y, X);
      launch_load_u8(Lh, img, m + 1, m, odd, odd_stride, 0, i0, 1, Y, X);
      launch_ring_s16(Lh, img, 0, m + 1, Y, X, B);  // compact planes: fill_border without the alias, one launch
    }
    if (n_copy > 0) {
      launch_load_u8(Lh, img, 2 * m + 1, n_copy, even, even_stride, 0, i0 + 1, 1, Y, X);
      launch_ring_s16(Lh, img, 2 * m + 1, n_copy, Y, X, B);
    }
    // non-invertible pyramid: scratch copy of one level-0 region per slot for the analysis passes
    short *tmpa = nullptr;
    const int tmpa_pitch = (X + 7) & ~7;
    const long long tmpa_stride = (long long)Y * tmpa_pitch;
    if (!pr && L > 0) TRY(s.get((size_t)tmpa_stride * nslots * sizeof(short), (void **)&tmpa));
    if (!pr) {
      // carried reference[0]: one pass of the non-invertible pyramid (the sub-pixel
      // synthesis/analysis pair of the reference is an exact identity and is skipped)
      auto used_state = [&](int slot0, int n) {
        if (n <= 0) return;
        dwt_analyze_via_copy(Lh, img, slot0, n, Y, X, L, tmpa, tmpa_stride, tmpa_pitch);
        for (int l = L - 1; l >= 0; --l)
          dwt_synthesize1_via_copy(Lh, img, slot0, n, desp(Y, l), desp(X, l), tmpa, tmpa_stride, tmpa_pitch);
      };
      if (!(first_global && i0 == 0)) used_state(0, 1);
      used_state(2 * m + 1, n_copy);
    }
    short *bufs[2] = {mv_out + (long long)i0 * field, mv_tmp};
    int j = 0;
    // invertible pyramid: level 0 is the frames' luma, which the sub-pixel levels need as byte planes
    // anyway -- made here, so that the level-0 search can run on bytes too
    const bool bytes0 = pr && bs == 16 && c->me_bytes0 != 0;
    if (pr) {
      launch_luma_to_plane(Lh, even + (long long)i0 * even_stride, even_stride, m + 1, Y, X, v[0],
                           (long long)vbytes[0], pitch[0]);
      launch_luma_to_plane(Lh, odd + (long long)i0 * odd_stride, odd_stride, m, Y, X,
                           v[0] + (size_t)(m + 1) * vbytes[0], (long long)vbytes[0], pitch[0]);
    }
    auto run_search = [&](int mode, int nby, int nbx, int lim, bool level0 = false) {
      SearchParams q;
      if (level0 && bytes0) {
        q.v0 = v[0];
        q.v0_slot_stride = (long long)vbytes[0];
        q.v0_pitch = pitch[0];
        q.v0_Y = Y;
        q.v0_X = X;
      }
      q.img = img;
      q.slots = d_slots;
      q.mv_out = bufs[(j + n_search - 1) & 1];
      q.mv_in = bufs[(j + n_search) & 1];
      q.BY = BY;
      q.BX = BX;
      q.nby = nby;
      q.nbx = nbx;
      q.bs = bs;
      q.bd = 0;
      q.mode = mode;
      q.lim = lim;
      launch_search(Lh, q, m);
      j++;
    };
    // For an invertible pyramid the descent only restores what the analysis overwrote:
    // snapshot each LL region before it is transformed and copy it back instead of
    // running the inverse transform (identical result, half the passes).
    short *snap = nullptr;
    std::vector<size_t> snap_off(L + 1, 0);
    std::vector<int> snap_pitch(L + 1, 0);
    size_t snap_per_slot = 0;
    if (pr && L > 0) {
      const int l0 = fuse0 ? 1 : 0;  // first level that needs a snapshot
      for (int l = l0; l < L; l++) {
        snap_off[l] = snap_per_slot;
        snap_pitch[l] = ((X >> l) + 7) & ~7;
        snap_per_slot += (size_t)(Y >> l) * snap_pitch[l];
      }
      TRY(s.get(std::max<size_t>(snap_per_slot, 8) * nslots * sizeof(short), (void **)&snap));
      for (int l = l0; l < L; l++) {
        launch_region_copy(Lh, img, 0, nslots, Y >> l, X >> l, snap + snap_off[l], (long long)snap_per_slot,
                           snap_pitch[l], true);
        // the snapshot is also the transform's input: rows and columns in one pass, no in-place hazard
        if (dwt_snap_supported(Y >> l, X >> l, snap_pitch[l], snap + snap_off[l], (long long)snap_per_slot))
          launch_dwt_snap(Lh, img, 0, nslots, snap + snap_off[l], (long long)snap_per_slot, snap_pitch[l], Y >> l,
                          X >> l);
        else
          launch_dwt_level(Lh, img, 0, nslots, Y >> l, X >> l, false);
      }
    } else {
      dwt_analyze_via_copy(Lh, img, 0, nslots, Y, X, L, tmpa, tmpa_stride, tmpa_pitch);
    }
    run_search(ME_INIT, desp(BY, L), desp(BX, L), 0, L == 0);
    for (int l = L - 1; l >= 0; --l) {
      if (snap && l == 0 && fuse0) {
        launch_load_u8(Lh, img, 0, m + 1, even, even_stride, 0, i0, 1, Y, X);
        launch_load_u8(Lh, img, m + 1, m, odd, odd_stride, 0, i0, 1, Y, X);
      } else if (snap)
        launch_region_copy(Lh, img, 0, nslots, Y >> l, X >> l, snap + snap_off[l], (long long)snap_per_slot,
                           snap_pitch[l], false);
      else
        dwt_synthesize1_via_copy(Lh, img, 0, nslots, desp(Y, l), desp(X, l), tmpa, tmpa_stride, tmpa_pitch);
      run_search(ME_DESCEND, desp(BY, l), desp(BX, l), sr, l == 0);
    }
    // byte planes of the level-0 interiors and their zero-high-band interpolations
    if (pr) {
      // the descent restored the pictures exactly: V_0 is the frames' luma (made above), every tile holds bytes
    } else {
      launch_plane_to_u8(Lh, img, 0, nslots, Y, X, v[0], (long long)vbytes[0], pitch[0], d_flags, tiles_x,
                         tiles_per_slot);
    }
    if (a == 2) {  // V_1 and V_2 from one read of V_0
      uint8_t *const outs[3] = {v[1], v[2], nullptr};
      const int ps[3] = {pitch[1], pitch[2], 0};
      const long long st[3] = {(long long)vbytes[1], (long long)vbytes[2], 0};
      launch_upsample_chain(Lh, v[0], Y, X, pitch[0], (long long)vbytes[0], 2, outs, ps, st, nslots);
    } else {
      for (int l = 1; l <= a; l++)
        launch_upsample2x(Lh, v[l - 1], Y << (l - 1), X << (l - 1), pitch[l - 1], (long long)vbytes[l - 1],
                          v[l], pitch[l], (long long)vbytes[l], nslots);
    }
    for (int l = 1; l <= a; l++) {
      CU(cudaMemsetAsync(d_slow, 0, sizeof(int), c->stream));
      SubpelParams q;
      q.b0 = img;
      q.slots = d_slots;
      q.v = v[l];
      q.v_slot_stride = (long long)vbytes[l];
      q.v_pitch = pitch[l];
      q.tile_bad = d_flags;
      q.tiles_x = tiles_x;
      q.tiles_per_slot = tiles_per_slot;
      q.mv_out = bufs[(j + n_search - 1) & 1];
      q.mv_in = bufs[(j + n_search) & 1];
      q.BY = BY;
      q.BX = BX;
      q.l = l;
      q.Y = Y;
      q.X = X;
      q.B = B;
      q.Bc = Bc;
      q.Ya = Y << a;
      q.Ba = B << a;
      q.size_field = (unsigned long long)heap_row_shorts(X << a, B << a) * 2ull | 1ull;
      q.lim = sr << a;
      q.slow_count = d_slow;
      q.slow_list = d_slow + 1;
      CU(cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
      q.bad_count = d_bad;
      q.bad_list = d_bad + 1;
      q.clean = cleanl[l];
      q.strip_top = stop[l];
      q.strip_left = sleft[l];
      q.strip_top_stride = (long long)top_sz[l];
      q.strip_left_stride = (long long)left_sz[l];
      // only the reference roles are border-filled: slots [0, m] and the carried copies
      launch_strips(Lh, q, l, 0, m + 1, stop[l], sleft[l], stop[1], sleft[1], (long long)top_sz[1],
                    (long long)left_sz[1], cleanl[1], v[1], (long long)vbytes[1], pitch[1]);
      launch_strips(Lh, q, l, 2 * m + 1, n_copy, stop[l], sleft[l], stop[1], sleft[1], (long long)top_sz[1],
                    (long long)left_sz[1], cleanl[1], v[1], (long long)vbytes[1], pitch[1]);
      q.v_rows_per_slot = (int)(vbytes[l] / pitch[l]);
      q.use_tma = (c->tma_mode != 0 &&
                   subpel_make_tensor_maps(v[l], pitch[l], (long long)q.v_rows_per_slot * nslots, bs << l,
                                           (bs << l) + 32, q.tm_p, q.tm_r))
                      ? 1
                      : 0;
      q.check_tiles = pr ? 0 : 1;
      {
        static const int dbg = getenv("QSVC_SUBPEL_DEBUG") ? atoi(getenv("QSVC_SUBPEL_DEBUG")) : 0;
        q.debug = dbg;
      }
      launch_subpel(Lh, q, bs << l, m);
      j++;
    }
    CU(cudaGetLastError());
    c->sad_ops += me_sad_ops(BY, BX, bs, 0, L, a) * m;
  }
  return QSVC_OK;
}

// even frame k at even + k*even_stride, odd frame i at odd + i*odd_stride (device).
static int me_level(qsvc_ctx *c, const uint8_t *even, long long even_stride, const uint8_t *odd,
                    long long odd_stride, int n_pairs, int X, int Y, int bs, int bd, int sr, int a,
                    int first_global, short *mv_out) {
  TRY(check_geometry(X, Y, bs, a));
  if (sr < 1 || bd < 0) return fail(QSVC_EINVAL, "bad search_range/border_size");
  const int BY = Y / bs, BX = X / bs;
  if (n_pairs <= 0 || BY == 0 || BX == 0) return QSVC_OK;
  int L = me_levels(sr);
  if (L < 0) L = 0;
  if (L >= 1 && ((Y >> (L - 1)) < 2 || (X >> (L - 1)) < 2))
    return fail(QSVC_EINVAL, "search_range %d needs %d pyramid levels: picture too small", sr, L);
  const int B = sr + bd;
  const int Ya = Y << a, Xa = X << a, Ba = B << a;
  // pyramid descent is an exact inverse of the analysis iff floor- and
  // ceil-halving agree on every level used
  bool pr = true;
  for (int l = 0; l < L; l++)
    if ((Y >> l) != desp(Y, l) || (X >> l) != desp(X, l)) pr = false;

  if (c->me_mode != 1 && me_fused_ok(X, Y, bs, bd, sr, a))
    return me_level_fused(c, even, even_stride, odd, odd_stride, n_pairs, X, Y, bs, sr, a, L, pr,
                          first_global, mv_out);
  if (c->me_mode == 2) return fail(QSVC_EINVAL, "fused ME path requested but not applicable");

  const long long field = 4LL * BY * BX;
  const int n_search = 1 + L + a;
  Launch Lh = c->L();

  // chunk the level so that the image planes fit the budget
  int S = heap_row_shorts(Xa, Ba);
  size_t slot_bytes = (size_t)((long long)Ya + 2LL * Ba + 2) * S * sizeof(short);
  int per_pair_slots = pr ? 2 : 3;
  if (pr) slot_bytes += (size_t)Y * (((X + 7) & ~7)) * sizeof(short) * 4 / 3 + 64;  // LL snapshots of the descent
  long long max_pairs = (long long)(c->me_budget / slot_bytes - 1) / per_pair_slots;
  if (max_pairs < 1) max_pairs = 1;

  for (int i0 = 0; i0 < n_pairs; i0 += (int)max_pairs) {
    const int m = (int)std::min<long long>(max_pairs, n_pairs - i0);
    Scratch s(c);
    const int n_copy = pr ? 0 : m - 1;  // separate R0-role buffers for frames i0+1 .. i0+m-1
    const int nslots = 2 * m + 1 + n_copy;
    PlaneAlloc pa;
    TRY(alloc_heap_planes(s, nslots, Ya, Xa, Ba, &pa));
    Plane img = pa.p;
    short *mv_tmp;
    int *d_slots;
    TRY(s.get((size_t)m * field * sizeof(short), (void **)&mv_tmp));
    TRY(s.get((size_t)m * 3 * sizeof(int), (void **)&d_slots));
    std::vector<int> slots(3 * m);
    for (int i = 0; i < m; i++) {
      slots[3 * i] = (!pr && i >= 1) ? 2 * m + i : i;
      slots[3 * i + 1] = i + 1;
      slots[3 * i + 2] = m + 1 + i;
    }
    CU(cudaMemcpyAsync(d_slots, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(pa.raw, 0, pa.bytes, c->stream));
    launch_size_fields(Lh, img, 0, nslots, Ya + 2 * Ba);
    // fresh(even_k): zero interior, luma in the top-left Y x X, fill_border with the
    // UNSHIFTED sizes (motion_estimate.cpp:803-820)
    launch_load_u8(Lh, img, 0, m + 1, even, even_stride, 0, i0, 1, Y, X);
    launch_load_u8(Lh, img, m + 1, m, odd, odd_stride, 0, i0, 1, Y, X);
    launch_fill_border(Lh, img, 0, m + 1, Y, X, B);
    if (n_copy > 0) {
      launch_load_u8(Lh, img, 2 * m + 1, n_copy, even, even_stride, 0, i0 + 1, 1, Y, X);
      launch_fill_border(Lh, img, 2 * m + 1, n_copy, Y, X, B);
    }
    if (!pr) {
      // reference[0] of every pair but the process's first is the buffer the
      // previous pair left behind: one full pass of the (non-invertible) pyramid.
      auto used_state = [&](int slot0, int n) {
        if (n <= 0) return;
        dwt_analyze(Lh, img, slot0, n, Y, X, L);
        for (int l = L - 1; l >= 0; --l) dwt_synthesize(Lh, img, slot0, n, desp(Y, l), desp(X, l), 1);
        for (int l = 1; l <= a; l++) dwt_synthesize(Lh, img, slot0, n, Y << l, X << l, 1);
        dwt_analyze(Lh, img, slot0, n, Ya, Xa, a);
      };
      if (!(first_global && i0 == 0)) used_state(0, 1);
      used_state(2 * m + 1, n_copy);
    }

    short *bufs[2] = {mv_out + (long long)i0 * field, mv_tmp};
    int j = 0;  // search index; search j writes bufs[(j + n_search - 1) & 1] so the last lands in mv_out
    auto run_search = [&](int mode, int nby, int nbx, int bsl, int bdl, int lim) {
      SearchParams q;
      q.img = img;
      q.slots = d_slots;
      q.mv_out = bufs[(j + n_search - 1) & 1];
      q.mv_in = bufs[(j + n_search) & 1];
      q.BY = BY;
      q.BX = BX;
      q.nby = nby;
      q.nbx = nbx;
      q.bs = bsl;
      q.bd = bdl;
      q.mode = mode;
      q.lim = lim;
      launch_search(Lh, q, m);
      j++;
    };
    // For an invertible pyramid the descent only restores what the analysis overwrote:
    // snapshot each LL region before it is transformed and copy it back instead of
    // running the inverse transform (identical result, half the passes).
    short *snap = nullptr;
    std::vector<size_t> snap_off(L + 1, 0);
    std::vector<int> snap_pitch(L + 1, 0);
    size_t snap_per_slot = 0;
    if (pr && L > 0) {
      for (int l = 0; l < L; l++) {
        snap_off[l] = snap_per_slot;
        snap_pitch[l] = ((X >> l) + 7) & ~7;
        snap_per_slot += (size_t)(Y >> l) * snap_pitch[l];
      }
      if (s.get(snap_per_slot * nslots * sizeof(short), (void **)&snap) != QSVC_OK) snap = nullptr;  // fall back
    }
    if (snap) {
      for (int l = 0; l < L; l++) {
        launch_region_copy(Lh, img, 0, nslots, Y >> l, X >> l, snap + snap_off[l], (long long)snap_per_slot,
                           snap_pitch[l], true);
        // the snapshot is also the transform's input: rows and columns in one pass, no in-place hazard
        if (dwt_snap_supported(Y >> l, X >> l, snap_pitch[l], snap + snap_off[l], (long long)snap_per_slot))
          launch_dwt_snap(Lh, img, 0, nslots, snap + snap_off[l], (long long)snap_per_slot, snap_pitch[l], Y >> l,
                          X >> l);
        else
          launch_dwt_level(Lh, img, 0, nslots, Y >> l, X >> l, false);
      }
    } else {
      dwt_analyze(Lh, img, 0, nslots, Y, X, L);
    }
    run_search(ME_INIT, desp(BY, L), desp(BX, L), bs, bd, 0);
    for (int l = L - 1; l >= 0; --l) {
      if (snap)
        launch_region_copy(Lh, img, 0, nslots, Y >> l, X >> l, snap + snap_off[l], (long long)snap_per_slot,
                           snap_pitch[l], false);
      else
        dwt_synthesize(Lh, img, 0, nslots, desp(Y, l), desp(X, l), 1);
      run_search(ME_DESCEND, desp(BY, l), desp(BX, l), bs, bd, sr);
    }
    for (int l = 1; l <= a; l++) {
      dwt_synthesize(Lh, img, 0, nslots, Y << l, X << l, 1);
      run_search(ME_SUBPEL, BY, BX, bs << l, bd >> l, sr << a);
    }
    CU(cudaGetLastError());
    c->sad_ops += me_sad_ops(BY, BX, bs, bd, L, a) * m;
    // the scratch planes are reused by later work on the same stream only
  }
  return QSVC_OK;
}

// ------------------------------------------------- decorrelate / correlate

// entropy.cpp:20-34 with the reference's expression shape: float division,
// logf on the float probability, quotient in double, float accumulate.  Runs on
// the host so that it goes through the same libm as the reference binary.
static float ref_entropy(const int *count, int n) {
  float entropy = 0.0f;
  int total = 0;
  for (int i = 0; i < n; i++) total += count[i];
  for (int i = 0; i < n; i++)
    if (count[i]) {
      float prob = (float)count[i] / total;
      entropy += prob * (float)(logf(prob) / log(2.0));
    }
  return -entropy;
}

// Up-sampled reference planes of one even frame (decorrelate.cpp:583-686):
// three components at luma size << a by zero-high-band synthesis.
static void prepare_reference(qsvc_ctx *c, Plane ref, int slot_set, const uint8_t *frame, int X,
                              int Y, int a) {
  Launch Lh = c->L();
  const int s0 = slot_set * 3;
  cudaMemsetAsync(ref.base + (long long)s0 * ref.slot_stride, 0,
                  (size_t)3 * ref.slot_stride * sizeof(short), c->stream);
  launch_load_u8(Lh, ref, s0, 1, frame, 0, 0, 0, 0, Y, X);
  launch_load_u8(Lh, ref, s0 + 1, 1, frame, 0, comp_offset(X, Y, 1), 0, 0, Y / 2, X / 2);
  launch_load_u8(Lh, ref, s0 + 2, 1, frame, 0, comp_offset(X, Y, 2), 0, 0, Y / 2, X / 2);
  dwt_synthesize(Lh, ref, s0 + 1, 2, Y, X, 1);
  for (int s = 1; s <= a; s++) dwt_synthesize(Lh, ref, s0, 3, Y << s, X << s, 1);
}

// The motion field of the level may still be in the making on the other lane (analyze_levels): whoever
// reads it first waits for it here.  The byte-plane path up-samples its reference planes before.
static int await_motion(qsvc_ctx *c) {
  if (c->mv_ready) {
    cudaEvent_t e = c->mv_ready;
    c->mv_ready = nullptr;
    CU(cudaStreamWaitEvent(c->stream, e, 0));
  }
  return QSVC_OK;
}

#include "mc_fused.inc"

// analysis != 0: decorrelate (in = odd frames, out = high frames);
// analysis == 0: correlate   (in = high frames, out = odd frames, types given).
static int mc_level(qsvc_ctx *c, int analysis, const uint8_t *even, long long even_stride,
                    const uint8_t *in, long long in_stride, const short *mv_in, int n_pairs, int X,
                    int Y, int bs, int ov, int sr, int a, int always_B, const char *types_in,
                    uint8_t *out, long long out_stride, std::string *types_out, short *mv_out,
                    uint8_t *prediction_out) {
  TRY(check_geometry(X, Y, bs, a));
  if (ov < 0) return fail(QSVC_EINVAL, "bad block_overlaping");
  int ov_levels = 0;
  if (ov > 0) {
    // decorrelate.cpp:84-88: levels of the per-block transform.  Areas no block covers keep the
    // previous pair's transformed leftovers and then go through the picture synthesis (A.2.6): the
    // literal path below replays exactly that (one prediction buffer per call, pairs in order), so
    // ragged pictures are exact within one call; a later GOP shard would need the whole buffer of
    // its left neighbour, which no exchange carries.
    if ((X % bs != 0 || Y % bs != 0) && c->tail_fn)
      return fail(QSVC_EINVAL, "block_overlaping > 0 on a picture that is not a multiple of the block size cannot be "
                               "GOP-sharded (the whole prediction buffer is carried from pair to pair)");
    if (!predict_obmc_supported(bs << a, ov << a))
      return fail(QSVC_EINVAL, "block_overlaping=%d: extended block does not fit shared memory", ov);
    ov_levels = (int)rint(log((double)(ov << a)) / log(2.0));
    if (ov_levels < 0) ov_levels = 0;
    if (((bs << a) >> ov_levels) == 0)
      return fail(QSVC_EINVAL, "block_overlaping=%d: more transform levels than the block has", ov);
  }
  const int BY = Y / bs, BX = X / bs;
  const long long field = 4LL * BY * BX;
  const long long fb = frame_bytes(X, Y);
  const int Ya = Y << a, Xa = X << a, bsa = bs << a;
  const int ba = (4 * sr + ov) << a;
  Launch Lh = c->L();
  Scratch s(c);
  const bool fused = c->mc_mode != 1 && mc_fused_ok(X, Y, bs, ov, a);
  if (c->mc_mode == 2 && !fused) return fail(QSVC_EINVAL, "fused MC path requested but not applicable");
  if (!fused) TRY(await_motion(c));
  if (!fused && analysis && c->upload_gops > 0 && c->cur_level == 1)
    CU(cudaStreamWaitEvent(c->stream, c->upload_events[c->upload_gops - 1], 0));  // qsvc_analyze: the clip has landed
  int *d_hist = nullptr;
  const int HS = 1024;  // per pair: 256 predicted, 256 residue, 257 motion (+pad)
  const bool need_hist = analysis && !always_B;
  if (need_hist) {
    TRY(s.get((size_t)n_pairs * HS * sizeof(int), (void **)&d_hist));
    CU(cudaMemsetAsync(d_hist, 0, (size_t)n_pairs * HS * sizeof(int), c->stream));
  }
  const int padh = heap_row_shorts(Xa, ba) - (Xa + 2 * ba);

  if (!fused && c->tail_fn && Y % bs != 0)
    return fail(QSVC_EINVAL, "the prediction tail exchange needs the byte-plane path (X %% 8 == 0, X %% block_size == 0)");
  if (fused) {
    TRY(mc_fused_core(c, analysis, even, even_stride, in, in_stride, mv_in, n_pairs, X, Y, bs, sr, a,
                      types_in, out, out_stride, prediction_out, need_hist ? d_hist : nullptr, HS));
    if (need_hist)
      for (int i = 0; i < n_pairs; i++)
        launch_mv_hist(Lh, mv_in + (long long)i * field, (int)field, d_hist + (long long)i * HS + 512);
  } else {
  PlaneAlloc ref, pred;
  TRY(alloc_dense_planes(s, 6, Ya, Xa, &ref));
  TRY(alloc_dense_planes(s, 3, Ya, Xa, &pred));
  CU(cudaMemsetAsync(pred.raw, 0, pred.bytes, c->stream));
  if (n_pairs > 0) prepare_reference(c, ref.p, 0, even, X, Y, a);
  for (int i = 0; i < n_pairs; i++) {
    prepare_reference(c, ref.p, (i + 1) & 1, even + (long long)(i + 1) * even_stride, X, Y, a);
    const short *mv = mv_in + (long long)i * field;
    PredictParams q;
    q.ref = ref.p;
    q.pred = pred.p;
    q.mv = mv;
    q.r0_slot = i & 1;
    q.r1_slot = (i + 1) & 1;
    q.BY = BY;
    q.BX = BX;
    q.bsa = bsa;
    q.Ya = Ya;
    q.Xa = Xa;
    q.ba = ba;
    q.padh = padh;
    if (ov > 0) {
      launch_predict_obmc(Lh, q, ov << a, ov_levels);
      dwt_synthesize(Lh, pred.p, 0, 3, Ya, Xa, ov_levels);
      launch_clip_uncovered(Lh, pred.p, Ya, Xa, 0, 0);  // decorrelate.cpp:841-848 over the whole picture
    } else {
      launch_predict(Lh, q);
      launch_clip_uncovered(Lh, pred.p, Ya, Xa, BY * bsa, BX * bsa);
    }
    dwt_analyze(Lh, pred.p, 0, 3, Ya, Xa, a);
    dwt_analyze(Lh, pred.p, 1, 2, Y, X, 1);
    ResidueParams r;
    r.pred = pred.p;
    r.odd = in + (long long)i * in_stride;
    r.out = out + (long long)i * out_stride;
    r.prediction = prediction_out ? prediction_out + (long long)i * fb : nullptr;
    r.hist = need_hist ? d_hist + (long long)i * HS : nullptr;
    r.X = X;
    r.Y = Y;
    r.synth = analysis ? 0 : 1;
    r.is_I = (!analysis && types_in[i] == 'I') ? 1 : 0;
    launch_residue(Lh, r);
    if (need_hist) launch_mv_hist(Lh, mv, (int)field, d_hist + (long long)i * HS + 512);
  }
  }
  CU(cudaGetLastError());
  if (!analysis) return QSVC_OK;

  types_out->assign((size_t)n_pairs, 'B');
  int rc = QSVC_OK;
  if (need_hist) {
    std::vector<int> hist((size_t)n_pairs * HS);
    CU(cudaMemcpyAsync(hist.data(), d_hist, hist.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n_pairs; i++) {
      const int *h = hist.data() + (size_t)i * HS;
      if (h[512 + 256] != 0) rc = 1;  // motion component outside the reference's histogram
      // decorrelate.cpp:934-979
      float predicted_entropy = ref_entropy(h, 256);
      float residue_entropy = ref_entropy(h + 256, 256);
      float motion_entropy = ref_entropy(h + 512, 256);
      int predicted_size = (int)(predicted_entropy * (float)Y * (float)X);
      int residue_size = (int)(residue_entropy * (float)Y * (float)X);
      int motion_size = (int)(motion_entropy * (float)BY * (float)BX);
      if (predicted_size <= (residue_size + motion_size)) (*types_out)[i] = 'I';
    }
  }
  // 'I': high = raw odd frame, motion_out = 0; 'B': motion_out = motion_in
  for (int i = 0; i < n_pairs;) {  // runs of equal frame types
    int j = i;
    while (j < n_pairs && (*types_out)[j] == (*types_out)[i]) j++;
    if ((*types_out)[i] == 'I') {
      launch_copy_strided(Lh, out + (long long)i * out_stride, out_stride, in + (long long)i * in_stride, in_stride,
                          (size_t)fb, j - i);
      if (mv_out) CU(cudaMemsetAsync(mv_out + (long long)i * field, 0, (size_t)(j - i) * field * sizeof(short), c->stream));
    } else if (mv_out && mv_out != mv_in) {
      launch_copy_bytes(Lh, mv_out + (long long)i * field, mv_in + (long long)i * field,
                        (size_t)(j - i) * field * sizeof(short));
    }
    i = j;
  }
  CU(cudaGetLastError());
  if (rc == 1)
    return fail(QSVC_EDOMAIN,
                "motion component outside [-128,127] with always_B=0: the reference indexes its "
                "256-bin histogram out of bounds here (decorrelate.cpp:809-812); use --always_B=1");
  return QSVC_OK;
}

// ------------------------------------------------------- update / un_update

static int update_level(qsvc_ctx *c, int inverse, const uint8_t *in, long long in_stride,
                        const uint8_t *high, long long high_stride, const short *mv,
                        const char *types, int n_pairs, int X, int Y, int bs, float uf, uint8_t *out,
                        long long out_stride) {
  TRY(check_geometry(X, Y, bs, 0));
  const int BY = Y / bs, BX = X / bs;
  const long long field = 4LL * BY * BX;
  const long long fb = frame_bytes(X, Y);
  Launch Lh = c->L();
  Scratch s(c);
  PlaneAlloc ref, res;
  TRY(alloc_dense_planes(s, 3, Y, X, &ref));
  TRY(alloc_dense_planes(s, 3, Y, X, &res));
  // residue[1|2] is only ever written in its top-left quarter (update.cpp:512-520,
  // UPDATE_STEP undefined); the rest stays at the allocator's zeros.
  CU(cudaMemsetAsync(res.raw, 0, res.bytes, c->stream));
  int *d_reach;
  TRY(s.get(256, (void **)&d_reach));
  if (BY == 0 || BX == 0 || uf == 0.0f) {
    // update_factor == 0 (analyze.py's default): every contribution is aux + (+-0) with aux
    // already in [0,255], and the chroma up/down pair is exact: the tool is a copy (A.4)
    launch_copy_strided(Lh, out, out_stride, in, in_stride, (size_t)fb, n_pairs + 1);
    CU(cudaGetLastError());
    return QSVC_OK;
  }
  if (c->boundary_fn == nullptr || n_pairs == 0) {
    // Frames are independent of each other (frame k takes pair k-1's NEXT update, then pair k's PREV
    // update): every frame of the level goes through the same handful of launches, chunked by HBM.
    const int tiles_x = (X + 15) / 16, tiles_y = (Y + 15) / 16, ntiles = tiles_x * tiles_y, CAP = 32;
    char *d_types;
    int *d_cnt, *d_list, *d_reach;
    TRY(s.get((size_t)n_pairs + 16, (void **)&d_types));
    TRY(s.get((size_t)2 * n_pairs * ntiles * sizeof(int), (void **)&d_cnt));
    TRY(s.get((size_t)2 * n_pairs * ntiles * CAP * sizeof(int), (void **)&d_list));
    int4 *d_geo = nullptr;  // the dyadic kernel reads the listed blocks' geometry instead of their vectors
    if (update_is_dyadic(uf)) TRY(s.get((size_t)2 * n_pairs * ntiles * CAP * sizeof(int4), (void **)&d_geo));
    TRY(s.get((size_t)2 * n_pairs * sizeof(int) + 16, (void **)&d_reach));
    CU(cudaMemcpyAsync(d_types, types, (size_t)n_pairs, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(d_cnt, 0, (size_t)2 * n_pairs * ntiles * sizeof(int), c->stream));
    CU(cudaMemsetAsync(d_reach, 0, (size_t)2 * n_pairs * sizeof(int), c->stream));
    UpdateBatchParams q;
    q.high = high;
    q.high_stride = high_stride;
    q.mv = mv;
    q.types = d_types;
    q.cnt = d_cnt;
    q.list = d_list;
    q.geo = d_geo;
    q.reach = d_reach;
    q.cap = CAP;
    q.n_pairs = n_pairs;
    q.BY = BY;
    q.BX = BX;
    q.bs = bs;
    q.Y = Y;
    q.X = X;
    q.tiles_x = tiles_x;
    q.tiles_y = tiles_y;
    q.uf = uf;
    q.inverse = inverse;
    q.slots_per_comp = 0;
    q.frame0 = 0;
    q.ref = ref.p;
    launch_update_bin(Lh, q);
    const size_t per_frame = (size_t)3 * Y * ((X + 7) & ~7) * sizeof(short);
    int max_frames = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_pairs + 1, (c->me_budget / 4) / per_frame));
    // dyadic factor, 4:2:0 geometry with word-aligned rows: luma is updated on the frames' own bytes, the chroma
    // components go to luma-sized planes and back in one pass each (no generic transform passes, no luma planes)
    const bool lean = update_is_dyadic(uf) && CAP <= 32 && X % 8 == 0 && Y % 2 == 0;
    for (int k0 = 0; lean && k0 <= n_pairs; k0 += max_frames) {
      const int m = std::min(max_frames, n_pairs + 1 - k0);
      Scratch sc(c);
      PlaneAlloc planes;
      TRY(alloc_dense_planes(sc, 2 * m, Y, X, &planes));
      launch_chroma_up_s16(Lh, planes.p, 0, m, in, in_stride, comp_offset(X, Y, 1), k0, Y, X);
      launch_chroma_up_s16(Lh, planes.p, m, m, in, in_stride, comp_offset(X, Y, 2), k0, Y, X);
      q.ref = planes.p;
      q.ref.base -= (long long)m * planes.p.slot_stride;  // component c >= 1 of frame f: slot (c - 1) * m + f
      q.slots_per_comp = m;
      q.frame0 = k0;
      q.luma_in = in;
      q.luma_in_stride = in_stride;
      q.luma_out = out;
      q.luma_out_stride = out_stride;
      launch_update_batch(Lh, q, m);
      launch_ll1_store_u8(Lh, planes.p, 0, m, out, out_stride, comp_offset(X, Y, 1), k0, Y, X);
      launch_ll1_store_u8(Lh, planes.p, m, m, out, out_stride, comp_offset(X, Y, 2), k0, Y, X);
    }
    for (int k0 = 0; !lean && k0 <= n_pairs; k0 += max_frames) {
      const int m = std::min(max_frames, n_pairs + 1 - k0);
      Scratch sc(c);
      PlaneAlloc planes;
      TRY(alloc_dense_planes(sc, 3 * m, Y, X, &planes));
      CU(cudaMemsetAsync(planes.raw, 0, planes.bytes, c->stream));
      launch_load_u8(Lh, planes.p, 0, m, in, in_stride, 0, k0, 1, Y, X);
      launch_load_u8(Lh, planes.p, m, m, in, in_stride, comp_offset(X, Y, 1), k0, 1, Y / 2, X / 2);
      launch_load_u8(Lh, planes.p, 2 * m, m, in, in_stride, comp_offset(X, Y, 2), k0, 1, Y / 2, X / 2);
      dwt_synthesize(Lh, planes.p, m, 2 * m, Y, X, 1);
      q.ref = planes.p;
      q.slots_per_comp = m;
      q.frame0 = k0;
      launch_update_batch(Lh, q, m);
      dwt_analyze(Lh, planes.p, m, 2 * m, Y, X, 1);
      launch_store_u8(Lh, planes.p, 0, m, out, out_stride, 0, k0, 1, Y, X);
      launch_store_u8(Lh, planes.p, m, m, out, out_stride, comp_offset(X, Y, 1), k0, 1, Y / 2, X / 2);
      launch_store_u8(Lh, planes.p, 2 * m, m, out, out_stride, comp_offset(X, Y, 2), k0, 1, Y / 2, X / 2);
    }
    CU(cudaGetLastError());
    return QSVC_OK;
  }
  // GOP shards: the frame shared with a neighbour receives the left
  // shard's NEXT update first and the right shard's PREV update second, and every contribution
  // is clamped and truncated, so the int16 planes travel from left to right between the two
  // passes and the finished frame travels back.  The last frame is processed first so that a
  // chain of shards does not serialise on this hand-over.
  std::vector<uint8_t> h_planes;
  const bool xch = c->boundary_fn != nullptr && n_pairs > 0;
  auto boundary = [&](int phase, void *data, long long bytes) -> int {
    CU(cudaStreamSynchronize(c->stream));
    const int r = c->boundary_fn(c->boundary_user, c->cur_level, inverse, phase, data, bytes);
    if (r < 0) return fail(QSVC_EINVAL, "boundary exchange callback failed (phase %d)", phase);
    return r;
  };
  bool last_from_right = false;
  for (int kk = 0; kk <= n_pairs; kk++) {
    const int k = xch ? (kk == 0 ? n_pairs : kk - 1) : kk;  // with an exchange: n, 0, 1, .., n-1
    const uint8_t *frame = in + (long long)k * in_stride;
    uint8_t *dst = out + (long long)k * out_stride;
    bool upd_next = k >= 1 && types[k - 1] == 'B';
    bool upd_prev = k < n_pairs && types[k] == 'B';
    const bool first_xch = xch && k == 0, last_xch = xch && k == n_pairs;
    // update_factor == 0: every contribution is aux + (+-0) with aux already in
    // [0,255], so the scatter is the identity (analyze.py's default)
    if (!first_xch && !last_xch && !(upd_next || upd_prev)) {
      // chroma up (zero-high synthesis) and down (analysis) are exact inverses
      launch_copy_bytes(Lh, dst, frame, (size_t)fb);
      continue;
    }
    CU(cudaMemsetAsync(ref.raw, 0, ref.bytes, c->stream));
    launch_load_u8(Lh, ref.p, 0, 1, frame, 0, 0, 0, 0, Y, X);
    launch_load_u8(Lh, ref.p, 1, 1, frame, 0, comp_offset(X, Y, 1), 0, 0, Y / 2, X / 2);
    launch_load_u8(Lh, ref.p, 2, 1, frame, 0, comp_offset(X, Y, 2), 0, 0, Y / 2, X / 2);
    dwt_synthesize(Lh, ref.p, 1, 2, Y, X, 1);
    if (first_xch) {
      // phase 0: the planes as the left neighbour left them after its NEXT pass (1: filled in)
      h_planes.resize(ref.bytes);
      const int got = boundary(0, h_planes.data(), (long long)ref.bytes);
      if (got < 0) return got;
      if (got > 0) {
        CU(cudaMemcpyAsync(ref.raw, h_planes.data(), ref.bytes, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
      }
    }
    for (int pass = 0; pass < 2; pass++) {
      // frame k first receives pair k-1's NEXT update, then pair k's PREV update
      int pair = pass == 0 ? k - 1 : k;
      if (pass == 0 ? !upd_next : !upd_prev) continue;
      const uint8_t *h = high + (long long)pair * high_stride;
      launch_load_residue(Lh, res.p, 0, h, Y, X);
      launch_load_residue(Lh, res.p, 1, h + comp_offset(X, Y, 1), Y / 2, X / 2);
      launch_load_residue(Lh, res.p, 2, h + comp_offset(X, Y, 2), Y / 2, X / 2);
      UpdateParams q;
      q.ref = ref.p;
      q.res = res.p;
      q.mv = mv + (long long)pair * field;
      q.dir = pass == 0 ? MV_NEXT_X : MV_PREV_X;
      launch_mv_reach(Lh, q.mv + (long long)q.dir * BY * BX, 2 * BY * BX, d_reach);
      q.reach = d_reach;
      q.BY = BY;
      q.BX = BX;
      q.bs = bs;
      q.Y = Y;
      q.X = X;
      q.uf = uf;
      q.inverse = inverse;
      launch_update(Lh, q);
    }
    if (last_xch) {
      // phase 1: the planes after this shard's NEXT pass, for the right neighbour (1: one exists
      // and will send the finished frame back in phase 3)
      h_planes.resize(ref.bytes);
      CU(cudaMemcpyAsync(h_planes.data(), ref.raw, ref.bytes, cudaMemcpyDeviceToHost, c->stream));
      const int has_right = boundary(1, h_planes.data(), (long long)ref.bytes);
      if (has_right < 0) return has_right;
      last_from_right = has_right > 0;
    }
    dwt_analyze(Lh, ref.p, 1, 2, Y, X, 1);
    launch_store_u8(Lh, ref.p, 0, 1, dst, 0, 0, 0, 0, Y, X);
    launch_store_u8(Lh, ref.p, 1, 1, dst, 0, comp_offset(X, Y, 1), 0, 0, Y / 2, X / 2);
    launch_store_u8(Lh, ref.p, 2, 1, dst, 0, comp_offset(X, Y, 2), 0, 0, Y / 2, X / 2);
    if (first_xch) {
      // phase 2: this shard's finished first frame, for the left neighbour (ignored by the first shard)
      std::vector<uint8_t> f((size_t)fb);
      CU(cudaMemcpyAsync(f.data(), dst, (size_t)fb, cudaMemcpyDeviceToHost, c->stream));
      const int r = boundary(2, f.data(), fb);
      if (r < 0) return r;
    }
  }
  if (last_from_right) {
    // phase 3: the finished shared frame from the right neighbour replaces this shard's last frame
    std::vector<uint8_t> f((size_t)fb);
    const int got = boundary(3, f.data(), fb);
    if (got < 0) return got;
    if (got > 0) {
      CU(cudaMemcpyAsync(out + (long long)n_pairs * out_stride, f.data(), (size_t)fb, cudaMemcpyHostToDevice, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
  }
  CU(cudaGetLastError());
  return QSVC_OK;
}

// ------------------------------------------------------------------- C ABI

extern "C" {

int qsvc_version(void) { return 1; }

int qsvc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char *qsvc_last_error(void) { return g_err.c_str(); }

qsvc_ctx *qsvc_create(int device) {
  int n = qsvc_device_count();
  if (n <= 0) {
    fail(QSVC_ECUDA, "no CUDA device: libqsvc_b200 has no CPU fallback");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    fail(QSVC_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    fail(QSVC_ECUDA, "cudaSetDevice(%d) failed", device);
    return nullptr;
  }
  qsvc_ctx *c = new qsvc_ctx();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->me_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
      cudaEventCreate(&c->ev2) != cudaSuccess || cudaEventCreate(&c->ev3) != cudaSuccess) {
    fail(QSVC_ECUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  int e = dwt_init_attributes();
  if (e != 0) {
    fail(QSVC_ECUDA, "kernel attribute setup failed: %s (is this an sm_100 device?)",
         cudaGetErrorString((cudaError_t)e));
    delete c;
    return nullptr;
  }
  if (const char *e = getenv("QSVC_ME_MODE")) c->me_mode = atoi(e);
  if (const char *e = getenv("QSVC_MC_MODE")) c->mc_mode = atoi(e);
  if (const char *e = getenv("QSVC_TMA")) c->tma_mode = atoi(e);
  if (const char *e = getenv("QSVC_MC_KERNEL")) c->mc_kernel = atoi(e);
  if (const char *e = getenv("QSVC_MC_RING")) c->mc_ring = atoi(e);
  if (const char *e = getenv("QSVC_ME_FUSE0")) c->me_fuse0 = atoi(e);
  if (const char *e = getenv("QSVC_ME_BYTES0")) c->me_bytes0 = atoi(e);
  if (const char *e = getenv("QSVC_OVERLAP")) c->overlap = atoi(e);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) c->me_budget = std::min<size_t>((size_t)64 << 30, free_b / 3);
  return c;
}

static void free_levels(qsvc_ctx *c) {
  for (auto &lv : c->levels) {
    pool_free(c, lv.high);
    pool_free(c, lv.motion);
    pool_free(c, lv.motion_filtered);
    pool_free(c, lv.low);
  }
  c->levels.clear();
}

void qsvc_destroy(qsvc_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto &b : c->pool)
    if (b.ptr) cudaFree(b.ptr);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaEventDestroy(c->ev2);
  cudaEventDestroy(c->ev3);
  for (auto e : c->level_events) cudaEventDestroy(e);
  for (auto e : c->upload_events) cudaEventDestroy(e);
  for (auto e : c->me_events) cudaEventDestroy(e);
  if (c->me_stream) cudaStreamDestroy(c->me_stream);
  cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->stream);
  delete c;
}

long long qsvc_launch_count(const qsvc_ctx *c) { return c ? c->launches : 0; }

int qsvc_timer_start(qsvc_ctx *c) {
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->ev0, c->stream));
  return QSVC_OK;
}
int qsvc_timer_stop(qsvc_ctx *c, float *ms) {
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaEventSynchronize(c->ev1));
  CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return QSVC_OK;
}
// Measured issue rate of the SAD instructions: SAD operations per second with
// packed bytes (__vsadu4) and with 32-bit lanes (__sad), best of 5.
int qsvc_int_peak(qsvc_ctx *c, double *u8_sad_ops_per_s, double *i32_sad_ops_per_s) {
  if (!c) return fail(QSVC_EINVAL, "null context");
  CU(cudaSetDevice(c->device));
  const int blocks = 148 * 8, iters = 4096;
  Scratch s(c);
  unsigned *d_out;
  TRY(s.get((size_t)blocks * 256 * sizeof(unsigned), (void **)&d_out));
  for (int packed = 1; packed >= 0; packed--) {
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
      CU(cudaEventRecord(c->ev0, c->stream));
      int e = run_int_peak(c->stream, d_out, blocks, iters, packed != 0);
      if (e) return fail(QSVC_ECUDA, "int peak launch: %s", cudaGetErrorString((cudaError_t)e));
      c->launches++;
      CU(cudaEventRecord(c->ev1, c->stream));
      CU(cudaEventSynchronize(c->ev1));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
      double ops = (double)blocks * 256 * iters * 8 * (packed ? 4 : 1);
      if (rep > 0 && ms > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    if (packed && u8_sad_ops_per_s) *u8_sad_ops_per_s = best;
    if (!packed && i32_sad_ops_per_s) *i32_sad_ops_per_s = best;
  }
  return QSVC_OK;
}
int qsvc_debug_tma_timeouts(void) { return subpel_tma_timeouts(); }
int qsvc_set_mc_mode(qsvc_ctx *c, int mode) {
  if (!c || mode < 0 || mode > 2) return fail(QSVC_EINVAL, "bad mc_mode");
  c->mc_mode = mode;
  return QSVC_OK;
}
int qsvc_set_tail_exchange(qsvc_ctx *c, qsvc_tail_fn fn, void *user) {
  if (!c) return fail(QSVC_EINVAL, "null context");
  c->tail_fn = fn;
  c->tail_user = user;
  c->tail_on_device = 0;
  return QSVC_OK;
}
int qsvc_set_tail_exchange_device(qsvc_ctx *c, qsvc_tail_fn fn, void *user) {
  if (!c) return fail(QSVC_EINVAL, "null context");
  c->tail_fn = fn;
  c->tail_user = user;
  c->tail_on_device = fn ? 1 : 0;
  return QSVC_OK;
}
int qsvc_host_register(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(QSVC_EINVAL, "bad arguments");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(QSVC_ECUDA, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e));
  }
  return QSVC_OK;
}
int qsvc_host_unregister(void *ptr) {
  if (!ptr) return QSVC_OK;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(QSVC_ECUDA, "cudaHostUnregister: %s", cudaGetErrorString(e));
  }
  return QSVC_OK;
}
int qsvc_set_boundary_exchange(qsvc_ctx *c, qsvc_boundary_fn fn, void *user) {
  if (!c) return fail(QSVC_EINVAL, "null context");
  c->boundary_fn = fn;
  c->boundary_user = user;
  return QSVC_OK;
}
int qsvc_set_overlap(qsvc_ctx *c, int on) {
  if (!c) return fail(QSVC_EINVAL, "null context");
  c->overlap = on ? 1 : 0;
  return QSVC_OK;
}
int qsvc_set_me_mode(qsvc_ctx *c, int mode) {
  if (!c || mode < 0 || mode > 2) return fail(QSVC_EINVAL, "bad me_mode");
  c->me_mode = mode;
  return QSVC_OK;
}
int qsvc_profile_enable(qsvc_ctx *c, int on) {
  if (!c) return fail(QSVC_EINVAL, "null context");
  c->prof.enabled = on != 0;
  c->prof.n = 0;
  return QSVC_OK;
}
// Sums the per-launch event pairs recorded since the last call, per kernel class.
int qsvc_profile_read(qsvc_ctx *c, float *ms, long long *launches, int n_classes) {
  if (!c) return fail(QSVC_EINVAL, "null context");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < n_classes; k++) {
    if (ms) ms[k] = 0.f;
    if (launches) launches[k] = 0;
  }
  for (int i = 0; i < c->prof.n; i++) {
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, c->prof.recs[i].a, c->prof.recs[i].b));
    int k = c->prof.recs[i].cls;
    if (k < n_classes) {
      if (ms) ms[k] += t;
      if (launches) launches[k]++;
    }
  }
  c->prof.n = 0;
  return QSVC_OK;
}
int qsvc_synchronize(qsvc_ctx *c) {
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  return QSVC_OK;
}

#define ENTER(c)                                            \
  if (!(c)) return fail(QSVC_EINVAL, "null context");       \
  CU(cudaSetDevice((c)->device));

int qsvc_motion_estimate(qsvc_ctx *c, const uint8_t *even, const uint8_t *odd, int n_pairs, int X,
                         int Y, int bs, int bd, int sr, int a, int first_global, int16_t *mv_out) {
  ENTER(c);
  if (n_pairs < 0 || !even || !odd || !mv_out) return fail(QSVC_EINVAL, "bad arguments");
  TRY(check_geometry(X, Y, bs, a));
  if (n_pairs == 0) return QSVC_OK;
  const long long fb = frame_bytes(X, Y), field = 4LL * (Y / bs) * (X / bs);
  Scratch s(c);
  uint8_t *d_even, *d_odd;
  short *d_mv;
  TRY(s.get((size_t)fb * (n_pairs + 1), (void **)&d_even));
  TRY(s.get((size_t)fb * n_pairs, (void **)&d_odd));
  TRY(s.get((size_t)std::max<long long>(field, 1) * n_pairs * sizeof(short), (void **)&d_mv));
  CU(cudaMemcpyAsync(d_even, even, (size_t)fb * (n_pairs + 1), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_odd, odd, (size_t)fb * n_pairs, cudaMemcpyHostToDevice, c->stream));
  TRY(me_level(c, d_even, fb, d_odd, fb, n_pairs, X, Y, bs, bd, sr, a, first_global, d_mv));
  CU(cudaMemcpyAsync(mv_out, d_mv, (size_t)field * n_pairs * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QSVC_OK;
}

int qsvc_decorrelate(qsvc_ctx *c, const uint8_t *even, const uint8_t *odd, const int16_t *mv_in,
                     int n_pairs, int X, int Y, int bs, int ov, int sr, int a, int always_B,
                     uint8_t *high_out, char *types_out, int16_t *mv_out, uint8_t *prediction_out) {
  ENTER(c);
  if (n_pairs < 0 || !even || !odd || !mv_in || !high_out || !types_out || !mv_out)
    return fail(QSVC_EINVAL, "bad arguments");
  TRY(check_geometry(X, Y, bs, a));
  if (n_pairs == 0) return QSVC_OK;
  const long long fb = frame_bytes(X, Y), field = 4LL * (Y / bs) * (X / bs);
  Scratch s(c);
  uint8_t *d_even, *d_odd, *d_high, *d_pred = nullptr;
  short *d_mv, *d_mvo;
  TRY(s.get((size_t)fb * (n_pairs + 1), (void **)&d_even));
  TRY(s.get((size_t)fb * n_pairs, (void **)&d_odd));
  TRY(s.get((size_t)fb * n_pairs, (void **)&d_high));
  if (prediction_out) TRY(s.get((size_t)fb * n_pairs, (void **)&d_pred));
  TRY(s.get((size_t)std::max<long long>(field, 1) * n_pairs * sizeof(short), (void **)&d_mv));
  TRY(s.get((size_t)std::max<long long>(field, 1) * n_pairs * sizeof(short), (void **)&d_mvo));
  CU(cudaMemcpyAsync(d_even, even, (size_t)fb * (n_pairs + 1), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_odd, odd, (size_t)fb * n_pairs, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_mv, mv_in, (size_t)field * n_pairs * sizeof(short), cudaMemcpyHostToDevice, c->stream));
  std::string types;
  TRY(mc_level(c, 1, d_even, fb, d_odd, fb, d_mv, n_pairs, X, Y, bs, ov, sr, a, always_B, nullptr,
               d_high, fb, &types, d_mvo, d_pred));
  CU(cudaMemcpyAsync(high_out, d_high, (size_t)fb * n_pairs, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(mv_out, d_mvo, (size_t)field * n_pairs * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
  if (prediction_out)
    CU(cudaMemcpyAsync(prediction_out, d_pred, (size_t)fb * n_pairs, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  memcpy(types_out, types.data(), (size_t)n_pairs);
  return QSVC_OK;
}

int qsvc_correlate(qsvc_ctx *c, const uint8_t *even, const uint8_t *high, const int16_t *mv_in,
                   const char *types, int n_pairs, int X, int Y, int bs, int ov, int sr, int a,
                   uint8_t *odd_out, uint8_t *prediction_out) {
  ENTER(c);
  if (n_pairs < 0 || !even || !high || !mv_in || !types || !odd_out)
    return fail(QSVC_EINVAL, "bad arguments");
  TRY(check_geometry(X, Y, bs, a));
  if (n_pairs == 0) return QSVC_OK;
  const long long fb = frame_bytes(X, Y), field = 4LL * (Y / bs) * (X / bs);
  Scratch s(c);
  uint8_t *d_even, *d_high, *d_odd, *d_pred = nullptr;
  short *d_mv;
  TRY(s.get((size_t)fb * (n_pairs + 1), (void **)&d_even));
  TRY(s.get((size_t)fb * n_pairs, (void **)&d_high));
  TRY(s.get((size_t)fb * n_pairs, (void **)&d_odd));
  if (prediction_out) TRY(s.get((size_t)fb * n_pairs, (void **)&d_pred));
  TRY(s.get((size_t)std::max<long long>(field, 1) * n_pairs * sizeof(short), (void **)&d_mv));
  CU(cudaMemcpyAsync(d_even, even, (size_t)fb * (n_pairs + 1), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_high, high, (size_t)fb * n_pairs, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_mv, mv_in, (size_t)field * n_pairs * sizeof(short), cudaMemcpyHostToDevice, c->stream));
  TRY(mc_level(c, 0, d_even, fb, d_high, fb, d_mv, n_pairs, X, Y, bs, ov, sr, a, 1, types, d_odd,
               fb, nullptr, nullptr, d_pred));
  CU(cudaMemcpyAsync(odd_out, d_odd, (size_t)fb * n_pairs, cudaMemcpyDeviceToHost, c->stream));
  if (prediction_out)
    CU(cudaMemcpyAsync(prediction_out, d_pred, (size_t)fb * n_pairs, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QSVC_OK;
}

int qsvc_update(qsvc_ctx *c, int inverse, const uint8_t *in, const uint8_t *high,
                const int16_t *mv, const char *types, int n_pairs, int X, int Y, int bs, float uf,
                uint8_t *out) {
  ENTER(c);
  if (n_pairs < 0 || !in || !out || (n_pairs > 0 && (!high || !mv || !types)))
    return fail(QSVC_EINVAL, "bad arguments");
  TRY(check_geometry(X, Y, bs, 0));
  const long long fb = frame_bytes(X, Y), field = 4LL * (Y / bs) * (X / bs);
  Scratch s(c);
  uint8_t *d_in, *d_high, *d_out;
  short *d_mv;
  TRY(s.get((size_t)fb * (n_pairs + 1), (void **)&d_in));
  TRY(s.get((size_t)fb * (n_pairs + 1), (void **)&d_out));
  TRY(s.get((size_t)fb * std::max(n_pairs, 1), (void **)&d_high));
  TRY(s.get((size_t)std::max<long long>(field, 1) * std::max(n_pairs, 1) * sizeof(short), (void **)&d_mv));
  CU(cudaMemcpyAsync(d_in, in, (size_t)fb * (n_pairs + 1), cudaMemcpyHostToDevice, c->stream));
  if (n_pairs > 0) {
    CU(cudaMemcpyAsync(d_high, high, (size_t)fb * n_pairs, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_mv, mv, (size_t)field * n_pairs * sizeof(short), cudaMemcpyHostToDevice, c->stream));
  }
  TRY(update_level(c, inverse, d_in, fb, d_high, fb, d_mv, types, n_pairs, X, Y, bs, uf, d_out, fb));
  CU(cudaMemcpyAsync(out, d_out, (size_t)fb * (n_pairs + 1), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QSVC_OK;
}

int qsvc_bidirectional_motion_decorrelate(qsvc_ctx *c, int inverse, const int16_t *in, int n_fields, int by, int bx,
                                          int16_t *out) {
  ENTER(c);
  if (n_fields < 0 || by < 0 || bx < 0 || (n_fields > 0 && by > 0 && bx > 0 && (!in || !out)))
    return fail(QSVC_EINVAL, "bad arguments");
  const size_t bytes = (size_t)n_fields * 4 * by * bx * sizeof(short);
  if (!bytes) return QSVC_OK;
  Scratch s(c);
  short *d_in, *d_out;
  TRY(s.get(bytes, (void **)&d_in));
  TRY(s.get(bytes, (void **)&d_out));
  CU(cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, c->stream));
  launch_mv_bidirectional(c->L(), d_in, d_out, n_fields, by * bx, inverse);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QSVC_OK;
}

int qsvc_interlevel_motion_decorrelate(qsvc_ctx *c, int inverse, const int16_t *in, int n_fields,
                                       const int16_t *reference, int n_reference, int by, int bx, int16_t *out) {
  ENTER(c);
  if (n_fields < 0 || n_reference < 0 || by < 0 || bx < 0 || (n_reference > 0 && !reference) ||
      (n_fields > 0 && by > 0 && bx > 0 && (!in || !out)))
    return fail(QSVC_EINVAL, "bad arguments");
  const size_t fsz = (size_t)4 * by * bx * sizeof(short), bytes = fsz * n_fields;
  if (!bytes) return QSVC_OK;
  // only the reference fields the reader's loop would reach matter
  const int n_ref = std::min(n_reference, (n_fields + 1) / 2);
  Scratch s(c);
  short *d_in, *d_out, *d_ref = nullptr;
  TRY(s.get(bytes, (void **)&d_in));
  TRY(s.get(bytes, (void **)&d_out));
  if (n_ref > 0) {
    TRY(s.get(fsz * n_ref, (void **)&d_ref));
    CU(cudaMemcpyAsync(d_ref, reference, fsz * n_ref, cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, c->stream));
  launch_mv_interlevel(c->L(), d_in, d_ref, d_out, n_fields, n_ref, by * bx, inverse);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QSVC_OK;
}

// Distortion between two byte streams per block (psnr.py:78-90 calls the external `snr` with
// --block_size = bytes per picture): sum of squared differences per block, exact.
int qsvc_sse(qsvc_ctx *c, const uint8_t *a, const uint8_t *b, long long block_bytes, int n_blocks,
             unsigned long long *sse_out) {
  ENTER(c);
  if (!a || !b || !sse_out || block_bytes <= 0 || n_blocks < 0) return fail(QSVC_EINVAL, "bad arguments");
  if (n_blocks == 0) return QSVC_OK;
  Scratch s(c);
  uint8_t *d_a, *d_b;
  unsigned long long *d_out;
  // chunks of at most 1 GiB per stream keep any file size inside the pool
  const int per = (int)std::max<long long>(1, std::min<long long>(n_blocks, ((long long)1 << 30) / block_bytes));
  TRY(s.get((size_t)per * block_bytes, (void **)&d_a));
  TRY(s.get((size_t)per * block_bytes, (void **)&d_b));
  TRY(s.get((size_t)n_blocks * sizeof(unsigned long long), (void **)&d_out));
  CU(cudaMemsetAsync(d_out, 0, (size_t)n_blocks * sizeof(unsigned long long), c->stream));
  for (int k0 = 0; k0 < n_blocks; k0 += per) {
    const int m = std::min(per, n_blocks - k0);
    CU(cudaMemcpyAsync(d_a, a + (long long)k0 * block_bytes, (size_t)m * block_bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_b, b + (long long)k0 * block_bytes, (size_t)m * block_bytes, cudaMemcpyHostToDevice, c->stream));
    launch_sse_u8(c->L(), d_a, d_b, block_bytes, m, d_out + k0);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(sse_out, d_out, (size_t)n_blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QSVC_OK;
}

// ------------------------------------------------------ resident sequence

int qsvc_resident_load(qsvc_ctx *c, const uint8_t *low0, int n_frames, int X, int Y) {
  ENTER(c);
  if (!low0 || n_frames < 1) return fail(QSVC_EINVAL, "bad arguments");
  TRY(check_geometry(X, Y, 1, 0));
  free_levels(c);
  pool_free(c, c->low0);
  c->low0 = nullptr;
  const long long fb = frame_bytes(X, Y);
  TRY(pool_alloc(c, (size_t)fb * n_frames, (void **)&c->low0));
  CU(cudaMemcpyAsync(c->low0, low0, (size_t)fb * n_frames, cudaMemcpyHostToDevice, c->stream));
  c->n_frames = n_frames;
  c->X = X;
  c->Y = Y;
  return QSVC_OK;
}

static int analyze_levels(qsvc_ctx *c, const qsvc_analyze_params *p, const qsvc_level_out *outs);

int qsvc_resident_analyze(qsvc_ctx *c, const qsvc_analyze_params *p) {
  ENTER(c);
  return analyze_levels(c, p, nullptr);
}

void *qsvc_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    fail(QSVC_ENOMEM, "cudaHostAlloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
void qsvc_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// analyze.py equivalent in one call with host buffers: upload, all levels, and each
// level's results copied back on a second stream while the next level computes.
int qsvc_analyze(qsvc_ctx *c, const qsvc_analyze_params *p, const uint8_t *low0, int n_frames,
                 const qsvc_level_out *outs) {
  ENTER(c);
  if (!p || !low0 || !outs) return fail(QSVC_EINVAL, "bad arguments");
  const int G = p->TRLs >= 2 && p->TRLs < 31 ? 1 << (p->TRLs - 1) : 0;
  if (G > 0 && n_frames > G + 1 && (n_frames - 1) % G == 0) {
    // upload GOP by GOP on the copy stream; level 1's motion estimation of GOP g starts as soon
    // as its frames (up to the boundary frame (g+1)*G) have arrived
    TRY(check_geometry(p->pixels_in_x, p->pixels_in_y, 1, 0));
    free_levels(c);
    pool_free(c, c->low0);
    c->low0 = nullptr;
    const long long fb = frame_bytes(p->pixels_in_x, p->pixels_in_y);
    TRY(pool_alloc(c, (size_t)fb * n_frames, (void **)&c->low0));
    c->n_frames = n_frames;
    c->X = p->pixels_in_x;
    c->Y = p->pixels_in_y;
    const int gops = (n_frames - 1) / G, per_gop = G / 2;  // pairs of a GOP at level 1
    // upload segments = ranges of level-1 pairs: whole GOPs, except that the first GOP starts with a quarter of
    // itself, so that the first motion estimation waits for a quarter of a GOP's frames, not a whole one
    c->upload_segs.clear();
    const int head = per_gop >= 16 ? per_gop / 4 : 0;
    if (head > 0) c->upload_segs.push_back(head);
    for (int g = 0; g < gops; g++) c->upload_segs.push_back((g + 1) * per_gop);  // end pair (exclusive) of each segment
    const int nseg = (int)c->upload_segs.size();
    while ((int)c->upload_events.size() < nseg) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->upload_events.push_back(e);
    }
    // the pool hands out memory that earlier work on the compute stream may still be using
    CU(cudaEventRecord(c->ev3, c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev3, 0));
    long long f0 = 0;
    for (int sgm = 0; sgm < nseg; sgm++) {
      const long long f1 = 2LL * c->upload_segs[sgm];  // last frame (inclusive) the segment's pairs read
      CU(cudaMemcpyAsync(c->low0 + f0 * fb, low0 + f0 * fb, (size_t)((f1 - f0 + 1) * fb), cudaMemcpyHostToDevice,
                         c->copy_stream));
      CU(cudaEventRecord(c->upload_events[sgm], c->copy_stream));
      f0 = f1 + 1;
    }
    c->upload_gops = nseg;
    c->upload_gop_frames = G;
  } else {
    TRY(qsvc_resident_load(c, low0, n_frames, p->pixels_in_x, p->pixels_in_y));
  }
  const int rc = analyze_levels(c, p, outs);
  c->upload_gops = 0;
  if (rc != QSVC_OK) {
    cudaStreamSynchronize(c->copy_stream);
    return rc;
  }
  CU(cudaStreamSynchronize(c->copy_stream));
  for (int t = 1; t < p->TRLs; t++)
    if (outs[t].frame_types) memcpy(outs[t].frame_types, c->levels[t].types.data(), (size_t)c->levels[t].n_pairs);
  return QSVC_OK;
}

// Runs `body` with the context switched to the motion-estimation lane (its own stream and its own
// share of the memory pool), then switches back.
struct MeLane {
  qsvc_ctx *c;
  cudaStream_t main;
  explicit MeLane(qsvc_ctx *ctx) : c(ctx), main(ctx->stream) {
    c->stream = c->me_stream;
    c->lane = 1;
  }
  ~MeLane() {
    c->stream = main;
    c->lane = 0;
  }
};

static int analyze_levels(qsvc_ctx *c, const qsvc_analyze_params *p, const qsvc_level_out *outs) {
  if (!p || !c->low0) return fail(QSVC_EINVAL, "no resident sequence");
  if (p->pixels_in_x != c->X || p->pixels_in_y != c->Y) return fail(QSVC_EINVAL, "geometry mismatch");
  const int X = c->X, Y = c->Y;
  int pictures = c->n_frames;
  if (p->TRLs < 1) return fail(QSVC_EINVAL, "bad TRLs");
  if (p->TRLs > 1 && (pictures - 1) % (1 << (p->TRLs - 1)) != 0)
    return fail(QSVC_EINVAL, "n_frames=%d is not GOPs*2^(TRLs-1)+1 for TRLs=%d", pictures, p->TRLs);
  free_levels(c);
  c->levels.resize(p->TRLs);
  c->sad_ops = 0;
  const long long fb = frame_bytes(X, Y);
  int sr = p->search_range, bs = p->block_size, bs_min = p->block_size_min;
  if (bs < bs_min) bs_min = bs;  // analyze.py:118-119
  // update_factor == 0 (analyze.py's default): update is the identity, low_t = even_t, so the inputs of
  // every level are frames of the resident clip (even_t[k] = low_0[k << t]) and the motion estimation of
  // level t+1 does not wait for the decorrelate of level t: two lanes, joined by one event per level.
  const bool lanes = c->overlap != 0 && p->update_factor == 0.0f && p->TRLs > 2 && c->me_stream != nullptr;
  // update_factor != 0: level t+1 needs low_t, so the levels stay in order -- but the motion estimation of a level
  // and the preparation of its decorrelate (up-sampled reference planes, border ring) are independent of each
  // other and run on the two streams side by side
  const bool side = !lanes && c->overlap != 0 && p->update_factor != 0.0f && c->me_stream != nullptr;
  {  // result buffers of every level first: nothing the pool hands out below is still in use elsewhere
    int pics = pictures, b = bs;
    for (int t = 1; t < p->TRLs; t++) {
      const int n = pics / 2;
      LevelResult &lv = c->levels[t];
      lv.n_pairs = n;
      lv.block_size = b;
      const long long field = 4LL * (Y / b) * (X / b);
      TRY(pool_alloc(c, (size_t)fb * n, (void **)&lv.high));
      TRY(pool_alloc(c, (size_t)std::max<long long>(field, 1) * n * sizeof(short), (void **)&lv.motion));
      TRY(pool_alloc(c, (size_t)std::max<long long>(field, 1) * n * sizeof(short), (void **)&lv.motion_filtered));
      TRY(pool_alloc(c, (size_t)fb * (n + 1), (void **)&lv.low));
      pics = (pics + 1) / 2;
      b = std::max(b / 2, bs_min);
    }
  }
  const uint8_t *low = c->low0;
  CU(cudaEventRecord(c->ev2, c->stream));
  if (lanes || side) {
    // the ME lane starts where the main stream stands now
    CU(cudaStreamWaitEvent(c->me_stream, c->ev2, 0));
    while ((int)c->me_events.size() < p->TRLs) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->me_events.push_back(e);
    }
  }
  // per-level parameters (analyze.py:144-151)
  struct LevelPlan {
    int n, sr, bs;
  };
  std::vector<LevelPlan> plan(p->TRLs);
  for (int t = 1; t < p->TRLs; t++) {
    plan[t] = {pictures / 2, sr, bs};
    c->levels[t].search_range = sr;
    pictures = (pictures + 1) / 2;
    sr = std::min(sr * 2, 128);       // analyze.py:144-147
    bs = std::max(bs / 2, bs_min);    // analyze.py:149-151
  }
  // motion estimation of level t (two lanes: on the ME stream, frames of the resident clip)
  auto run_me = [&](int t, const uint8_t *low) -> int {
    const int n = plan[t].n, bsz = plan[t].bs;
    LevelResult &lv = c->levels[t];
    const long long field = 4LL * (Y / bsz) * (X / bsz);
    // split (split.cpp:229-341) is index arithmetic: even k = frame 2k, odd i = frame 2i+1 of low_{t-1}
    const long long in_stride = lanes ? (fb << t) : 2 * fb;
    const uint8_t *even = lanes ? c->low0 : low, *odd = even + in_stride / 2;
    if (side && t > 1) {
      // low_{t-1} is ready where the main stream stands now
      while ((int)c->level_events.size() <= p->TRLs + t) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->level_events.push_back(e);
      }
      CU(cudaEventRecord(c->level_events[p->TRLs + t], c->stream));
      CU(cudaStreamWaitEvent(c->me_stream, c->level_events[p->TRLs + t], 0));
    }
    std::unique_ptr<MeLane> lane_guard;
    if (lanes || side) lane_guard.reset(new MeLane(c));
    if (t == 1 && c->upload_gops > 0) {
      long long p0 = 0;  // first pair of the segment
      for (int g = 0; g < c->upload_gops; g++) {
        CU(cudaStreamWaitEvent(c->stream, c->upload_events[g], 0));
        const int np = (int)(c->upload_segs[g] - p0);
        TRY(me_level(c, even + p0 * in_stride, in_stride, odd + p0 * in_stride, in_stride, np, X, Y, bsz,
                     p->border_size, plan[t].sr, p->subpixel_accuracy, g == 0 ? p->first_gop_is_global_first : 0,
                     lv.motion + p0 * field));
        p0 = c->upload_segs[g];
      }
    } else {
      TRY(me_level(c, even, in_stride, odd, in_stride, n, X, Y, bsz, p->border_size, plan[t].sr, p->subpixel_accuracy,
                   p->first_gop_is_global_first, lv.motion));
    }
    if (lanes || side) CU(cudaEventRecord(c->me_events[t], c->stream));
    return QSVC_OK;
  };
  // Two lanes: the ME lane is enqueued one level ahead of the decorrelate lane, so that whatever blocks the host
  // inside a decorrelate (the frame-type decision, a GOP shard waiting for its neighbour's prediction tail)
  // leaves the GPU with the next level's motion estimation to run.
  if (lanes && p->TRLs > 1) TRY(run_me(1, low));
  for (int t = 1; t < p->TRLs; t++) {
    const int n = plan[t].n, bsz = plan[t].bs;
    LevelResult &lv = c->levels[t];
    const long long field = 4LL * (Y / bsz) * (X / bsz);
    const long long in_stride = lanes ? (fb << t) : 2 * fb;
    const uint8_t *even = lanes ? c->low0 : low, *odd = even + in_stride / 2;
    c->cur_level = t;
    if (!lanes) TRY(run_me(t, low));
    else if (t + 1 < p->TRLs) TRY(run_me(t + 1, low));
    if (lanes || side) c->mv_ready = c->me_events[t];  // awaited inside, after the reference planes are up-sampled
    // (the decorrelate of level 1 reads the clip on this stream before it meets the ME lane, which is the one
    // that waited segment by segment: mc_level waits for the upload itself -- the byte-plane path frame by
    // frame as the segments land, the literal path for the whole clip)
    {
      const int rc = mc_level(c, 1, even, in_stride, odd, in_stride, lv.motion, n, X, Y, bsz, p->block_overlaping,
                              plan[t].sr, p->subpixel_accuracy, p->always_B, nullptr, lv.high, fb, &lv.types,
                              lv.motion_filtered, nullptr);
      const int rw = await_motion(c);  // nothing read it (no pairs): the lanes still have to meet
      if (rc != QSVC_OK) return rc;
      if (rw != QSVC_OK) return rw;
    }
    TRY(update_level(c, 0, even, in_stride, lv.high, fb, lv.motion_filtered, lv.types.c_str(), n, X, Y,
                     bsz, p->update_factor, lv.low, fb));
    if (outs) {
      // stream this level's results to the host behind the next level's compute
      while ((int)c->level_events.size() <= t) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->level_events.push_back(e);
      }
      CU(cudaEventRecord(c->level_events[t], c->stream));
      CU(cudaStreamWaitEvent(c->copy_stream, c->level_events[t], 0));
      const qsvc_level_out &o = outs[t];
      if (o.high) CU(cudaMemcpyAsync(o.high, lv.high, (size_t)fb * n, cudaMemcpyDeviceToHost, c->copy_stream));
      if (o.motion)
        CU(cudaMemcpyAsync(o.motion, lv.motion, (size_t)field * n * sizeof(short), cudaMemcpyDeviceToHost, c->copy_stream));
      if (o.motion_filtered)
        CU(cudaMemcpyAsync(o.motion_filtered, lv.motion_filtered, (size_t)field * n * sizeof(short),
                           cudaMemcpyDeviceToHost, c->copy_stream));
      if (o.low) CU(cudaMemcpyAsync(o.low, lv.low, (size_t)fb * (n + 1), cudaMemcpyDeviceToHost, c->copy_stream));
    }
    low = lv.low;
  }
  c->cur_level = 0;
  CU(cudaEventRecord(c->ev3, c->stream));
  CU(cudaEventSynchronize(c->ev3));
  CU(cudaEventElapsedTime(&c->total_ms, c->ev2, c->ev3));
  return QSVC_OK;
}

int qsvc_resident_fetch(qsvc_ctx *c, int t, uint8_t *high, int16_t *motion, int16_t *motion_filtered,
                        char *types, uint8_t *low) {
  ENTER(c);
  if (t < 1 || t >= (int)c->levels.size()) return fail(QSVC_EINVAL, "no such level %d", t);
  LevelResult &lv = c->levels[t];
  const long long fb = frame_bytes(c->X, c->Y);
  const long long field = 4LL * (c->Y / lv.block_size) * (c->X / lv.block_size);
  if (high) CU(cudaMemcpyAsync(high, lv.high, (size_t)fb * lv.n_pairs, cudaMemcpyDeviceToHost, c->stream));
  if (motion)
    CU(cudaMemcpyAsync(motion, lv.motion, (size_t)field * lv.n_pairs * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
  if (motion_filtered)
    CU(cudaMemcpyAsync(motion_filtered, lv.motion_filtered, (size_t)field * lv.n_pairs * sizeof(short),
                       cudaMemcpyDeviceToHost, c->stream));
  if (low) CU(cudaMemcpyAsync(low, lv.low, (size_t)fb * (lv.n_pairs + 1), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (types) memcpy(types, lv.types.data(), (size_t)lv.n_pairs);
  return QSVC_OK;
}

int qsvc_resident_fetch_motion_residue(qsvc_ctx *c, int t, int16_t *residue) {
  ENTER(c);
  const int T = (int)c->levels.size();
  if (t < 1 || t >= T || !residue) return fail(QSVC_EINVAL, "no such level %d", t);
  LevelResult &lv = c->levels[t];
  if (!lv.motion_filtered) return fail(QSVC_EINVAL, "level %d holds no motion_filtered fields", t);
  const int by = c->Y / lv.block_size, bx = c->X / lv.block_size;
  const size_t bytes = (size_t)lv.n_pairs * 4 * by * bx * sizeof(short);
  if (!bytes) return QSVC_OK;
  Scratch s(c);
  short *d_out;
  TRY(s.get(bytes, (void **)&d_out));
  if (t == T - 1) {
    launch_mv_bidirectional(c->L(), lv.motion_filtered, d_out, lv.n_pairs, by * bx, 0);
  } else {
    LevelResult &up = c->levels[t + 1];
    if (up.block_size != lv.block_size)
      return fail(QSVC_EINVAL, "levels %d and %d use different block sizes", t, t + 1);
    launch_mv_interlevel(c->L(), lv.motion_filtered, up.motion_filtered, d_out, lv.n_pairs,
                         up.motion_filtered ? up.n_pairs : 0, by * bx, 0);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(residue, d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QSVC_OK;
}

int qsvc_resident_stats(qsvc_ctx *c, double *sad_ops, float *search_ms, float *total_ms) {
  if (!c) return fail(QSVC_EINVAL, "null context");
  if (sad_ops) *sad_ops = c->sad_ops;
  if (search_ms) *search_ms = c->search_ms;
  if (total_ms) *total_ms = c->total_ms;
  return QSVC_OK;
}

int qsvc_resident_push(qsvc_ctx *c, int t, int n_pairs, const uint8_t *high, const int16_t *motion,
                       const char *types, const uint8_t *low_top, int X, int Y, int bs) {
  ENTER(c);
  if (t < 1 || n_pairs < 1 || !high || !motion || !types) return fail(QSVC_EINVAL, "bad arguments");
  TRY(check_geometry(X, Y, bs, 0));
  if ((int)c->levels.size() <= t) c->levels.resize(t + 1);
  if (c->X != X || c->Y != Y) {
    c->X = X;
    c->Y = Y;
  }
  LevelResult &lv = c->levels[t];
  pool_free(c, lv.high);
  pool_free(c, lv.motion);
  pool_free(c, lv.low);
  lv.high = nullptr;
  lv.motion = nullptr;
  lv.low = nullptr;
  lv.n_pairs = n_pairs;
  lv.block_size = bs;
  lv.types.assign(types, (size_t)n_pairs);
  const long long fb = frame_bytes(X, Y), field = 4LL * (Y / bs) * (X / bs);
  TRY(pool_alloc(c, (size_t)fb * n_pairs, (void **)&lv.high));
  TRY(pool_alloc(c, (size_t)std::max<long long>(field, 1) * n_pairs * sizeof(short), (void **)&lv.motion));
  CU(cudaMemcpyAsync(lv.high, high, (size_t)fb * n_pairs, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(lv.motion, motion, (size_t)field * n_pairs * sizeof(short), cudaMemcpyHostToDevice, c->stream));
  if (low_top) {
    TRY(pool_alloc(c, (size_t)fb * (n_pairs + 1), (void **)&lv.low));
    CU(cudaMemcpyAsync(lv.low, low_top, (size_t)fb * (n_pairs + 1), cudaMemcpyHostToDevice, c->stream));
  }
  return QSVC_OK;
}

int qsvc_resident_synthesize(qsvc_ctx *c, const qsvc_analyze_params *p) {
  ENTER(c);
  if (!p) return fail(QSVC_EINVAL, "null params");
  const int T = p->TRLs;
  if (T < 2 || (int)c->levels.size() < T || !c->levels[T - 1].low)
    return fail(QSVC_EINVAL, "push every level (and low_{TRLs-1}) before synthesize");
  const int X = c->X, Y = c->Y;
  const long long fb = frame_bytes(X, Y);
  CU(cudaEventRecord(c->ev2, c->stream));
  for (int t = T - 1; t >= 1; t--) {
    LevelResult &lv = c->levels[t];
    if (!lv.high || !lv.motion) return fail(QSVC_EINVAL, "level %d was not pushed", t);
    const int n = lv.n_pairs;
    // search_range of level t (synthesize.py:113-121,143-153)
    int sr = p->search_range;
    for (int j = 1; j < t; j++) sr = std::min(sr * 2, 128);
    uint8_t *dst;  // low_{t-1}: 2n+1 frames, even_t at even positions, odd_t at odd positions
    c->cur_level = t;
    Scratch guard(c);  // an error below gives the block back to the pool
    TRY(guard.get((size_t)fb * (2 * n + 1), (void **)&dst));
    if (t - 1 >= 1 && c->levels[t - 1].n_pairs != 2 * n)
      return fail(QSVC_EINVAL, "level %d has %d pairs, expected %d", t - 1, c->levels[t - 1].n_pairs, 2 * n);
    TRY(update_level(c, 1, lv.low, fb, lv.high, fb, lv.motion, lv.types.c_str(), n, X, Y,
                     lv.block_size, p->update_factor, dst, 2 * fb));
    TRY(mc_level(c, 0, dst, 2 * fb, lv.high, fb, lv.motion, n, X, Y, lv.block_size,
                 p->block_overlaping, sr, p->subpixel_accuracy, 1, lv.types.c_str(), dst + fb, 2 * fb,
                 nullptr, nullptr, nullptr));
    guard.detach(dst);
    if (t - 1 >= 1) {
      pool_free(c, c->levels[t - 1].low);
      c->levels[t - 1].low = dst;
    } else {
      pool_free(c, c->low0);
      c->low0 = dst;
      c->n_frames = 2 * n + 1;
    }
  }
  CU(cudaEventRecord(c->ev3, c->stream));
  CU(cudaEventSynchronize(c->ev3));
  CU(cudaEventElapsedTime(&c->total_ms, c->ev2, c->ev3));
  return QSVC_OK;
}

int qsvc_resident_fetch_low0(qsvc_ctx *c, uint8_t *low0, int n_frames) {
  ENTER(c);
  if (!c->low0 || n_frames != c->n_frames) return fail(QSVC_EINVAL, "resident low_0 has %d frames", c->n_frames);
  CU(cudaMemcpyAsync(low0, c->low0, (size_t)frame_bytes(c->X, c->Y) * n_frames, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return QSVC_OK;
}

}  // extern "C"
