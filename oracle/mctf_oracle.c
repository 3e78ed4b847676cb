/*
 * mctf_oracle.c -- CPU restatement of the QSVC MCTF hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The
 * product path (qsvc_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py and the committed
 * fixtures under tests/golden/ (made by oracle/make_golden.py from the
 * UNMODIFIED reference tools compiled into oracle/_ref/) check every entry
 * point below byte-for-byte against the reference binaries.
 *
 * What is restated (all paths relative to /root/reference/trunk/src):
 *   5_3.cpp:39-115        integer 5/3 lifting, truncating division
 *   Haar.cpp:39-89        Haar lifting (motion-field up-sampling)
 *   dwt2d.cpp:76-175      in-place Mallat 2-D analysis / synthesis
 *   texture.cpp:34-113    bordered image allocation (incl. the row-pointer
 *                         shift bug) and edge replication (incl. region 6)
 *   motion_estimate.cpp:70-481,714-907   hierarchical +-1 bidirectional search
 *   decorrelate.cpp:69-189,508-1075      prediction, residue, I/B decision,
 *                                        and the inverse (correlate)
 *   update.cpp:50-148,439-679            update lifting step and its inverse
 *   entropy.cpp:20-34                    zero-order entropy
 *
 * The reference's results depend on where glibc places its row allocations
 * (SURVEY.md A.3): rows of a texture are separate malloc chunks, and some
 * border reads land in the chunk header or in the tail of the physically
 * preceding row.  To reproduce that independently of the real allocator this
 * file carves every texture from a private, zero-filled arena using glibc's
 * chunk geometry (16-byte granules, 8-byte size field in front of the user
 * pointer, PREV_INUSE bit set).  Everything else is a plain restatement of the
 * reference's loops.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#define PREV 0
#define NEXT 1
#define X_FIELD 0
#define Y_FIELD 1
#define LINE_MAX_SAMPLES 16384 /* texture.cpp:9 */

/* ------------------------------------------------------------------ arena */

typedef struct {
  unsigned char *base;
  size_t cap, cur;
} arena_t;

static int arena_init(arena_t *A, size_t cap) {
  A->base = mmap(NULL, cap, PROT_READ | PROT_WRITE,
                 MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (A->base == MAP_FAILED) return -1;
  A->cap = cap;
  A->cur = 0;
  return 0;
}

static void arena_free(arena_t *A) {
  if (A->base && A->base != MAP_FAILED) munmap(A->base, A->cap);
  A->base = NULL;
}

/* glibc malloc(req) on a fresh brk heap: chunk = align16(req + 8), >= 32;
 * the 8-byte size field (chunk | PREV_INUSE) sits right before the user
 * pointer; the next chunk follows immediately. */
static void *arena_new(arena_t *A, size_t req) {
  size_t chunk = (req + 8 + 15) & ~(size_t)15;
  if (chunk < 32) chunk = 32;
  if (A->cur + chunk + 32 > A->cap) {
    fprintf(stderr, "mctf_oracle: arena exhausted\n");
    abort();
  }
  uint64_t sz = (uint64_t)chunk | 1u;
  memcpy(A->base + A->cur + 8, &sz, 8);
  void *p = A->base + A->cur + 16;
  A->cur += chunk;
  return p;
}

static size_t tex_bytes(long y_dim, long x_dim, long b) {
  size_t rows = (size_t)(y_dim + 2 * b);
  size_t row_chunk = (((size_t)(x_dim + 2 * b) * 2 + 8 + 15) & ~(size_t)15);
  if (row_chunk < 32) row_chunk = 32;
  return rows * row_chunk + rows * 8 + 64;
}

/* ---------------------------------------------------------------- texture */

/* texture.cpp:34-46.  Only the first y_dim row pointers are shifted by the
 * border; the remaining 2*border rows stay unshifted (reference bug, kept). */
static short **tex_alloc(arena_t *A, int y_dim, int x_dim, int border) {
  short **data = arena_new(A, sizeof(short *) * (size_t)(y_dim + border * 2));
  for (int y = 0; y < y_dim + border * 2; y++)
    data[y] = arena_new(A, sizeof(short) * (size_t)(x_dim + border * 2));
  for (int y = 0; y < y_dim; y++) data[y] += border;
  data += border;
  return data;
}

/* texture.cpp:55-113, regions in source order; region 6 replicates the
 * bottom-RIGHT pixel into the bottom-left corner (reference bug, kept). */
static void tex_fill_border(short **d, int y_dim, int x_dim, int b) {
  for (int y = -b; y < 0; y++)
    for (int x = -b; x < 0; x++) d[y][x] = d[0][0];
  for (int y = -b; y < 0; y++)
    for (int x = 0; x < x_dim; x++) d[y][x] = d[0][x];
  for (int y = -b; y < 0; y++)
    for (int x = x_dim; x < x_dim + b; x++) d[y][x] = d[0][x_dim - 1];
  for (int y = 0; y < y_dim; y++)
    for (int x = -b; x < 0; x++) d[y][x] = d[y][0];
  for (int y = 0; y < y_dim; y++)
    for (int x = x_dim; x < x_dim + b; x++) d[y][x] = d[y][x_dim - 1];
  for (int y = y_dim; y < y_dim + b; y++)
    for (int x = -b; x < 0; x++) d[y][x] = d[y_dim - 1][x_dim - 1];
  for (int y = y_dim; y < y_dim + b; y++)
    for (int x = 0; x < x_dim; x++) d[y][x] = d[y_dim - 1][x];
  for (int y = y_dim; y < y_dim + b; y++)
    for (int x = x_dim; x < x_dim + b; x++) d[y][x] = d[y_dim - 1][x_dim - 1];
}

/* texture.cpp:122-129 */
static const uint8_t *tex_read(const uint8_t *src, short **img, int y_dim, int x_dim) {
  for (int y = 0; y < y_dim; y++)
    for (int x = 0; x < x_dim; x++) img[y][x] = *src++;
  return src;
}

/* texture.cpp:137-144: short -> unsigned char conversion truncates mod 256. */
static uint8_t *tex_write(uint8_t *dst, short **img, int y_dim, int x_dim) {
  for (int y = 0; y < y_dim; y++)
    for (int x = 0; x < x_dim; x++) *dst++ = (uint8_t)img[y][x];
  return dst;
}

/* ------------------------------------------------------------ 5/3 lifting */

typedef void (*lift_fn)(short *a, short *b, short *c, int n);

/* 5_3.cpp:39-52 */
static void f53_even_analyze(short *s, short *l, short *h, int n) {
  int i;
  for (i = 0; i < n / 2 - 1; i++) {
    int i2 = i << 1;
    h[i] = s[i2 + 1] - (s[i2] + s[i2 + 2]) / 2;
  }
  h[i] = s[n - 1] - s[n - 2];
  l[0] = s[0] + h[0] / 2;
  for (i = 1; i < n / 2; i++) {
    int i2 = i << 1;
    l[i] = s[i2] + (h[i] + h[i - 1]) / 4;
  }
}

/* 5_3.cpp:60-73 */
static void f53_odd_analyze(short *s, short *l, short *h, int n) {
  int i;
  for (i = 0; i < n / 2; i++) {
    int i2 = i << 1;
    h[i] = s[i2 + 1] - (s[i2] + s[i2 + 2]) / 2;
  }
  l[0] = s[0] + h[0] / 2;
  for (i = 1; i < n / 2; i++) {
    int i2 = i << 1;
    l[i] = s[i2] + (h[i] + h[i - 1]) / 4;
  }
  l[i] = s[n - 1] + h[i - 1] / 2;
}

/* 5_3.cpp:81-94 */
static void f53_even_synthesize(short *s, short *l, short *h, int n) {
  int i;
  s[0] = l[0] - h[0] / 2;
  for (i = 1; i < n / 2; i++) {
    int i2 = i << 1;
    s[i2] = l[i] - (h[i] + h[i - 1]) / 4;
  }
  for (i = 0; i < n / 2 - 1; i++) {
    int i2 = i << 1;
    s[i2 + 1] = h[i] + (s[i2] + s[i2 + 2]) / 2;
  }
  s[n - 1] = h[i] + s[n - 2];
}

/* 5_3.cpp:102-115 */
static void f53_odd_synthesize(short *s, short *l, short *h, int n) {
  int i;
  s[0] = l[0] - h[0] / 2;
  for (i = 1; i < n / 2; i++) {
    int i2 = i << 1;
    s[i2] = l[i] - (h[i] + h[i - 1]) / 4;
  }
  s[n - 1] = l[i] - h[i - 1] / 2;
  for (i = 0; i < n / 2; i++) {
    int i2 = i << 1;
    s[i2 + 1] = h[i] + (s[i2] + s[i2 + 2]) / 2;
  }
}

/* Haar.cpp:68-89 (only synthesis is used on the hot path) */
static void haar_even_synthesize(short *s, short *l, short *h, int n) {
  int i, k;
  for (i = k = 0; k < n; i++, k += 2) {
    s[k] = l[i] - h[i] / 2;
    s[k + 1] = s[k] + h[i];
  }
}
static void haar_odd_synthesize(short *s, short *l, short *h, int n) {
  int i, k;
  for (i = k = 0; k < (n - 1); i++, k += 2) {
    s[k] = l[i] - h[i] / 2;
    s[k + 1] = s[k] + h[i];
  }
  s[k] = l[i];
}

/* ------------------------------------------------------------------ dwt2d */

typedef struct {
  lift_fn even_analyze, odd_analyze, even_synthesize, odd_synthesize;
  short *in_line, *out_line;
} dwt_t;

static void dwt_init(dwt_t *d, int haar) {
  if (haar) {
    d->even_analyze = d->odd_analyze = NULL;
    d->even_synthesize = haar_even_synthesize;
    d->odd_synthesize = haar_odd_synthesize;
  } else {
    d->even_analyze = f53_even_analyze;
    d->odd_analyze = f53_odd_analyze;
    d->even_synthesize = f53_even_synthesize;
    d->odd_synthesize = f53_odd_synthesize;
  }
  d->in_line = calloc(LINE_MAX_SAMPLES + 8, sizeof(short));
  d->out_line = calloc(LINE_MAX_SAMPLES + 8, sizeof(short));
}
static void dwt_done(dwt_t *d) {
  free(d->in_line);
  free(d->out_line);
}

/* dwt2d.cpp:76-119: rows, then columns; sizes halve with floor. */
static void dwt_analyze(dwt_t *d, short **sig, int y, int x, int levels) {
  short *in_line = d->in_line, *out_line = d->out_line;
  for (int lv = 0; lv < levels; lv++) {
    int nx = x;
    x >>= 1;
    int ny = y;
    y >>= 1;
    if (y == 0) y = 1;
    if (x == 0) x = 1;
    if (nx & 1) {
      for (int j = 0; j < ny; j++) {
        memcpy(in_line, sig[j], nx * sizeof(short));
        d->odd_analyze(in_line, sig[j], sig[j] + x + 1, nx);
      }
    } else {
      for (int j = 0; j < ny; j++) {
        memcpy(in_line, sig[j], nx * sizeof(short));
        d->even_analyze(in_line, sig[j], sig[j] + x, nx);
      }
    }
    if (ny & 1) {
      for (int i = 0; i < nx; i++) {
        for (int j = 0; j < ny; j++) in_line[j] = sig[j][i];
        d->odd_analyze(in_line, out_line, out_line + y + 1, ny);
        for (int j = 0; j < ny; j++) sig[j][i] = out_line[j];
      }
    } else {
      for (int i = 0; i < nx; i++) {
        for (int j = 0; j < ny; j++) in_line[j] = sig[j][i];
        d->even_analyze(in_line, out_line, out_line + y, ny);
        for (int j = 0; j < ny; j++) sig[j][i] = out_line[j];
      }
    }
  }
}

/* dwt2d.cpp:128-175: columns, then rows. */
static void dwt_synthesize(dwt_t *d, short **sig, int y, int x, int levels) {
  short *in_line = d->in_line, *out_line = d->out_line;
  int nx = x >> levels;
  int ny = y >> levels;
  for (int lv = levels - 1; lv >= 0; lv--) {
    int mx = nx;
    nx = x >> lv;
    int my = ny;
    ny = y >> lv;
    if (nx == 0) nx = 1;
    if (ny == 0) ny = 1;
    if (ny & 1) {
      for (int i = 0; i < nx; i++) {
        for (int j = 0; j < ny; j++) in_line[j] = sig[j][i];
        d->odd_synthesize(out_line, in_line, in_line + my + 1, ny);
        for (int j = 0; j < ny; j++) sig[j][i] = out_line[j];
      }
    } else {
      for (int i = 0; i < nx; i++) {
        for (int j = 0; j < ny; j++) in_line[j] = sig[j][i];
        d->even_synthesize(out_line, in_line, in_line + my, ny);
        for (int j = 0; j < ny; j++) sig[j][i] = out_line[j];
      }
    }
    if (nx & 1) {
      for (int j = 0; j < ny; j++) {
        memcpy(in_line, sig[j], nx * sizeof(short));
        d->odd_synthesize(sig[j], in_line, in_line + mx + 1, nx);
      }
    } else {
      for (int j = 0; j < ny; j++) {
        memcpy(in_line, sig[j], nx * sizeof(short));
        d->even_synthesize(sig[j], in_line, in_line + mx, nx);
      }
    }
  }
}

/* ----------------------------------------------------------------- motion */

/* motion.cpp:36-48: [PREV|NEXT][X|Y][by][bx] */
typedef struct {
  short **p[2][2];
  int by, bx;
} mvf_t;

static void mvf_alloc(mvf_t *m, int by, int bx) {
  m->by = by;
  m->bx = bx;
  for (int i = 0; i < 2; i++)
    for (int f = 0; f < 2; f++) {
      m->p[i][f] = calloc((size_t)(by > 0 ? by : 1), sizeof(short *));
      for (int y = 0; y < by; y++) m->p[i][f][y] = calloc((size_t)(bx > 0 ? bx : 1) + 2, sizeof(short));
    }
}
static void mvf_free(mvf_t *m) {
  for (int i = 0; i < 2; i++)
    for (int f = 0; f < 2; f++) {
      for (int y = 0; y < m->by; y++) free(m->p[i][f][y]);
      free(m->p[i][f]);
    }
}
/* motion.cpp:72-101: planes PREV.X, PREV.Y, NEXT.X, NEXT.Y, row-major int16 */
static const int16_t *mvf_read(const int16_t *src, mvf_t *m) {
  for (int i = 0; i < 2; i++)
    for (int f = 0; f < 2; f++)
      for (int y = 0; y < m->by; y++) {
        memcpy(m->p[i][f][y], src, sizeof(short) * (size_t)m->bx);
        src += m->bx;
      }
  return src;
}
static int16_t *mvf_write(int16_t *dst, mvf_t *m) {
  for (int i = 0; i < 2; i++)
    for (int f = 0; f < 2; f++)
      for (int y = 0; y < m->by; y++) {
        memcpy(dst, m->p[i][f][y], sizeof(short) * (size_t)m->bx);
        dst += m->bx;
      }
  return dst;
}

/* -------------------------------------------------------- motion_estimate */

/* motion_estimate.cpp:232-236 */
static int desp(int x, int y) {
  for (int i = 0; i < y; i++) x = (x + 1) / 2;
  return x;
}

/* motion_estimate.cpp:70-184.  Nine candidates in the reference's order,
 * PREV tests centre+delta, NEXT tests centre-delta, "<=" keeps the later
 * candidate among equal errors; both directions start from the OLD centre. */
static void local_me_for_block(mvf_t *mv, short ***ref, short **pred, int luby, int lubx,
                               int rbby, int rbbx, int by, int bx) {
  static const int cand[9][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}, {-1, 0},
                                 {1, 0},   {0, 1},  {0, -1}, {0, 0}};
  int min_error[2] = {0, 0};
  int vy[2] = {0, 0}, vx[2] = {0, 0};
  short c_py = mv->p[PREV][Y_FIELD][by][bx], c_px = mv->p[PREV][X_FIELD][by][bx];
  short c_ny = mv->p[NEXT][Y_FIELD][by][bx], c_nx = mv->p[NEXT][X_FIELD][by][bx];
  for (int k = 0; k < 9; k++) {
    int dy = cand[k][0], dx = cand[k][1];
    short y[2] = {(short)(c_py + dy), (short)(c_ny - dy)};
    short x[2] = {(short)(c_px + dx), (short)(c_nx - dx)};
    int error[2] = {0, 0};
    for (int py = luby; py < rbby; py++) {
      short *pred_py = pred[py];
      for (int px = lubx; px < rbbx; px++) {
        error[PREV] += abs(pred_py[px] - ref[PREV][py + y[PREV]][px + x[PREV]]);
        error[NEXT] += abs(pred_py[px] - ref[NEXT][py + y[NEXT]][px + x[NEXT]]);
      }
    }
    for (int d = 0; d < 2; d++)
      if (k == 0 || error[d] <= min_error[d]) {
        vy[d] = y[d];
        vx[d] = x[d];
        min_error[d] = error[d];
      }
  }
  mv->p[PREV][Y_FIELD][by][bx] = vy[PREV];
  mv->p[PREV][X_FIELD][by][bx] = vx[PREV];
  mv->p[NEXT][Y_FIELD][by][bx] = vy[NEXT];
  mv->p[NEXT][X_FIELD][by][bx] = vx[NEXT];
}

/* motion_estimate.cpp:196-225 */
static void local_me_for_image(mvf_t *mv, short ***ref, short **pred, int block_size,
                               int border_size, int blocks_in_y, int blocks_in_x) {
  for (int by = 0; by < blocks_in_y; by++)
    for (int bx = 0; bx < blocks_in_x; bx++)
      local_me_for_block(mv, ref, pred, by * block_size - border_size,
                         bx * block_size - border_size, (by + 1) * block_size + border_size,
                         (bx + 1) * block_size + border_size, by, bx);
}

static void mv_double_clamp(mvf_t *mv, int by, int bx, int lim) {
  for (int y = 0; y < by; y++)
    for (int x = 0; x < bx; x++)
      for (int i = 0; i < 2; i++)
        for (int f = 0; f < 2; f++) {
          short *p = &mv->p[i][f][y][x];
          *p *= 2;
          if (*p > lim) *p = lim;
          if (*p < -lim) *p = -lim;
        }
}

/* motion_estimate.cpp:260-413 (FAST_SEARCH branch) */
static void me_for_image(mvf_t *mv, short ***ref, short **pred, int Y, int X, int bs, int bd,
                         int a, int sr, int by, int bx, dwt_t *pic, dwt_t *mvd) {
  int L = (int)rint(log((double)sr) / log(2.0)) - 1;
  dwt_analyze(pic, ref[PREV], Y, X, L);
  dwt_analyze(pic, ref[NEXT], Y, X, L);
  dwt_analyze(pic, pred, Y, X, L);
  local_me_for_image(mv, ref, pred, bs, bd, desp(by, L), desp(bx, L));
  for (int l = L - 1; l >= 0; --l) {
    int Y_l = desp(Y, l), X_l = desp(X, l);
    int by_l = desp(by, l), bx_l = desp(bx, l);
    dwt_synthesize(pic, ref[PREV], Y_l, X_l, 1);
    dwt_synthesize(pic, ref[NEXT], Y_l, X_l, 1);
    dwt_synthesize(pic, pred, Y_l, X_l, 1);
    dwt_synthesize(mvd, mv->p[PREV][Y_FIELD], by_l, bx_l, 1);
    dwt_synthesize(mvd, mv->p[NEXT][Y_FIELD], by_l, bx_l, 1);
    dwt_synthesize(mvd, mv->p[PREV][X_FIELD], by_l, bx_l, 1);
    dwt_synthesize(mvd, mv->p[NEXT][X_FIELD], by_l, bx_l, 1);
    mv_double_clamp(mv, by_l, bx_l, sr);
    local_me_for_image(mv, ref, pred, bs, bd, by_l, bx_l);
  }
  for (int l = 1; l <= a; l++) {
    dwt_synthesize(pic, ref[PREV], Y << l, X << l, 1);
    dwt_synthesize(pic, ref[NEXT], Y << l, X << l, 1);
    dwt_synthesize(pic, pred, Y << l, X << l, 1);
    mv_double_clamp(mv, by, bx, sr << a);
    local_me_for_image(mv, ref, pred, bs << l, bd >> l, by, bx);
  }
  dwt_analyze(pic, ref[PREV], Y << a, X << a, a);
  dwt_analyze(pic, ref[NEXT], Y << a, X << a, a);
  dwt_analyze(pic, pred, Y << a, X << a, a);
}

static size_t frame_bytes(int X, int Y) {
  return (size_t)X * Y + 2 * (size_t)(X / 2) * (Y / 2);
}

/* motion_estimate.cpp:714-907.  even: pictures/2+1 I420 frames, odd: pictures/2.
 * mv_out: pictures/2 fields of 4*by*bx int16. */
int orc_motion_estimate(const uint8_t *even, const uint8_t *odd, int pictures, int X, int Y,
                        int block_size, int border_size, int search_range, int a,
                        int16_t *mv_out) {
  int pb = search_range + border_size;
  arena_t A;
  if (arena_init(&A, 3 * tex_bytes((long)Y << a, (long)X << a, (long)pb << a) + (1 << 20))) return -1;
  short **reference[2];
  for (int i = 0; i < 2; i++) {
    reference[i] = tex_alloc(&A, Y << a, X << a, pb << a);
    for (int y = 0; y < Y << a; y++)
      for (int x = 0; x < X << a; x++) reference[i][y][x] = 0;
  }
  short **predicted = tex_alloc(&A, Y << a, X << a, pb << a);
  for (int y = 0; y < Y << a; y++)
    for (int x = 0; x < X << a; x++) predicted[y][x] = 0;
  int by = Y / block_size, bx = X / block_size;
  mvf_t mv;
  mvf_alloc(&mv, by, bx);
  dwt_t pic, mvd;
  dwt_init(&pic, 0);
  dwt_init(&mvd, 1);
  size_t fb = frame_bytes(X, Y);

  tex_read(even, reference[0], Y, X);
  tex_fill_border(reference[0], Y, X, pb);
  for (int i = 0; i < pictures / 2; i++) {
    tex_read(odd + (size_t)i * fb, predicted, Y, X);
    for (int y = 0; y < Y << a; y++)
      for (int x = 0; x < X << a; x++) reference[1][y][x] = 0;
    tex_read(even + (size_t)(i + 1) * fb, reference[1], Y, X);
    tex_fill_border(reference[1], Y, X, pb);
    for (int y = 0; y < by; y++)
      for (int x = 0; x < bx; x++)
        mv.p[0][0][y][x] = mv.p[0][1][y][x] = mv.p[1][0][y][x] = mv.p[1][1][y][x] = 0;
    me_for_image(&mv, reference, predicted, Y, X, block_size, border_size, a, search_range, by,
                 bx, &pic, &mvd);
    mv_out = mvf_write(mv_out, &mv);
    short **tmp = reference[0];
    reference[0] = reference[1];
    reference[1] = tmp;
  }
  dwt_done(&pic);
  dwt_done(&mvd);
  mvf_free(&mv);
  arena_free(&A);
  return 0;
}

/* ---------------------------------------------------------------- entropy */

/* entropy.cpp:20-34.  `log(prob)` on a float argument resolves to the float
 * overload in the reference's C++ (logf); the quotient is taken in double and
 * rounded to float before the float multiply-accumulate. */
static float entropy(const int *count, int n) {
  float e = 0.0f;
  int total = 0;
  for (int i = 0; i < n; i++) total += count[i];
  for (int i = 0; i < n; i++)
    if (count[i]) {
      float prob = (float)count[i] / total;
      e += prob * (float)(logf(prob) / log(2.0));
    }
  return -e;
}

float orc_entropy(const int *count, int n) { return entropy(count, n); }

/* ---------------------------------------------------- decorrelate/correlate */

/* decorrelate.cpp:69-189 */
static void predict(int ov, int bs, int by, int bx, int comps, int Y, int X, mvf_t *mv,
                    dwt_t *dwt, short **blk, short ***pred_pic, short ****ref) {
  int dwt_border = ov;
  int levels = 0;
  if (ov > 0) levels = (int)rint(log((double)ov) / log(2.0));
  for (int c = 0; c < comps; c++) {
    for (int yb = 0; yb < by; yb++) {
      for (int xb = 0; xb < bx; xb++) {
        int mvy0 = mv->p[PREV][Y_FIELD][yb][xb] + yb * bs;
        int mvy1 = mv->p[NEXT][Y_FIELD][yb][xb] + yb * bs;
        int mvx0 = mv->p[PREV][X_FIELD][yb][xb] + xb * bs;
        int mvx1 = mv->p[NEXT][X_FIELD][yb][xb] + xb * bs;
        for (int y = -dwt_border; y < bs + dwt_border; y++)
          for (int x = -dwt_border; x < bs + dwt_border; x++)
            blk[y + dwt_border][x + dwt_border] =
                (ref[PREV][c][mvy0 + y][mvx0 + x] + ref[NEXT][c][mvy1 + y][mvx1 + x]) / 2;
        dwt_analyze(dwt, blk, bs + dwt_border * 2, bs + dwt_border * 2, levels);
        for (int l = 1; l <= levels; l++) {
          int s = bs >> l;
          for (int y = 0; y < s; y++)
            for (int x = 0; x < s; x++) {
              pred_pic[c][yb * s + y][(X >> l) + xb * s + x] =
                  blk[(dwt_border >> l) + y][((bs + dwt_border * 3) >> l) + x];
              pred_pic[c][(Y >> l) + yb * s + y][xb * s + x] =
                  blk[((bs + dwt_border * 3) >> l) + y][(dwt_border >> l) + x];
              pred_pic[c][(Y >> l) + yb * s + y][(X >> l) + xb * s + x] =
                  blk[((bs + dwt_border * 3) >> l) + y][((bs + dwt_border * 3) >> l) + x];
            }
        }
        int s = bs >> levels;
        for (int y = 0; y < s; y++)
          for (int x = 0; x < s; x++)
            pred_pic[c][yb * s + y][xb * s + x] =
                blk[(dwt_border >> levels) + y][(dwt_border >> levels) + x];
      }
    }
    dwt_synthesize(dwt, pred_pic[c], Y, X, levels);
  }
}

/* decorrelate.cpp:583-686 / 732-788: load the three components, bring chroma to
 * luma size and every component to 2^a resolution by zero-high-band synthesis,
 * then replicate edges. */
static const uint8_t *load_reference(const uint8_t *src, short ***ref, const int *piy,
                                     const int *pix, int a, int pb, dwt_t *dwt) {
  for (int c = 0; c < 3; c++) src = tex_read(src, ref[c], piy[c], pix[c]);
  for (int c = 1; c < 3; c++) {
    for (int y = 0; y < piy[0] / 2; y++)
      memset(ref[c][y] + pix[0] / 2, 0, (pix[0] * sizeof(short)) / 2);
    for (int y = piy[0] / 2; y < piy[0]; y++) memset(ref[c][y], 0, pix[0] * sizeof(short));
    dwt_synthesize(dwt, ref[c], piy[0], pix[0], 1);
  }
  for (int c = 0; c < 3; c++) {
    for (int s = 1; s <= a; s++) {
      for (int y = 0; y < (piy[0] << s) / 2; y++)
        memset(ref[c][y] + (pix[0] << s) / 2, 0, ((pix[0] << s) / 2) * sizeof(short));
      for (int y = (piy[0] << s) / 2; y < (piy[0] << s); y++)
        memset(ref[c][y], 0, (pix[0] << s) * sizeof(short));
      dwt_synthesize(dwt, ref[c], piy[0] << s, pix[0] << s, 1);
    }
    tex_fill_border(ref[c], piy[0] << a, pix[0] << a, pb << a);
  }
  return src;
}

/* decorrelate.cpp:199-1078.  analyze!=0: decorrelate (reads odd, writes high,
 * frame_types, mv_out); analyze==0: correlate (reads high, frame_types, writes
 * odd).  prediction_out (nullable) receives the prediction_<even_fn> side file.
 * Returns 0, or 1 when always_B==0 and a motion component fell outside the
 * 256-bin histogram (undefined behaviour in the reference; here such
 * components are dropped from the histogram). */
int orc_decorrelate(int analyze, const uint8_t *even, const uint8_t *odd_in,
                    const uint8_t *high_in, const int16_t *mv_in, const char *types_in,
                    int pictures, int X, int Y, int block_size, int ov, int search_range, int a,
                    int always_B, uint8_t *high_out, uint8_t *odd_out, char *types_out,
                    int16_t *mv_out, uint8_t *prediction_out) {
  int rc = 0;
  int pix[3] = {X, X / 2, X / 2}, piy[3] = {Y, Y / 2, Y / 2};
  int by = piy[0] / block_size, bx = pix[0] / block_size;
  int pb = 4 * search_range + ov;
  size_t need = 6 * tex_bytes((long)Y << a, (long)X << a, (long)pb << a) +
                3 * tex_bytes((long)Y << a, (long)X << a, 0) + 6 * tex_bytes(Y, X, pb) +
                tex_bytes((long)(block_size + 2 * ov + 2) << a, (long)(block_size + 2 * ov + 2) << a, 0) +
                (1 << 20);
  arena_t A;
  if (arena_init(&A, need)) return -1;
  dwt_t dwt;
  dwt_init(&dwt, 0);
  mvf_t mv, zeroes;
  mvf_alloc(&mv, by, bx);
  mvf_alloc(&zeroes, by, bx);
  short **blk = tex_alloc(&A, (by ? piy[0] / by + ov * 2 : 1) << a, (bx ? pix[0] / bx + ov * 2 : 1) << a, 0);
  short **refbuf[2][3];
  short ***reference[2] = {refbuf[0], refbuf[1]};
  for (int i = 0; i < 2; i++)
    for (int c = 0; c < 3; c++) refbuf[i][c] = tex_alloc(&A, piy[0] << a, pix[0] << a, pb << a);
  short **predicted[3], **prediction[3], **residue[3];
  for (int c = 0; c < 3; c++) predicted[c] = tex_alloc(&A, piy[c], pix[c], pb);
  for (int c = 0; c < 3; c++) prediction[c] = tex_alloc(&A, piy[0] << a, pix[0] << a, 0);
  for (int c = 0; c < 3; c++) residue[c] = tex_alloc(&A, piy[c], pix[c], 0);
  size_t fb = frame_bytes(X, Y);

  load_reference(even, reference[0], piy, pix, a, pb, &dwt);
  for (int i = 0; i < pictures / 2; i++) {
    if (analyze) {
      const uint8_t *s = odd_in + (size_t)i * fb;
      for (int c = 0; c < 3; c++) s = tex_read(s, predicted[c], piy[c], pix[c]);
    } else {
      const uint8_t *s = high_in + (size_t)i * fb;
      for (int c = 0; c < 3; c++) {
        s = tex_read(s, residue[c], piy[c], pix[c]);
        for (int y = 0; y < piy[c]; y++)
          for (int x = 0; x < pix[c]; x++) residue[c][y][x] -= 128;
      }
    }
    load_reference(even + (size_t)(i + 1) * fb, reference[1], piy, pix, a, pb, &dwt);
    mv_in = mvf_read(mv_in, &mv);

    float motion_entropy = 0.0f;
    if (analyze && !always_B) {
      int count[256];
      memset(count, 0, sizeof count);
      for (int y = 0; y < by; y++)
        for (int x = 0; x < bx; x++)
          for (int k = 0; k < 4; k++) {
            /* reference order: PREV.Y, PREV.X, NEXT.Y, NEXT.X (decorrelate.cpp:809-812) */
            int v = mv.p[k >> 1][(k & 1) ? X_FIELD : Y_FIELD][y][x] + 128;
            if (v < 0 || v > 255) {
              rc = 1;
              continue;
            }
            count[v]++;
          }
      motion_entropy = entropy(count, 256);
    }

    {
      short ***refs[2] = {reference[0], reference[1]};
      predict(ov << a, block_size << a, by, bx, 3, piy[0] << a, pix[0] << a, &mv, &dwt, blk,
              prediction, refs);
    }
    for (int c = 0; c < 3; c++)
      for (int y = 0; y < piy[0] << a; y++)
        for (int x = 0; x < pix[0] << a; x++) {
          if (prediction[c][y][x] < 0)
            prediction[c][y][x] = 0;
          else if (prediction[c][y][x] > 255)
            prediction[c][y][x] = 255;
        }
    for (int c = 0; c < 3; c++) dwt_analyze(&dwt, prediction[c], piy[0] << a, pix[0] << a, a);
    dwt_analyze(&dwt, prediction[1], piy[0], pix[0], 1);
    dwt_analyze(&dwt, prediction[2], piy[0], pix[0], 1);
    if (prediction_out)
      for (int c = 0; c < 3; c++) prediction_out = tex_write(prediction_out, prediction[c], piy[c], pix[c]);

    if (analyze) {
      for (int c = 0; c < 3; c++)
        for (int y = 0; y < piy[c]; y++)
          for (int x = 0; x < pix[c]; x++) {
            int val = predicted[c][y][x] - prediction[c][y][x];
            if (val < -128)
              val = -128;
            else if (val > 127)
              val = 127;
            residue[c][y][x] = val;
          }
      float residue_entropy = 0.0f, predicted_entropy = 1.0f;
      if (!always_B) {
        int predicted_count[256], residue_count[256];
        memset(predicted_count, 0, sizeof predicted_count);
        memset(residue_count, 0, sizeof residue_count);
        for (int y = 0; y < piy[0]; y++)
          for (int x = 0; x < pix[0]; x++) {
            predicted_count[predicted[0][y][x]]++;
            residue_count[residue[0][y][x] + 128]++;
          }
        predicted_entropy = entropy(predicted_count, 256);
        residue_entropy = entropy(residue_count, 256);
      }
      int predicted_size = (int)(predicted_entropy * (float)piy[0] * (float)pix[0]);
      int residue_size = (int)(residue_entropy * (float)piy[0] * (float)pix[0]);
      int motion_size = (int)(motion_entropy * (float)by * (float)bx);
      if (predicted_size <= (residue_size + motion_size)) {
        *types_out++ = 'I';
        for (int c = 0; c < 3; c++)
          for (int y = 0; y < piy[c]; y++)
            for (int x = 0; x < pix[c]; x++) residue[c][y][x] = predicted[c][y][x];
        for (int c = 0; c < 3; c++) high_out = tex_write(high_out, residue[c], piy[c], pix[c]);
        mv_out = mvf_write(mv_out, &zeroes);
      } else {
        *types_out++ = 'B';
        for (int c = 0; c < 3; c++) {
          for (int y = 0; y < piy[c]; y++)
            for (int x = 0; x < pix[c]; x++) {
              int val = residue[c][y][x] + 128;
              if (val < 0)
                val = 0;
              else if (val > 255)
                val = 255;
              residue[c][y][x] = val;
            }
          high_out = tex_write(high_out, residue[c], piy[c], pix[c]);
        }
        mv_out = mvf_write(mv_out, &mv);
      }
    } else {
      if (*types_in++ == 'I') {
        for (int c = 0; c < 3; c++)
          for (int y = 0; y < piy[c]; y++)
            for (int x = 0; x < pix[c]; x++) predicted[c][y][x] = residue[c][y][x] + 128;
      } else {
        for (int c = 0; c < 3; c++)
          for (int y = 0; y < piy[c]; y++)
            for (int x = 0; x < pix[c]; x++) {
              int val = residue[c][y][x] + prediction[c][y][x];
              if (val < 0)
                val = 0;
              else if (val > 255)
                val = 255;
              predicted[c][y][x] = val;
            }
      }
      for (int c = 0; c < 3; c++) odd_out = tex_write(odd_out, predicted[c], piy[c], pix[c]);
    }
    short ***tmp = reference[0];
    reference[0] = reference[1];
    reference[1] = tmp;
  }
  mvf_free(&mv);
  mvf_free(&zeroes);
  dwt_done(&dwt);
  arena_free(&A);
  return rc;
}

/* ------------------------------------------------------- update/un_update */

/* update.cpp:50-54 */
static int clip(int x, int dim) {
  if (x < 0) return 0;
  if (x >= dim) return dim - 1;
  return x;
}

/* update.cpp:71-148.  Sequential in-place scatter; float multiply and add are
 * separate roundings (the reference is built for baseline x86-64: no FMA). */
static void update_step(int analyze, int bs, int by, int bx, mvf_t *mv, int Y, int X,
                        short ****ref, short ***residue, float uf) {
  for (int c = 0; c < 3; c++)
    for (int yb = 0; yb < by; yb++)
      for (int xb = 0; xb < bx; xb++)
        for (int y = 0; y < bs; y++)
          for (int x = 0; x < bs; x++)
            for (int d = 0; d < 2; d++) {
              short *t = &ref[d][c][clip(yb * bs + y + mv->p[d][Y_FIELD][yb][xb], Y)]
                                   [clip(xb * bs + x + mv->p[d][X_FIELD][yb][xb], X)];
              volatile float prod = residue[c][yb * bs + y][xb * bs + x] * uf;
              float aux = *t;
              if (analyze)
                aux += prod;
              else
                aux -= prod;
              if (aux > 255)
                aux = 255;
              else if (aux < 0)
                aux = 0;
              *t = aux;
            }
}

static const uint8_t *load_update_reference(const uint8_t *src, short ***ref, const int *piy,
                                            const int *pix, dwt_t *dwt) {
  for (int c = 0; c < 3; c++) src = tex_read(src, ref[c], piy[c], pix[c]);
  for (int c = 1; c < 3; c++) {
    for (int y = 0; y < piy[0] / 2; y++)
      for (int x = pix[0] / 2; x < pix[0]; x++) ref[c][y][x] = 0;
    for (int y = piy[0] / 2; y < piy[0]; y++)
      for (int x = 0; x < pix[0]; x++) ref[c][y][x] = 0;
    dwt_synthesize(dwt, ref[c], piy[0], pix[0], 1);
  }
  return src;
}

/* update.cpp:158-684.  analyze!=0: update (in = even_t, out = low_t);
 * analyze==0: un_update (in = low_t, out = even_t).  in/out hold
 * pictures/2+1 frames, high pictures/2 frames. */
int orc_update(int analyze, const uint8_t *in, const uint8_t *high, const int16_t *mv_in,
               const char *types, int pictures, int X, int Y, int block_size, float uf,
               uint8_t *out) {
  int pix[3] = {X, X / 2, X / 2}, piy[3] = {Y, Y / 2, Y / 2};
  int by = piy[0] / block_size, bx = pix[0] / block_size;
  arena_t A;
  if (arena_init(&A, 9 * tex_bytes(Y, X, 0) + (1 << 20))) return -1;
  dwt_t dwt;
  dwt_init(&dwt, 0);
  mvf_t mv;
  mvf_alloc(&mv, by, bx);
  short **refbuf[2][3], **residue[3];
  short ***reference[2] = {refbuf[0], refbuf[1]};
  for (int i = 0; i < 2; i++)
    for (int c = 0; c < 3; c++) refbuf[i][c] = tex_alloc(&A, piy[0], pix[0], 0);
  for (int c = 0; c < 3; c++) residue[c] = tex_alloc(&A, piy[0], pix[0], 0);
  size_t fb = frame_bytes(X, Y);

  load_update_reference(in, reference[0], piy, pix, &dwt);
  int i = 0;
  for (; i < pictures / 2; i++) {
    const uint8_t *s = high + (size_t)i * fb;
    for (int c = 0; c < 3; c++) {
      s = tex_read(s, residue[c], piy[c], pix[c]);
      for (int y = 0; y < piy[c]; y++)
        for (int x = 0; x < pix[c]; x++) residue[c][y][x] -= 128;
    }
    load_update_reference(in + (size_t)(i + 1) * fb, reference[1], piy, pix, &dwt);
    mv_in = mvf_read(mv_in, &mv);
    if (types[i] == 'B') {
      short ***refs[2] = {reference[0], reference[1]};
      update_step(analyze, block_size, by, bx, &mv, piy[0], pix[0], refs, residue, uf);
    }
    dwt_analyze(&dwt, reference[0][1], piy[0], pix[0], 1);
    dwt_analyze(&dwt, reference[0][2], piy[0], pix[0], 1);
    for (int c = 0; c < 3; c++) out = tex_write(out, reference[0][c], piy[c], pix[c]);
    short ***tmp = reference[0];
    reference[0] = reference[1];
    reference[1] = tmp;
  }
  dwt_analyze(&dwt, reference[0][1], piy[0], pix[0], 1);
  dwt_analyze(&dwt, reference[0][2], piy[0], pix[0], 1);
  for (int c = 0; c < 3; c++) out = tex_write(out, reference[0][c], piy[c], pix[c]);
  mvf_free(&mv);
  dwt_done(&dwt);
  arena_free(&A);
  return 0;
}

/* --------------------------------------------- small kernels for unit tests */

/* In-place 2-D transforms on a dense y*x int16 image (row stride = stride). */
int orc_dwt53(int16_t *img, int stride, int y, int x, int levels, int synth) {
  short **rows = malloc(sizeof(short *) * (size_t)(y > 0 ? y : 1));
  for (int j = 0; j < y; j++) rows[j] = img + (size_t)j * stride;
  dwt_t d;
  dwt_init(&d, 0);
  if (synth)
    dwt_synthesize(&d, rows, y, x, levels);
  else
    dwt_analyze(&d, rows, y, x, levels);
  dwt_done(&d);
  free(rows);
  return 0;
}

/* texture::alloc + read + fill_border as seen through data[y][x] for
 * y in [-b_alloc, y_alloc+b_alloc), x in [-b_alloc, x_alloc+b_alloc); used to
 * test border emulation.  out has (y_alloc+2b)*(x_alloc+2b) entries. */
int orc_bordered_view(const uint8_t *luma, int y_img, int x_img, int y_alloc, int x_alloc,
                      int b_alloc, int b_fill, int16_t *out) {
  arena_t A;
  if (arena_init(&A, tex_bytes(y_alloc, x_alloc, b_alloc) + (1 << 20))) return -1;
  short **t = tex_alloc(&A, y_alloc, x_alloc, b_alloc);
  for (int y = 0; y < y_alloc; y++)
    for (int x = 0; x < x_alloc; x++) t[y][x] = 0;
  tex_read(luma, t, y_img, x_img);
  tex_fill_border(t, y_img, x_img, b_fill);
  for (int y = -b_alloc; y < y_alloc + b_alloc; y++)
    for (int x = -b_alloc; x < x_alloc + b_alloc; x++) *out++ = t[y][x];
  arena_free(&A);
  return 0;
}

/* ------------------------------------------------ motion-field (de)correlation */

/* bidirectional_motion_decorrelate.cpp:25-52 (kernel), :195-215 (field loop).
 * analyze != 0: NEXT -= PREV; else NEXT += PREV.  Fields are [2][2][by][bx] shorts in file order. */
int orc_bidirectional_motion(int analyze, const int16_t *in, int fields, int by, int bx, int16_t *out) {
  const long plane = (long)by * bx;
  for (int i = 0; i < fields; i++) {
    const int16_t *f = in + (long)i * 4 * plane;
    int16_t *o = out + (long)i * 4 * plane;
    memcpy(o, f, sizeof(int16_t) * 4 * plane);
    for (long k = 0; k < 2 * plane; k++) /* X then Y of NEXT against X then Y of PREV */
      o[2 * plane + k] = (int16_t)(analyze ? f[2 * plane + k] - f[k] : f[2 * plane + k] + f[k]);
  }
  return 0;
}

/* interlevel_motion_decorrelate.cpp:32-69 (kernel), :250-297 (reader loop): one reference field
 * per iteration (fread past the end leaves the buffer as it was; a missing file is /dev/zero),
 * then up to two fields of the predicted (analyze) or residue stream, stopping at its end.
 * ref == NULL: zeros.  Returns the number of fields written. */
int orc_interlevel_motion(int analyze, const int16_t *in, int n_in, const int16_t *ref, int n_ref,
                          int fields_in_predicted, int by, int bx, int16_t *out) {
  const long fsz = 4L * by * bx;
  int16_t *reference = calloc((size_t)(fsz > 0 ? fsz : 1), sizeof(int16_t));
  int rd = 0, wr = 0, rr = 0;
  for (int i = 0; i < fields_in_predicted; i++) {
    if (ref && rr < n_ref) memcpy(reference, ref + (long)rr++ * fsz, sizeof(int16_t) * fsz);
    for (int p = 0; p < 2; p++) {
      if (rd >= n_in) break; /* feof after the failed read */
      const int16_t *f = in + (long)rd++ * fsz;
      int16_t *o = out + (long)wr++ * fsz;
      for (long k = 0; k < fsz; k++)
        o[k] = (int16_t)(analyze ? f[k] - reference[k] / 2 : f[k] + reference[k] / 2);
    }
  }
  free(reference);
  return wr;
}
