// kernels_img.cu -- frame load/store and border emulation.
//
// Reference: texture.cpp:122-144 (u8 <-> i16 row I/O, truncating store),
// texture.cpp:55-113 (fill_border, 8 regions), texture.cpp:34-46 (alloc).
#include "kernels.cuh"

#include <stdlib.h>

#define COUNT(L) (++*(L).counter)

ProfScope::ProfScope(const Launch &l, int cls) : L(l), idx(-1) {
  Profiler *p = L.prof;
  if (!p || !p->enabled) return;
  if (p->n == p->cap) {
    int ncap = p->cap ? p->cap * 2 : 4096;
    p->recs = (Profiler::Rec *)realloc(p->recs, sizeof(Profiler::Rec) * ncap);
    for (int i = p->cap; i < ncap; i++) {
      cudaEventCreate(&p->recs[i].a);
      cudaEventCreate(&p->recs[i].b);
    }
    p->cap = ncap;
  }
  idx = p->n++;
  p->recs[idx].cls = cls;
  cudaEventRecord(p->recs[idx].a, L.stream);
}

ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(L.prof->recs[idx].b, L.stream);
}

__global__ void k_load_u8(Plane dst, int slot0, const uint8_t *__restrict__ src,
                          long long frame_stride, long long comp_off, int f0, int fstep, int h,
                          int w) {
  int s = blockIdx.z;
  const uint8_t *f = src + (long long)(f0 + s * fstep) * frame_stride + comp_off;
  for (int y = blockIdx.y; y < h; y += gridDim.y) {
    short *row = dst.row(slot0 + s, y);
    const uint8_t *srow = f + (long long)y * w;
    // 8 samples per thread when both rows allow 64-bit loads / 128-bit stores
    const bool vec = ((((uintptr_t)row) & 15) | (((uintptr_t)srow) & 7)) == 0;
    const int nv = vec ? (w >> 3) : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
      const uint2 b = reinterpret_cast<const uint2 *>(srow)[i];
      uint4 o;
      o.x = __byte_perm(b.x, 0, 0x4140);
      o.y = __byte_perm(b.x, 0, 0x4342);
      o.z = __byte_perm(b.y, 0, 0x4140);
      o.w = __byte_perm(b.y, 0, 0x4342);
      reinterpret_cast<uint4 *>(row)[i] = o;
    }
    for (int x = (nv << 3) + blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x)
      row[x] = srow[x];
  }
}

void launch_load_u8(const Launch &L, Plane dst, int slot0, int nslots, const uint8_t *src,
                    long long frame_stride, long long comp_off, int f0, int fstep, int h, int w) {
  if (nslots <= 0 || h <= 0 || w <= 0) return;
  dim3 grid((w / 8 + 255) / 256 > 0 ? (w / 8 + 255) / 256 : 1, h < 1024 ? h : 1024, nslots);
  ProfScope ps_(L, KC_IMG);
  k_load_u8<<<grid, 256, 0, L.stream>>>(dst, slot0, src, frame_stride, comp_off, f0, fstep, h, w);
  COUNT(L);
}

__global__ void k_store_u8(Plane src, int slot0, uint8_t *__restrict__ dst, long long frame_stride,
                           long long comp_off, int f0, int fstep, int h, int w) {
  int s = blockIdx.z;
  uint8_t *f = dst + (long long)(f0 + s * fstep) * frame_stride + comp_off;
  for (int y = blockIdx.y; y < h; y += gridDim.y) {
    const short *row = src.row(slot0 + s, y);
    uint8_t *drow = f + (long long)y * w;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x)
      drow[x] = (uint8_t)row[x];  // texture.cpp:139-141: truncation mod 256, no clamp
  }
}

void launch_store_u8(const Launch &L, Plane src, int slot0, int nslots, uint8_t *dst,
                     long long frame_stride, long long comp_off, int f0, int fstep, int h, int w) {
  if (nslots <= 0 || h <= 0 || w <= 0) return;
  dim3 grid((w + 255) / 256, h < 1024 ? h : 1024, nslots);
  ProfScope ps_(L, KC_IMG);
  k_store_u8<<<grid, 256, 0, L.stream>>>(src, slot0, dst, frame_stride, comp_off, f0, fstep, h, w);
  COUNT(L);
}

// glibc chunk size field (chunk | PREV_INUSE) as the 4 shorts before each row.
__global__ void k_size_fields(Plane p, int slot0, int rows) {
  int s = blockIdx.y;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  unsigned long long sz = ((unsigned long long)p.S * 2ull) | 1ull;
  short *u = p.base + (long long)(slot0 + s) * p.slot_stride + (long long)r * p.S;
  u[-4] = (short)(sz & 0xffff);
  u[-3] = (short)((sz >> 16) & 0xffff);
  u[-2] = (short)((sz >> 32) & 0xffff);
  u[-1] = (short)((sz >> 48) & 0xffff);
}

void launch_size_fields(const Launch &L, Plane p, int slot0, int nslots, int rows) {
  if (nslots <= 0) return;
  dim3 grid((rows + 127) / 128, nslots);
  ProfScope ps_(L, KC_IMG);
  k_size_fields<<<grid, 128, 0, L.stream>>>(p, slot0, rows);
  COUNT(L);
}

// One fill_border region per launch: destinations of different regions can
// alias through the row-pointer bug, so the source order is preserved by
// stream order.  Within a region sources are interior cells and never alias a
// destination.
__global__ void k_fill_region(Plane p, int slot0, int region, int Y, int X, int b) {
  int s = slot0 + blockIdx.z;
  // destination rectangle and source rule per region (texture.cpp:57-112)
  int y_lo, y_hi, x_lo, x_hi;
  switch (region) {
    case 1: y_lo = -b; y_hi = 0; x_lo = -b; x_hi = 0; break;
    case 2: y_lo = -b; y_hi = 0; x_lo = 0; x_hi = X; break;
    case 3: y_lo = -b; y_hi = 0; x_lo = X; x_hi = X + b; break;
    case 4: y_lo = 0; y_hi = Y; x_lo = -b; x_hi = 0; break;
    case 5: y_lo = 0; y_hi = Y; x_lo = X; x_hi = X + b; break;
    case 6: y_lo = Y; y_hi = Y + b; x_lo = -b; x_hi = 0; break;
    case 7: y_lo = Y; y_hi = Y + b; x_lo = 0; x_hi = X; break;
    default: y_lo = Y; y_hi = Y + b; x_lo = X; x_hi = X + b; break;
  }
  for (int y = y_lo + blockIdx.y; y < y_hi; y += gridDim.y) {
    int sy = y < 0 ? 0 : (y >= Y ? Y - 1 : y);
    const short *srow = p.row(s, sy);
    short *drow = p.row(s, y);
    for (int x = x_lo + blockIdx.x * blockDim.x + threadIdx.x; x < x_hi;
         x += gridDim.x * blockDim.x) {
      int sx = x < 0 ? 0 : (x >= X ? X - 1 : x);
      if (region == 6) sx = X - 1;  // texture.cpp:95: bottom-left takes the bottom-right pixel
      drow[x] = srow[sx];
    }
  }
}

void launch_fill_border(const Launch &L, Plane p, int slot0, int nslots, int Y, int X, int b) {
  if (nslots <= 0 || b <= 0) return;
  for (int region = 1; region <= 8; region++) {
    int h = (region >= 4 && region <= 5) ? Y : b;
    int w = (region == 2 || region == 7) ? X : b;
    dim3 grid((w + 127) / 128, h < 512 ? h : 512, nslots);
    ProfScope ps_(L, KC_IMG);
    k_fill_region<<<grid, 128, 0, L.stream>>>(p, slot0, region, Y, X, b);
    COUNT(L);
  }
}

// texture::fill_border(data, Y, X, b) of a plain bordered plane (every row pointer shifted: no
// aliasing) whose interior would hold the luma of frame f0 + z: the ring straight from the frame's
// bytes, one launch (edge replication; the bottom-left corner takes the bottom-RIGHT pixel,
// texture.cpp:92-97).  The interior itself is not touched.
__global__ void __launch_bounds__(256) k_ring_u8(Plane p, int slot0, const uint8_t *__restrict__ src,
                                                  long long frame_stride, int f0, int Y, int X, int b) {
  const int slot = slot0 + blockIdx.z;
  const uint8_t *frame = src + (long long)(f0 + blockIdx.z) * frame_stride;
  const int W = X + 2 * b, nA = 2 * b * W, nB = 2 * b * Y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nA + nB; i += gridDim.x * blockDim.x) {
    int y, x;
    if (i < nA) {
      const int r = i / W;
      x = i - r * W - b;
      y = r < b ? r - b : Y + (r - b);
    } else {
      const int j = i - nA;
      y = j / (2 * b);
      const int k = j - y * (2 * b);
      x = k < b ? k - b : X + (k - b);
    }
    const int sy = y < 0 ? 0 : (y >= Y ? Y - 1 : y);
    int sx = x < 0 ? 0 : (x >= X ? X - 1 : x);
    if (y >= Y && x < 0) sx = X - 1;
    p.row(slot, y)[x] = frame[(long long)sy * X + sx];
  }
}

// The same from the plane's own interior (planes whose interior is already loaded).  Only for planes without
// the row-pointer alias (compact planes: every row pointer shifted), where no ring cell is anybody's source.
__global__ void __launch_bounds__(256) k_ring_s16(Plane p, int slot0, int Y, int X, int b) {
  const int slot = slot0 + blockIdx.z;
  const int W = X + 2 * b, nA = 2 * b * W, nB = 2 * b * Y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nA + nB; i += gridDim.x * blockDim.x) {
    int y, x;
    if (i < nA) {
      const int r = i / W;
      x = i - r * W - b;
      y = r < b ? r - b : Y + (r - b);
    } else {
      const int j = i - nA;
      y = j / (2 * b);
      const int k = j - y * (2 * b);
      x = k < b ? k - b : X + (k - b);
    }
    const int sy = y < 0 ? 0 : (y >= Y ? Y - 1 : y);
    int sx = x < 0 ? 0 : (x >= X ? X - 1 : x);
    if (y >= Y && x < 0) sx = X - 1;
    p.row(slot, y)[x] = p.row(slot, sy)[sx];
  }
}

void launch_ring_s16(const Launch &L, Plane p, int slot0, int nslots, int Y, int X, int b) {
  if (nslots <= 0 || b <= 0) return;
  ProfScope ps_(L, KC_IMG);
  k_ring_s16<<<dim3(64, 1, nslots), 256, 0, L.stream>>>(p, slot0, Y, X, b);
  COUNT(L);
}

void launch_ring_u8(const Launch &L, Plane p, int slot0, int nslots, const uint8_t *src, long long frame_stride,
                    int f0, int Y, int X, int b) {
  if (nslots <= 0 || b <= 0) return;
  const long long cells = 2LL * b * (X + 2 * b) + 2LL * b * Y;
  int blocks = (int)((cells + 1023) / 1024);
  if (blocks > 64) blocks = 64;
  ProfScope ps_(L, KC_IMG);
  k_ring_u8<<<dim3(blocks, 1, nslots), 256, 0, L.stream>>>(p, slot0, src, frame_stride, f0, Y, X, b);
  COUNT(L);
}

// Copies the top-left h x w region of every slot between the planes and a dense
// snapshot buffer (row pitch `pitch` shorts, `snap_slot_stride` shorts per slot).
__global__ void k_region_copy(Plane p, int slot0, int h, int w, short *snap, long long snap_slot_stride,
                              int pitch, int to_snapshot) {
  const int s = blockIdx.z;
  for (int y = blockIdx.y; y < h; y += gridDim.y) {
    short *row = p.row(slot0 + s, y);
    short *srow = snap + (long long)s * snap_slot_stride + (long long)y * pitch;
    const bool vec = (((uintptr_t)row | (uintptr_t)srow) & 15) == 0;
    const int nv = vec ? (w >> 3) : 0;
    uint4 *a = reinterpret_cast<uint4 *>(to_snapshot ? srow : row);
    const uint4 *b = reinterpret_cast<const uint4 *>(to_snapshot ? row : srow);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) a[i] = b[i];
    for (int x = (nv << 3) + blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) {
      if (to_snapshot) srow[x] = row[x];
      else row[x] = srow[x];
    }
  }
}

void launch_region_copy(const Launch &L, Plane p, int slot0, int nslots, int h, int w, short *snap,
                        long long snap_slot_stride, int pitch, bool to_snapshot) {
  if (nslots <= 0 || h <= 0 || w <= 0) return;
  dim3 grid((w / 8 + 127) / 128 > 0 ? (w / 8 + 127) / 128 : 1, h < 2048 ? h : 2048, nslots);
  ProfScope ps_(L, KC_IMG);
  k_region_copy<<<grid, 128, 0, L.stream>>>(p, slot0, h, w, snap, snap_slot_stride, pitch, to_snapshot ? 1 : 0);
  COUNT(L);
}
