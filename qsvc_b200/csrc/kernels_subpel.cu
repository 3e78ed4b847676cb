// kernels_subpel.cu -- sub-pixel levels of the motion search without
// materialising the 2^a-times up-sampled int16 images.
//
// Reference semantics (motion_estimate.cpp:361-407): for l = 1..a the three
// images are synthesised one more 5/3 level in place (high bands = whatever the
// buffer holds there: zeros, plus the fill_border replicas that the un-shifted
// fill_border call left inside the first high-band rows/columns), vectors are
// doubled and clamped, and every block is refined by +-1 at block size bs<<l.
//
// Here:
//  * B0 = the level-0 buffer after the over-pixel descent lives in a compact
//    bordered int16 plane (interior + fill ring + zeros);
//  * V_l = up^l(u8(B0 interior)) are dense u8 planes made by k_upsample2x: the
//    exact zero-high-band synthesis for byte data (s[2i] = l[i],
//    s[2i+1] = (l[i] + l[i+1]) >> 1, last = l[last]; columns first, then rows);
//  * a block whose windows lie in the region where B_l == V_l (away from the
//    polluted top/left strips, inside the picture, all samples bytes) takes the
//    FAST path: windows are copied as aligned 32-bit words and the nine SADs per
//    direction are accumulated with __vsadu4 (VABSDIFF4.U8.ACC, 4 SAD-ops/instr);
//  * every other block is queued and handled by the EXACT path, which rebuilds
//    the int16 windows from B0 with the literal synthesis formulas (including
//    the high-band reads) and accumulates with __sad.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

#define COUNT(L) (++*(L).counter)

__constant__ int c_cand9[9][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}, {-1, 0},
                                  {1, 0},   {0, 1},  {0, -1}, {0, 0}};

// ------------------------------------------------------------ u8 planes

__global__ void __launch_bounds__(256) k_plane_to_u8(Plane src, int slot0, int Y, int X,
                                                     uint8_t *dst, long long dst_slot_stride,
                                                     int pitch, uint8_t *tile_bad, int tiles_x,
                                                     int tiles_per_slot) {
  // tile_bad[slot][y >> TS][x >> TS] != 0: that tile holds a sample outside [0,255]
  const int s = blockIdx.z;
  for (int y = blockIdx.y; y < Y; y += gridDim.y) {
    const short *row = src.row(slot0 + s, y);
    uint8_t *drow = dst + (long long)s * dst_slot_stride + (long long)y * pitch;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < X; x += gridDim.x * blockDim.x) {
      int v = row[x];
      if ((unsigned)v > 255u) tile_bad[(long long)s * tiles_per_slot + (y >> SUBPEL_TILE_SHIFT) * tiles_x + (x >> SUBPEL_TILE_SHIFT)] = 1;
      drow[x] = (uint8_t)v;
    }
  }
}

void launch_plane_to_u8(const Launch &L, Plane src, int slot0, int nslots, int Y, int X,
                        uint8_t *dst, long long dst_slot_stride, int pitch, uint8_t *tile_bad,
                        int tiles_x, int tiles_per_slot) {
  if (nslots <= 0) return;
  dim3 grid((X + 1023) / 1024, Y < 512 ? Y : 512, nslots);
  ProfScope ps_(L, KC_IMG);
  k_plane_to_u8<<<grid, 256, 0, L.stream>>>(src, slot0, Y, X, dst, dst_slot_stride, pitch, tile_bad,
                                            tiles_x, tiles_per_slot);
  COUNT(L);
}

// V_0 planes straight from the frames' luma (pyramids whose descent restores the picture exactly)
__global__ void __launch_bounds__(256) k_luma_to_plane(const uint8_t *__restrict__ src, long long frame_stride,
                                                       int Y, int X, uint8_t *__restrict__ dst,
                                                       long long dst_slot_stride, int pitch) {
  const uint8_t *f = src + (long long)blockIdx.z * frame_stride;
  uint8_t *d = dst + (long long)blockIdx.z * dst_slot_stride;
  for (int y = blockIdx.y; y < Y; y += gridDim.y) {
    const uint8_t *srow = f + (long long)y * X;
    uint8_t *drow = d + (long long)y * pitch;
    const bool vec = (((uintptr_t)srow | (uintptr_t)drow) & 15) == 0;
    const int nv = vec ? (X >> 4) : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x)
      reinterpret_cast<uint4 *>(drow)[i] = reinterpret_cast<const uint4 *>(srow)[i];
    for (int x = (nv << 4) + blockIdx.x * blockDim.x + threadIdx.x; x < X; x += gridDim.x * blockDim.x)
      drow[x] = srow[x];
  }
}

void launch_luma_to_plane(const Launch &L, const uint8_t *src, long long frame_stride, int nframes, int Y, int X,
                          uint8_t *dst, long long dst_slot_stride, int pitch) {
  if (nframes <= 0) return;
  dim3 grid((X / 16 + 255) / 256 > 0 ? (X / 16 + 255) / 256 : 1, Y < 512 ? Y : 512, nframes);
  ProfScope ps_(L, KC_IMG);
  k_luma_to_plane<<<grid, 256, 0, L.stream>>>(src, frame_stride, Y, X, dst, dst_slot_stride, pitch);
  COUNT(L);
}

// out (2n x 2m) = rows(cols(in (n x m))) of the zero-high-band 5/3 synthesis on bytes
// (5_3.cpp:81-94 with h = 0, dwt2d.cpp:139-172: columns first, then rows).
// One thread: 4 input pixels of input rows i and i+1 -> 8 output pixels of rows 2i, 2i+1.
__global__ void __launch_bounds__(256) k_upsample2x(const uint8_t *__restrict__ in, int n, int m,
                                                    int pitch_in, long long in_slot_stride,
                                                    uint8_t *__restrict__ out, int pitch_out,
                                                    long long out_slot_stride) {
  const int s = blockIdx.z;
  const uint8_t *src = in + (long long)s * in_slot_stride;
  uint8_t *dst = out + (long long)s * out_slot_stride;
  const int mw = (m + 3) >> 2;
  for (int i = blockIdx.y; i < n; i += gridDim.y) {
    const uint8_t *r0 = src + (long long)i * pitch_in;
    const uint8_t *r1 = src + (long long)(i + 1 < n ? i + 1 : i) * pitch_in;  // last odd row = last row
    uint8_t *o0 = dst + (long long)(2 * i) * pitch_out;
    uint8_t *o1 = o0 + pitch_out;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < mw; w += gridDim.x * blockDim.x) {
      const int j = w << 2;
      unsigned a = *reinterpret_cast<const unsigned *>(r0 + j);
      unsigned b = *reinterpret_cast<const unsigned *>(r1 + j);
      // next pixel to the right (replicated at the right edge: last odd column = last column)
      int jn = j + 4 < m ? j + 4 : m - 1;
      unsigned an = r0[jn], bn = r1[jn];
      if (j + 4 > m) {  // partial word at the right edge: replicate the last valid pixel
        int last = m - 1 - j;
        unsigned la = (a >> (8 * last)) & 0xff, lb = (b >> (8 * last)) & 0xff;
        for (int k = last + 1; k < 4; k++) {
          a = (a & ~(0xffu << (8 * k))) | (la << (8 * k));
          b = (b & ~(0xffu << (8 * k))) | (lb << (8 * k));
        }
        an = la;
        bn = lb;
      }
      unsigned t0 = a;                 // column pass, even output row
      unsigned t1 = __vhaddu4(a, b);   // column pass, odd output row: floor((a + b) / 2)
      unsigned t0n = an, t1n = (an + bn) >> 1;
      // row pass: out[2j] = t[j], out[2j+1] = floor((t[j] + t[j+1]) / 2)
      unsigned s0 = __funnelshift_r(t0, t0n, 8), s1 = __funnelshift_r(t1, t1n, 8);
      unsigned h0 = __vhaddu4(t0, s0), h1 = __vhaddu4(t1, s1);
      uint2 q0, q1;
      q0.x = __byte_perm(t0, h0, 0x5140);
      q0.y = __byte_perm(t0, h0, 0x7362);
      q1.x = __byte_perm(t1, h1, 0x5140);
      q1.y = __byte_perm(t1, h1, 0x7362);
      *reinterpret_cast<uint2 *>(o0 + 2 * j) = q0;
      *reinterpret_cast<uint2 *>(o1 + 2 * j) = q1;
    }
  }
}

void launch_upsample2x(const Launch &L, const uint8_t *in, int n, int m, int pitch_in,
                       long long in_slot_stride, uint8_t *out, int pitch_out,
                       long long out_slot_stride, int nslots) {
  if (nslots <= 0) return;
  int mw = (m + 3) >> 2;
  dim3 grid((mw + 255) / 256, n < 1024 ? n : 1024, nslots);
  ProfScope ps_(L, KC_IMG);
  k_upsample2x<<<grid, 256, 0, L.stream>>>(in, n, m, pitch_in, in_slot_stride, out, pitch_out,
                                           out_slot_stride);
  COUNT(L);
}

// NST consecutive x2 up-samplings in one pass over the input: the CTA stages a (TH+1) x (TW+1)
// input tile (one halo sample to the right and below, clamped at the picture edge = the
// "last odd sample = last sample" rule), runs every stage but the last into shared memory and
// streams the last one to global memory; intermediate planes are written only where a consumer
// exists (outs[k] != nullptr).  Same arithmetic as k_upsample2x per stage, so the results are
// identical to chaining it; the traffic is one read of the input and one write per wanted plane.
// Every stage works on PAIRS of output rows (2i, 2i+1) and 16-byte output groups: the two source
// rows i, i+1 are read once, the even row is the horizontal expansion of row i, the odd row that of
// their packed-byte average.
struct UpChainOut {
  uint8_t *p[3];          // plane of stage k + 1 (k = 0 .. NST-1), nullptr: not materialised
  int pitch[3];
  long long slot_stride[3];
};

// horizontal expansion of eight samples (two words) + the sample to their right
__device__ __forceinline__ uint4 up_expand(unsigned ax, unsigned ay, unsigned an) {
  const unsigned h0 = __vhaddu4(ax, __funnelshift_r(ax, ay, 8)), h1 = __vhaddu4(ay, __funnelshift_r(ay, an, 8));
  return make_uint4(__byte_perm(ax, h0, 0x5140), __byte_perm(ax, h0, 0x7362), __byte_perm(ay, h1, 0x5140),
                    __byte_perm(ay, h1, 0x7362));
}
// output rows 2i and 2i+1, columns 16g .. 16g+15, from source rows i and i+1 (row pitch sp, a multiple of 8)
__device__ __forceinline__ void up_pair16(const uint8_t *src, int sp, int i, int g, uint4 &even, uint4 &odd) {
  const uint8_t *r0 = src + i * sp + 8 * g;
  const uint2 a = *reinterpret_cast<const uint2 *>(r0), b = *reinterpret_cast<const uint2 *>(r0 + sp);
  const unsigned an = r0[8], bn = r0[sp + 8];
  even = up_expand(a.x, a.y, an);
  odd = up_expand(__vhaddu4(a.x, b.x), __vhaddu4(a.y, b.y), (an + bn) >> 1);
}

template <int NST>
__global__ void __launch_bounds__(256) k_upsample_chain(const uint8_t *__restrict__ in, int n, int m, int pitch_in,
                                                        long long in_slot_stride, UpChainOut o) {
  constexpr int TH = 64 >> NST, TW = 512 >> NST;  // input tile; the last stage emits 64 x 512
  constexpr int P0 = TW + 16, P1 = 2 * TW + 16, P2 = 4 * TW + 16;  // stage buffer pitches (bytes, % 16 == 0)
  __shared__ __align__(16) uint8_t s0[(TH + 2) * P0];
  __shared__ __align__(16) uint8_t s1[(2 * TH + 2) * P1];
  __shared__ __align__(16) uint8_t s2[NST == 3 ? (4 * TH + 2) * P2 : 16];
  const int slot = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
  const uint8_t *src = in + (long long)slot * in_slot_stride;
  {
    // input tile rows y0 .. y0 + TH (halo row), columns x0 .. x0 + TW + 15, clamped to the picture
    constexpr int G0 = P0 / 16;
    const bool al = ((((uintptr_t)src) | (unsigned)pitch_in | (unsigned)x0) & 15) == 0;
    for (int i = threadIdx.x; i < (TH + 1) * G0; i += 256) {
      const int r = i / G0, c = (i - r * G0) * 16;
      const int y = min(y0 + r, n - 1);
      const uint8_t *row = src + (long long)y * pitch_in;
      uint4 v;
      if (al && x0 + c + 16 <= m) {
        v = *reinterpret_cast<const uint4 *>(row + x0 + c);
      } else {
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          w[k] = 0;
#pragma unroll
          for (int b = 0; b < 4; b++) w[k] |= (unsigned)row[min(x0 + c + 4 * k + b, m - 1)] << (8 * b);
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
      *reinterpret_cast<uint4 *>(s0 + r * P0 + c) = v;
    }
  }
  __syncthreads();
  const uint8_t *cur = s0;
  int cp = P0;
#pragma unroll
  for (int k = 0; k < NST; k++) {
    const int srows = TH << k, scols = TW << k;           // this stage's SOURCE tile (without halo)
    const int gy0 = y0 << (k + 1), gx0 = x0 << (k + 1), gn = n << (k + 1), gm = m << (k + 1);
    uint8_t *g = o.p[k] ? o.p[k] + (long long)slot * o.slot_stride[k] + (long long)gy0 * o.pitch[k] + gx0 : nullptr;
    const bool al16 = (o.pitch[k] & 15) == 0 && (gm & 15) == 0;
    const int pk = o.pitch[k];
    if (k < NST - 1) {
      uint8_t *nxt = k == 0 ? s1 : s2;
      const int np = k == 0 ? P1 : P2;
      // source row pairs (i, i+1) for i = 0 .. srows - 1 give output rows 0 .. 2 srows - 1; the halo row
      // 2 srows of the output is the even row of source pair srows (whose odd row nobody reads)
      const int gr = scols / 8 + 1;  // 16-byte output groups per row, one more for the halo column
      for (int t = threadIdx.x; t < (srows + 1) * gr; t += 256) {
        const int i = t / gr, q = t - i * gr;
        uint4 ev, od;
        up_pair16(cur, cp, i, q, ev, od);
        *reinterpret_cast<uint4 *>(nxt + (2 * i) * np + 16 * q) = ev;
        *reinterpret_cast<uint4 *>(nxt + (2 * i + 1) * np + 16 * q) = od;
        if (g && i < srows && q < gr - 1 && gx0 + 16 * q < gm) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            if (gy0 + 2 * i + h >= gn) break;
            uint8_t *d = g + (long long)(2 * i + h) * pk + 16 * q;
            const uint4 v = h ? od : ev;
            if (al16) {
              *reinterpret_cast<uint4 *>(d) = v;
            } else {  // picture width a multiple of 8 only
              *reinterpret_cast<uint2 *>(d) = make_uint2(v.x, v.y);
              if (gx0 + 16 * q + 8 < gm) *reinterpret_cast<uint2 *>(d + 8) = make_uint2(v.z, v.w);
            }
          }
        }
      }
      __syncthreads();
      cur = nxt;
      cp = np;
    } else {
      // last stage: 32 source rows x 32 groups of 8 source bytes; 256 threads = 8 source rows x 32 groups,
      // each thread walks down four source rows: pointers advance by constants
      const int q = threadIdx.x & 31, ti = threadIdx.x >> 5;
      if (gx0 + 16 * q < gm) {
        uint8_t *d = g + (long long)(2 * ti) * pk + 16 * q;
        const long long dstep = 16LL * pk;
        const bool tail8 = gx0 + 16 * q + 8 < gm;
#pragma unroll
        for (int i = ti; i < srows; i += 8, d += dstep) {
          if (gy0 + 2 * i >= gn) break;
          uint4 ev, od;
          up_pair16(cur, cp, i, q, ev, od);
          if (al16) {
            *reinterpret_cast<uint4 *>(d) = ev;
            if (gy0 + 2 * i + 1 < gn) *reinterpret_cast<uint4 *>(d + pk) = od;
          } else {
            *reinterpret_cast<uint2 *>(d) = make_uint2(ev.x, ev.y);
            if (tail8) *reinterpret_cast<uint2 *>(d + 8) = make_uint2(ev.z, ev.w);
            if (gy0 + 2 * i + 1 < gn) {
              *reinterpret_cast<uint2 *>(d + pk) = make_uint2(od.x, od.y);
              if (tail8) *reinterpret_cast<uint2 *>(d + pk + 8) = make_uint2(od.z, od.w);
            }
          }
        }
      }
    }
  }
}

// in (n x m) -> stage planes; outs[k] = plane of size (n << (k+1)) x (m << (k+1)) or nullptr
// (the last one is mandatory).  m << 1 must be a multiple of 8.
void launch_upsample_chain(const Launch &L, const uint8_t *in, int n, int m, int pitch_in, long long in_slot_stride,
                           int nst, uint8_t *const outs[3], const int pitches[3], const long long strides[3],
                           int nslots) {
  if (nslots <= 0) return;
  UpChainOut o;
  for (int k = 0; k < 3; k++) {
    o.p[k] = k < nst ? outs[k] : nullptr;
    o.pitch[k] = k < nst ? pitches[k] : 0;
    o.slot_stride[k] = k < nst ? strides[k] : 0;
  }
  ProfScope ps_(L, KC_IMG);
  if (nst == 3) {
    dim3 grid((m + 63) / 64, (n + 7) / 8, nslots);
    k_upsample_chain<3><<<grid, 256, 0, L.stream>>>(in, n, m, pitch_in, in_slot_stride, o);
  } else {
    dim3 grid((m + 127) / 128, (n + 15) / 16, nslots);
    k_upsample_chain<2><<<grid, 256, 0, L.stream>>>(in, n, m, pitch_in, in_slot_stride, o);
  }
  COUNT(L);
}

// ------------------------------------------------------------- fast path

__device__ __forceinline__ void subpel_centre(const SubpelParams &q, int pair, int by, int bx,
                                              short c[4]) {
  const long long plane = (long long)q.BY * q.BX;
  const short *mvi = q.mv_in + (long long)pair * 4 * plane;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    short v = mvi[k * plane + (long long)by * q.BX + bx];
    v = (short)(v * 2);
    if (v > q.lim) v = (short)q.lim;
    if (v < -q.lim) v = (short)(-q.lim);
    c[k] = v;
  }
}

__device__ __forceinline__ void subpel_store(const SubpelParams &q, int pair, int by, int bx,
                                             const short c[4], const int *err /*[2][9]*/) {
  const long long plane = (long long)q.BY * q.BX;
  short *mvo = q.mv_out + (long long)pair * 4 * plane;
  for (int d = 0; d < 2; d++) {
    int best = 0, min_error = 0;
    for (int k = 0; k < 9; k++) {
      int e = err[d * 9 + k];
      if (k == 0 || e <= min_error) {
        min_error = e;
        best = k;
      }
    }
    int sgn = d ? -1 : 1;
    long long dst = (long long)by * q.BX + bx;
    mvo[(2 * d) * plane + dst] = (short)(c[2 * d] + sgn * c_cand9[best][1]);
    mvo[(2 * d + 1) * plane + dst] = (short)(c[2 * d + 1] + sgn * c_cand9[best][0]);
  }
}

// 0: fast byte path, 1: strip path (a window touches the polluted strips or leaves the
// picture), 2: exact generator (a level-0 pixel under a window is not a byte).
template <int W, int NT>
__device__ __forceinline__ int classify_block(const SubpelParams &q, int l, int r0s, int r1s, int ps,
                                              int py0, int px0, const int wy[2], const int wx[2]) {
  const int Yl = q.Y << l, Xl = q.X << l;
  bool fast = true;
#pragma unroll
  for (int d = 0; d < 2; d++)
    fast = fast && wy[d] >= q.clean && wx[d] >= q.clean && wy[d] + W + 2 <= Yl && wx[d] + W + 2 <= Xl;
  int bad = 0;
  if (q.check_tiles) {
    for (int img = 0; img < 3; img++) {
      const int slot = img == 0 ? r0s : (img == 1 ? r1s : ps);
      const int y0 = img == 2 ? py0 : wy[img], x0 = img == 2 ? px0 : wx[img];
      const int span = img == 2 ? W : W + 2;
      const int pyl = max(y0 >> l, 0), pyh = min(((y0 + span - 1) >> l) + 1, q.Y - 1);
      const int pxl = max(x0 >> l, 0), pxh = min(((x0 + span - 1) >> l) + 1, q.X - 1);
      if (pyl > pyh || pxl > pxh) continue;
      constexpr int TS = SUBPEL_TILE_SHIFT;
      const int ty0 = pyl >> TS, nty = (pyh >> TS) - ty0 + 1, tx0 = pxl >> TS, ntx = (pxh >> TS) - tx0 + 1;
      const uint8_t *map = q.tile_bad + (long long)slot * q.tiles_per_slot;
      for (int i = threadIdx.x; i < nty * ntx; i += NT)
        bad |= map[(ty0 + i / ntx) * q.tiles_x + tx0 + i % ntx];
    }
  }
  bad = __syncthreads_or(bad);
  if (q.debug & 1) fast = false;               // debug: no block takes the TMA kernel
  if ((q.debug & 2) && !fast) bad = 1;         // debug: the strip kernel is bypassed
  return bad ? 2 : (fast ? 0 : 1);
}

__device__ __forceinline__ void queue_block(const SubpelParams &q, int kind, int pair, int by, int bx) {
  if (threadIdx.x == 0) {
    int *cnt = kind == 2 ? q.bad_count : q.slow_count;
    int *list = kind == 2 ? q.bad_list : q.slow_list;
    list[atomicAdd(cnt, 1)] = (pair * q.BY + by) * q.BX + bx;
  }
}

// W = block size at this level (bs << l), W in {16, 32, 64}.  Thread t owns the
// 32-bit word column j = t % (W/4) of RPT consecutive block rows.
template <int W>
__global__ void __launch_bounds__((W / 4) * (W / 8)) k_subpel_fast(SubpelParams q) {
  constexpr int WPR = W / 4;       // P words per row
  constexpr int RPT = 8;           // rows per thread
  constexpr int NT = WPR * (W / RPT);
  constexpr int PS = WPR + 2;      // smem row strides in words (8*stride % 32 == 16)
  constexpr int RWORDS = WPR + 2;  // aligned words covering W + 2 bytes at any alignment
  constexpr int RS = RWORDS + ((RWORDS % 4 == 2) ? 0 : ((6 - RWORDS % 4) % 4));
  __shared__ unsigned sP[W * PS];
  __shared__ unsigned sR[2][(W + 2) * RS];
  __shared__ int s_err[NT / 32 > 0 ? NT / 32 : 1][18];
  __shared__ int s_fin[18];

  const int bx = blockIdx.x, by = blockIdx.y, pair = blockIdx.z;
  const int l = q.l;
  short c[4];
  subpel_centre(q, pair, by, bx, c);
  const int r0s = q.slots[3 * pair], r1s = q.slots[3 * pair + 1], ps = q.slots[3 * pair + 2];
  const int py0 = by * W, px0 = bx * W;
  const int wy[2] = {py0 + c[MV_PREV_Y] - 1, py0 + c[MV_NEXT_Y] - 1};
  const int wx[2] = {px0 + c[MV_PREV_X] - 1, px0 + c[MV_NEXT_X] - 1};
  {
    const int kind = classify_block<W, NT>(q, l, r0s, r1s, ps, py0, px0, wy, wx);
    if (kind) {
      queue_block(q, kind, pair, by, bx);
      return;
    }
  }

  const uint8_t *vP = q.v + (long long)ps * q.v_slot_stride;
  for (int i = threadIdx.x; i < W * WPR; i += NT) {
    int y = i / WPR, w = i % WPR;
    sP[y * PS + w] = *reinterpret_cast<const unsigned *>(vP + (long long)(py0 + y) * q.v_pitch + px0 + 4 * w);
  }
#pragma unroll
  for (int d = 0; d < 2; d++) {
    const uint8_t *vR = q.v + (long long)(d ? r1s : r0s) * q.v_slot_stride;
    const int xa = wx[d] & ~3;
    for (int i = threadIdx.x; i < (W + 2) * RWORDS; i += NT) {
      int y = i / RWORDS, w = i % RWORDS;
      sR[d][y * RS + w] =
          *reinterpret_cast<const unsigned *>(vR + (long long)(wy[d] + y) * q.v_pitch + xa + 4 * w);
    }
  }
  __syncthreads();

  const int j = threadIdx.x % WPR, g = threadIdx.x / WPR;
  unsigned p[RPT];
#pragma unroll
  for (int r = 0; r < RPT; r++) p[r] = sP[(g * RPT + r) * PS + j];
  constexpr int DY[9] = {-1, -1, 1, 1, -1, 1, 0, 0, 0};
  constexpr int DX[9] = {-1, 1, -1, 1, 0, 0, 1, -1, 0};
  unsigned acc[18];
#pragma unroll
  for (int k = 0; k < 18; k++) acc[k] = 0;
#pragma unroll
  for (int d = 0; d < 2; d++) {
    // window column of candidate dx for P byte 0 of word j: (wx & 3) + 1 + sgn*dx + 4j
    const int sgn = d ? -1 : 1;
    const int base = (wx[d] & 3) + 1;
    const int o_m = base - 1, o_0 = base, o_p = base + 1;  // window byte offsets for shift -1, 0, +1
    const unsigned *R = sR[d] + (g * RPT) * RS + j;
#pragma unroll
    for (int rr = 0; rr < RPT + 2; rr++) {
      const unsigned *rw = R + rr * RS;
      unsigned w0 = rw[0], w1 = rw[1], w2 = rw[2];
      // shifted words for x offsets -1, 0, +1 (window coordinates)
      unsigned sh[3];
      {
        int o = o_m;
        sh[0] = __funnelshift_r(o < 4 ? w0 : w1, o < 4 ? w1 : w2, 8 * (o & 3));
        o = o_0;
        sh[1] = __funnelshift_r(o < 4 ? w0 : w1, o < 4 ? w1 : w2, 8 * (o & 3));
        o = o_p;
        sh[2] = __funnelshift_r(o < 4 ? w0 : w1, o < 4 ? w1 : w2, 8 * (o & 3));
      }
      // window row rr pairs with P row r = rr - 1 - wdy for window shifts wdy in {-1,0,1}
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const int wdy = sgn * DY[k], wdx = sgn * DX[k];
        const int r = rr - 1 - wdy;
        if (r >= 0 && r < RPT) acc[d * 9 + k] = __vsadu4(p[r], sh[wdx + 1]) + acc[d * 9 + k];
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 18; k++) {
    unsigned v = __reduce_add_sync(0xffffffffu, acc[k]);
    if (lane == 0) s_err[warp][k] = (int)v;
  }
  __syncthreads();
  if (threadIdx.x < 18) {
    int e = 0;
    for (int w = 0; w < (NT + 31) / 32; w++) e += s_err[w][threadIdx.x];
    s_fin[threadIdx.x] = e;
  }
  __syncthreads();
  if (threadIdx.x == 0) subpel_store(q, pair, by, bx, c, s_fin);
}

__device__ int g_tma_timeouts = 0;
int subpel_tma_timeouts() {
  int v = -1;
  cudaMemcpyFromSymbol(&v, g_tma_timeouts, sizeof(int));
  return v;
}

// ---- fast path, TMA variant ----
// The predicted block and the two (W+2)-row windows are fetched by three
// cp.async.bulk.tensor.2d loads (UTMALDG) at element coordinates, so the windows
// arrive already aligned to their first column: the three horizontal shifts are
// byte offsets 0/1/2 of two adjacent words.  Half of the threads take PREV, the
// other half NEXT; each thread owns one 32-bit word column of 16 block rows and
// accumulates the nine window shifts with VABSDIFF4.U8.ACC.
__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int x, int y, void *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

// ---- packed-byte SAD core shared by the TMA kernel and the strip kernel ----
// 2*W threads per block: thread = (direction d, row group g of RPT = W/4 block rows, word column
// j of W/4).  A thread keeps its RPT block words in registers, walks the RPT + 2 window rows
// below them once and accumulates the nine window shifts with VABSDIFF4.U8.ACC.  S = byte
// offset of window column 0 inside word 0 of the row pointer (0 when the window was staged
// aligned); EX adds the sums of a second byte window (the "excess" of int16 samples outside
// [0,255], see k_subpel_strip) under the same shifts.
// Funnel shift right by 8 * k bits.  FM: as two integer multiplies by a power of two the compiler
// cannot see (IMAD.HI + IMAD on the FMA pipe) instead of one SHF on the ALU pipe, which the SAD
// instructions saturate (profiles/r2_pipe_probe.txt: VABSDIFF4 and SHF share a pipe, IMAD does not).
template <bool FM>
__device__ __forceinline__ unsigned fshr(unsigned lo, unsigned hi, int k, const unsigned *km) {
  if (!FM) return __funnelshift_r(lo, hi, 8 * k);
  unsigned t, r;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(lo), "r"(km[k - 1]));
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(hi), "r"(km[k - 1]), "r"(t));
  return r;
}

template <int S, bool FM>
__device__ __forceinline__ void shifted3(const unsigned *rw, unsigned sh[3], const unsigned *km) {
  const unsigned w0 = rw[0], w1 = rw[1];
  if (S == 0) {
    sh[0] = w0;
    sh[1] = fshr<FM>(w0, w1, 1, km);
    sh[2] = fshr<FM>(w0, w1, 2, km);
  } else if (S == 1) {
    sh[0] = fshr<FM>(w0, w1, 1, km);
    sh[1] = fshr<FM>(w0, w1, 2, km);
    sh[2] = fshr<FM>(w0, w1, 3, km);
  } else if (S == 2) {
    sh[0] = fshr<FM>(w0, w1, 2, km);
    sh[1] = fshr<FM>(w0, w1, 3, km);
    sh[2] = w1;
  } else {
    const unsigned w2 = rw[2];
    sh[0] = fshr<FM>(w0, w1, 3, km);
    sh[1] = w1;
    sh[2] = fshr<FM>(w1, w2, 1, km);
  }
}

// COLS adjacent word columns per thread: their window words overlap, so the loads (and the fixed
// cost per thread) are shared.
template <int W, int S, bool EX, bool FM = false, int COLS = 1>
__device__ __forceinline__ void sad_rows(const unsigned *P4, const unsigned *R4, const unsigned *E4, int rp,
                                         unsigned acc[9], const unsigned *km = nullptr) {
  constexpr int WPR = W / 4, RPT = W / 4;
  unsigned p[RPT][COLS];
#pragma unroll
  for (int r = 0; r < RPT; r++)
#pragma unroll
    for (int cc = 0; cc < COLS; cc++) p[r][cc] = P4[r * WPR + cc];
#pragma unroll
  for (int rr = 0; rr < RPT + 2; rr++) {
    unsigned sh[COLS][3], se[COLS][3];
#pragma unroll
    for (int cc = 0; cc < COLS; cc++) {
      shifted3<S, FM>(R4 + rr * rp + cc, sh[cc], km);
      if (EX) shifted3<S, FM>(E4 + rr * rp + cc, se[cc], km);
    }
#pragma unroll
    for (int wdy = -1; wdy <= 1; wdy++) {
      const int r = rr - 1 - wdy;  // block row paired with window row rr under vertical shift wdy
      if (r >= 0 && r < RPT) {
#pragma unroll
        for (int wdx = -1; wdx <= 1; wdx++) {
          const int a = (wdy + 1) * 3 + wdx + 1;
#pragma unroll
          for (int cc = 0; cc < COLS; cc++) {
            acc[a] = __vsadu4(p[r][cc], sh[cc][wdx + 1]) + acc[a];
            if (EX) acc[a] = __vsadu4(se[cc][wdx + 1], 0u) + acc[a];
          }
        }
      }
    }
  }
}

// Warp sums -> block sums per direction -> arg-min with the reference's `<=` rule -> vectors.
// s_err: [2 * W / 32][9]; one direction per warp (W >= 32).
template <int W, int COLS = 1>
__device__ __forceinline__ void sad_finish(const SubpelParams &q, const unsigned acc[9], int (*s_err)[9], int pair,
                                           int by, int bx, const short c[4]) {
  constexpr int WPD = W / 32 / COLS;  // warps per direction
  static_assert(WPD >= 1, "one direction per warp at least");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const unsigned v = __reduce_add_sync(0xffffffffu, acc[k]);
    if (lane == 0) s_err[warp][k] = (int)v;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    // candidate k of direction d tests the window shift sgn * (dy, dx)
    const int d = threadIdx.x, sgn = d ? -1 : 1;
    int best = 0, min_error = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const int slot = (sgn * c_cand9[k][0] + 1) * 3 + sgn * c_cand9[k][1] + 1;
      int e = 0;
#pragma unroll
      for (int w = 0; w < WPD; w++) e += s_err[d * WPD + w][slot];
      if (k == 0 || e <= min_error) {
        min_error = e;
        best = k;
      }
    }
    const long long plane = (long long)q.BY * q.BX, dst = (long long)by * q.BX + bx;
    short *mvo = q.mv_out + (long long)pair * 4 * plane;
    mvo[(2 * d) * plane + dst] = (short)(c[2 * d] + sgn * c_cand9[best][1]);
    mvo[(2 * d + 1) * plane + dst] = (short)(c[2 * d + 1] + sgn * c_cand9[best][0]);
  }
}

// ---- fast path, TMA ----
// The predicted block and the two (W+2)-row windows are fetched by three
// cp.async.bulk.tensor.2d loads (UTMALDG) on one mbarrier.  ALIGNED = 1: windows are fetched
// from the 16-byte aligned column below wx (W + 32 bytes wide) and the SAD loop absorbs the
// byte offset (four instantiations).  ALIGNED = 0 would start the box at wx itself (W + 16
// bytes): measured on B200, a u8 box at a column that is not a multiple of 16 raises
// "illegal instruction", so only ALIGNED = 1 is instantiated.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar) {
  unsigned done = 0;
  const unsigned addr = smem_u32(bar);
  unsigned long long t0 = 0;
  for (int spin = 0; !done; spin++) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr)
        : "memory");
    if (!done && (spin & 1023) == 1023) {  // never hang the GPU: give up after 50 ms
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 50000000ull) { atomicAdd(&g_tma_timeouts, 1); break; }
    }
  }
}

template <int W, int ALIGNED, int COLS>
__global__ void __launch_bounds__(2 * W / COLS) k_subpel_tma(SubpelParams q, const __grid_constant__ CUtensorMap tmP,
                                                             const __grid_constant__ CUtensorMap tmR) {
  constexpr bool FM = false;
  constexpr int WPR = W / 4, RPT = W / 4, NT = 2 * W / COLS, NTD = W / COLS, WPRC = WPR / COLS;
  constexpr int BOXW = ALIGNED ? W + 32 : W + 16;  // TMA landing pitch in bytes
  constexpr int TP = BOXW / 4;
  constexpr int TBYTES = BOXW * (W + 2);
  constexpr int TB = (TBYTES + 127) & ~127;
  __shared__ __align__(128) unsigned char sP[W * W];
  __shared__ __align__(128) unsigned char sT[2][TB];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ int s_err[NT / 32][9];
  __shared__ int s_blk[8];  // doubled + clamped centre vectors, R0 / R1 / P slots

  // block order: bx fastest, then the pairs of a group, then by, then the groups -- pair i's
  // PREV window and pair i-1's NEXT window come from the same even frame and now meet in L2
  const int bx = blockIdx.x, by = blockIdx.y / q.pair_group;
  const int pair = blockIdx.z * q.pair_group + blockIdx.y % q.pair_group;
  if (pair >= q.npairs) return;
  const int l = q.l, Yl = q.Y << l, Xl = q.X << l;
  const int py0 = by * W, px0 = bx * W;
  const int lane = threadIdx.x & 31;
  int blk[7] = {0, 0, 0, 0, 0, 0, 0};
  bool fast;
  auto geometry = [&](int wy[2], int wx[2]) {
    wy[0] = py0 + blk[MV_PREV_Y] - 1;
    wy[1] = py0 + blk[MV_NEXT_Y] - 1;
    wx[0] = px0 + blk[MV_PREV_X] - 1;
    wx[1] = px0 + blk[MV_NEXT_X] - 1;
    bool f = true;
#pragma unroll
    for (int d = 0; d < 2; d++)
      f = f && wy[d] >= q.clean && wx[d] >= q.clean && wy[d] + W + 2 <= Yl && wx[d] + W + 2 <= Xl;
    return f && !(q.debug & 1);
  };
  int wy[2], wx[2];
  if (threadIdx.x < 32) {
    // warp 0: one global load per lane (four vector components, three slots), shared by
    // shuffles; lane 0 starts the TMA loads before anybody else has looked at the block
    const long long plane = (long long)q.BY * q.BX;
    int v = 0;
    if (lane < 4) {
      short m = q.mv_in[(long long)pair * 4 * plane + lane * plane + (long long)by * q.BX + bx];
      m = (short)(m * 2);
      if (m > q.lim) m = (short)q.lim;
      if (m < -q.lim) m = (short)(-q.lim);
      v = m;
    } else if (lane < 7) {
      v = q.slots[3 * pair + lane - 4];
    }
#pragma unroll
    for (int k = 0; k < 7; k++) blk[k] = __shfl_sync(0xffffffffu, v, k);
    fast = geometry(wy, wx);
    if (lane < 7) s_blk[lane] = v;
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (fast) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)),
                     "r"(W * W + 2 * TBYTES)
                     : "memory");
        tma_load_2d(sP, &tmP, px0, blk[6] * q.v_rows_per_slot + py0, &bar);
        tma_load_2d(sT[0], &tmR, ALIGNED ? (wx[0] & ~15) : wx[0], blk[4] * q.v_rows_per_slot + wy[0], &bar);
        tma_load_2d(sT[1], &tmR, ALIGNED ? (wx[1] & ~15) : wx[1], blk[5] * q.v_rows_per_slot + wy[1], &bar);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) {
#pragma unroll
    for (int k = 0; k < 7; k++) blk[k] = s_blk[k];
    fast = geometry(wy, wx);
  }
  int bad = 0;
  if (q.check_tiles) {
    // level-0 tiles under the three images' footprints: at most TG x TG each (footprint <= 35 pixels)
    constexpr int TS = SUBPEL_TILE_SHIFT, TG = (35 >> TS) + 2;
    for (int i = threadIdx.x; i < 3 * TG * TG; i += NT) {
      const int img = i / (TG * TG), r = i - img * (TG * TG), ty = r / TG, tx = r - ty * TG;
      const int slot = img == 0 ? blk[4] : (img == 1 ? blk[5] : blk[6]);
      const int y0 = img == 0 ? wy[0] : (img == 1 ? wy[1] : py0), x0 = img == 0 ? wx[0] : (img == 1 ? wx[1] : px0);
      const int span = img == 2 ? W : W + 2;
      const int pyl = max(y0 >> l, 0), pyh = min(((y0 + span - 1) >> l) + 1, q.Y - 1);
      const int pxl = max(x0 >> l, 0), pxh = min(((x0 + span - 1) >> l) + 1, q.X - 1);
      const int tyy = (pyl >> TS) + ty, txx = (pxl >> TS) + tx;
      if (pyl <= pyh && pxl <= pxh && tyy <= (pyh >> TS) && txx <= (pxh >> TS))
        bad |= q.tile_bad[(long long)slot * q.tiles_per_slot + tyy * q.tiles_x + txx];
    }
    bad = __syncthreads_or(bad);
  }
  if ((q.debug & 2) && !fast) bad = 1;  // debug: the strip kernel is bypassed
  if (bad || !fast) {
    if (fast) mbar_wait(&bar);  // the loads in flight must land before the block retires
    queue_block(q, bad ? 2 : 1, pair, by, bx);
    return;
  }
  mbar_wait(&bar);
  const int d = threadIdx.x / NTD, t = threadIdx.x % NTD;  // whole warps per direction
  const int j = (t % WPRC) * COLS, g = t / WPRC;
  const unsigned *P4 = reinterpret_cast<const unsigned *>(sP) + (g * RPT) * WPR + j;
  const unsigned *R4 = reinterpret_cast<const unsigned *>(sT[d]) + (g * RPT) * TP + j;
  unsigned acc[9];
#pragma unroll
  for (int k = 0; k < 9; k++) acc[k] = 0;
  if (ALIGNED) {
    R4 += (wx[d] & 15) >> 2;
    switch (wx[d] & 3) {
      case 0: sad_rows<W, 0, false, FM, COLS>(P4, R4, nullptr, TP, acc, q.kmul); break;
      case 1: sad_rows<W, 1, false, FM, COLS>(P4, R4, nullptr, TP, acc, q.kmul); break;
      case 2: sad_rows<W, 2, false, FM, COLS>(P4, R4, nullptr, TP, acc, q.kmul); break;
      default: sad_rows<W, 3, false, FM, COLS>(P4, R4, nullptr, TP, acc, q.kmul); break;
    }
  } else {
    sad_rows<W, 0, false, FM, COLS>(P4, R4, nullptr, TP, acc, q.kmul);
  }
  short c[4] = {(short)blk[0], (short)blk[1], (short)blk[2], (short)blk[3]};
  sad_finish<W, COLS>(q, acc, s_err, pair, by, bx, c);
}

// ------------------------------------------------------------ exact path

// Flattened 2-D loop over h x w cells by 256 threads: one division per thread, not per cell.
#define FOR_CELLS(r, c, h, w)                                                                  \
  for (int i_ = threadIdx.x, r = i_ / (w), c = i_ - r * (w), dc_ = 256 % (w), dr_ = 256 / (w); \
       i_ < (h) * (w); i_ += 256, c += dc_, r += dr_, r += (c >= (w)), c -= (c >= (w)) ? (w) : 0)


struct B0View {
  Plane p;
  int Y, X, B, Bc;       // picture, fill border, compact border
  int Ya, Ba;            // alloc geometry of the reference's buffer (rows)
  unsigned long long size_field;
};

// Cell of the reference's level-0 buffer at logical (y, x): the compact plane where
// it is materialised, else zero (never-written fresh heap), except the malloc size
// field seen at x in [-4, 0) of the rows whose pointer is not shifted (A.3).
__device__ __forceinline__ int b0_cell(const B0View &v, int slot, int y, int x) {
  if (y >= -v.Bc && y < v.Y + v.Bc && x >= -v.Bc && x < v.X + v.Bc) return v.p.row(slot, y)[x];
  if (x < 0 && x >= -4 && y >= v.Ya - v.Ba && y < v.Ya + v.Ba)
    return (short)((v.size_field >> (16 * (x + 4))) & 0xffff);
  return 0;
}
// High-band cell (row >= Y or column >= X inside the level-1 domain): zero beyond the
// fill_border replicas.
__device__ __forceinline__ int b0_high(const B0View &v, int slot, int y, int x) {
  if (y >= v.Y + v.B || x >= v.X + v.B) return 0;
  return v.p.row(slot, y)[x];
}

// Column pass of the level-1 synthesis (5_3.cpp:81-94 on column xp of [0,2Y) x [0,2X)).
__device__ int t1_even(const B0View &v, int slot, int i, int xp) {
  int low = xp < v.X ? (int)v.p.row(slot, i)[xp] : b0_high(v, slot, i, xp);
  int h = (i == 0) ? tdiv2(b0_high(v, slot, v.Y, xp))
                   : tdiv4(b0_high(v, slot, v.Y + i, xp) + b0_high(v, slot, v.Y + i - 1, xp));
  return (short)(low - h);
}
__device__ int t1_cell(const B0View &v, int slot, int y, int xp) {
  int i = y >> 1;
  if (!(y & 1)) return t1_even(v, slot, i, xp);
  int e0 = t1_even(v, slot, i, xp);
  int h = b0_high(v, slot, v.Y + i, xp);
  if (i < v.Y - 1) return (short)(h + tdiv2(e0 + t1_even(v, slot, i + 1, xp)));
  return (short)(h + e0);
}
// Level-1 image cell inside [0,2Y) x [0,2X) (row pass on top of the column pass).
__device__ int b1_even(const B0View &v, int slot, int y, int j) {
  int h = (j == 0) ? tdiv2(t1_cell(v, slot, y, v.X))
                   : tdiv4(t1_cell(v, slot, y, v.X + j) + t1_cell(v, slot, y, v.X + j - 1));
  return (short)(t1_cell(v, slot, y, j) - h);
}
__device__ int b1_inside(const B0View &v, int slot, int y, int x) {
  int j = x >> 1;
  if (y >= 2 * v.B + 2 && x >= 2 * v.B + 2) {
    // every high-band sample this cell depends on is zero: plain bilinear x2
    // (5_3.cpp:81-94 with h = 0; columns first, then rows)
    const int i = y >> 1;
    const short *r0 = v.p.row(slot, i);
    const short *r1 = v.p.row(slot, i + 1 < v.Y ? i + 1 : i);
    const int jn = j + 1 < v.X ? j + 1 : j;
    int t0 = (y & 1) ? (short)(tdiv2(r0[j] + r1[j])) : r0[j];
    if (!(x & 1)) return t0;
    int t1 = (y & 1) ? (short)(tdiv2(r0[jn] + r1[jn])) : r0[jn];
    return (short)(tdiv2(t0 + t1));
  }
  if (!(x & 1)) return b1_even(v, slot, y, j);
  int e0 = b1_even(v, slot, y, j);
  int h = t1_cell(v, slot, y, v.X + j);
  if (j < v.X - 1) return (short)(h + tdiv2(e0 + b1_even(v, slot, y, j + 1)));
  return (short)(h + e0);
}
__device__ __forceinline__ int b1_cell(const B0View &v, int slot, int y, int x) {
  if (y >= 0 && y < 2 * v.Y && x >= 0 && x < 2 * v.X) return b1_inside(v, slot, y, x);
  return b0_cell(v, slot, y, x);
}

// Level-1 window (h x w at (y0, x0)) into dst.  The four level-0 sub-windows it depends
// on -- low/low, low/high (columns >= X), high/low (rows >= Y), high/high -- are staged in
// shared memory first (high-band cells are zero beyond the fill_border replicas), then the
// column pass T (5_3.cpp:81-94 on columns) and the row pass run from shared memory.
__device__ void gen_level1(const B0View &v, int slot, int y0, int x0, int h, int w, short *dst,
                           short *TL, short *TH, short *stage, int nthreads) {
  const int Y1 = 2 * v.Y, X1 = 2 * v.X;
  const int ya = max(y0, 0), yb = min(y0 + h, Y1);
  const int xa = max(x0, 0), xb = min(x0 + w, X1);
  int j0 = 0, nj = 0, hj0 = 0, nh = 0;
  if (ya < yb && xa < xb) {
    j0 = xa >> 1;
    const int j1 = min(((xb - 1) >> 1) + 1, v.X - 1);
    nj = j1 - j0 + 1;
    hj0 = max(j0 - 1, 0);
    nh = j1 - hj0 + 1;
    const int ilo = max((ya >> 1) - 1, 0), ihi = min(((yb - 1) >> 1) + 1, v.Y - 1);
    const int ni = ihi - ilo + 1;
    short *SLL = stage, *SHL = SLL + ni * nh, *SLH = SHL + ni * nh, *SHH = SLH + ni * nh;
    FOR_CELLS(r, cc, ni, nh) {
      const int yy = ilo + r, xx = hj0 + cc, i = r * nh + cc;
      SLL[i] = v.p.row(slot, yy)[xx];
      SHL[i] = (short)b0_high(v, slot, yy, v.X + xx);
      SLH[i] = (short)b0_high(v, slot, v.Y + yy, xx);
      SHH[i] = (short)b0_high(v, slot, v.Y + yy, v.X + xx);
    }
    __syncthreads();
    const int rows = yb - ya;
    // column pass for the low columns (TL) and the high columns (TH)
    FOR_CELLS(r2, cc, rows * 2, nh) {
      const int which = r2 >= rows;  // 0: low columns, 1: high columns
      const int r = which ? r2 - rows : r2;
      const short *lo = (which ? SHL : SLL) + cc, *hi = (which ? SHH : SLH) + cc;
      const int y = ya + r, ii = y >> 1, li = ii - ilo;
      auto te = [&](int q) -> int {  // even sample 2*(ilo+q) of the column
        const int gi = ilo + q;
        const int hh = gi == 0 ? tdiv2(hi[0]) : tdiv4(hi[q * nh] + hi[(q - 1) * nh]);
        return (short)(lo[q * nh] - hh);
      };
      int val;
      if (!(y & 1)) {
        val = te(li);
      } else {
        const int e0 = te(li);
        val = (ii < v.Y - 1) ? (short)(hi[li * nh] + tdiv2(e0 + te(li + 1))) : (short)(hi[li * nh] + e0);
      }
      if (which) TH[r * nh + cc] = (short)val;
      else if (cc >= j0 - hj0) TL[r * nj + cc - (j0 - hj0)] = (short)val;
    }
  }
  __syncthreads();
  FOR_CELLS(ry, rx, h, w) {
    const int y = y0 + ry, x = x0 + rx, i = ry * w + rx;
    int val;
    if (y >= 0 && y < Y1 && x >= 0 && x < X1) {
      const short *tl = TL + (y - ya) * nj - j0;
      const short *th = TH + (y - ya) * nh - hj0;
      auto even = [&](int jj) -> int {
        int hh = (jj == 0) ? tdiv2(th[0]) : tdiv4(th[jj] + th[jj - 1]);
        return (short)(tl[jj] - hh);
      };
      const int j = x >> 1;
      if (!(x & 1)) {
        val = even(j);
      } else {
        int e0 = even(j);
        val = (j < v.X - 1) ? (short)(th[j] + tdiv2(e0 + even(j + 1))) : (short)(th[j] + e0);
      }
    } else {
      val = b0_cell(v, slot, y, x);
    }
    dst[i] = (short)val;
  }
}

// Fills dst (H x WW, row stride WW) with the level-l window whose top-left is (y0, x0).
// Level 2 is built from a level-1 window staged in `tmp` (high bands of the second
// synthesis are zero because B <= min(X, Y): checked on the host), separably: column
// pass into `vbuf` (H x tw), then row pass.
template <int H, int WW>
__device__ void gen_window(const B0View &v, int slot, int l, int y0, int x0, short *dst, short *tmp,
                           short *TL, short *TH, short *stage, int nthreads) {
  if (l == 1) {
    gen_level1(v, slot, y0, x0, H, WW, dst, TL, TH, stage, nthreads);
    return;
  }
  // level-1 window covering rows [y0>>1, ((y0+H-1)>>1)+1], same for columns (floor division)
  const int ty0 = y0 >> 1, tx0 = x0 >> 1;
  const int th = ((y0 + H - 1) >> 1) - ty0 + 2, tw = ((x0 + WW - 1) >> 1) - tx0 + 2;
  const int Y2 = 4 * v.Y, X2 = 4 * v.X;
  gen_level1(v, slot, ty0, tx0, th, tw, tmp, TL, TH, stage, nthreads);
  __syncthreads();
  short *vbuf = TL;  // free again after gen_level1
  // column pass (zero high band): vbuf[yy][xp] = T(y0 + yy, tx0 + xp)
  FOR_CELLS(yy, xp, H, tw) {
    const int y = y0 + yy, i = yy * tw + xp;
    int val = 0;
    if (y >= 0 && y < Y2) {
      const int i0 = (y >> 1) - ty0;
      const int a = tmp[i0 * tw + xp];
      val = (!(y & 1) || y == Y2 - 1) ? a : (short)tdiv2(a + tmp[(i0 + 1) * tw + xp]);
    }
    vbuf[i] = (short)val;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * WW; i += nthreads) {
    const int yy = i / WW, xx = i - yy * WW;  // WW is a compile-time constant
    const int y = y0 + yy, x = x0 + xx;
    int val;
    if (y >= 0 && y < Y2 && x >= 0 && x < X2) {
      const short *r = vbuf + yy * tw + ((x >> 1) - tx0);
      val = (!(x & 1) || x == X2 - 1) ? (int)r[0] : (short)tdiv2(r[0] + r[1]);
    } else {
      val = b0_cell(v, slot, y, x);
    }
    dst[i] = (short)val;
  }
}

// SAD of the W x W block against the two (W+2) x (W+2) int16 windows for the nine window
// shifts per direction; 256 threads; result s_fin[direction * 9 + candidate].
template <int W>
__device__ __forceinline__ void sad_windows(const short *Ps, const short *Rs0, const short *Rs1,
                                            int (*s_part)[18], int *s_fin) {
  constexpr int RW = W + 2;
    // SAD: thread = (block row, segment of SEG pixels); accumulators indexed by window shift
    constexpr int SEG = W * W / 256, SPR = W / SEG;  // 16 px x 4 segments (W=64), 4 px x 8 (W=32)
    unsigned acc[18];
#pragma unroll
    for (int k = 0; k < 18; k++) acc[k] = 0;
    {
      const int y = threadIdx.x / SPR, x0s = (threadIdx.x % SPR) * SEG;
      int p[SEG];
#pragma unroll
      for (int i = 0; i < SEG; i++) p[i] = Ps[y * W + x0s + i];
#pragma unroll
      for (int d = 0; d < 2; d++) {
        const short *R = (d ? Rs1 : Rs0) + y * RW + x0s;
#pragma unroll
        for (int wr = 0; wr < 3; wr++) {
          int vv[SEG + 2];
#pragma unroll
          for (int i = 0; i < SEG + 2; i++) vv[i] = R[wr * RW + i];
#pragma unroll
          for (int wc = 0; wc < 3; wc++)
#pragma unroll
            for (int i = 0; i < SEG; i++) acc[d * 9 + wr * 3 + wc] = __sad(p[i], vv[i + wc], acc[d * 9 + wr * 3 + wc]);
        }
      }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 18; k++) {
      unsigned sacc = __reduce_add_sync(0xffffffffu, acc[k]);
      if (lane == 0) s_part[warp][k] = (int)sacc;
    }
    __syncthreads();
    if (threadIdx.x < 18) {
      // candidate k of direction dd tests the window shift sgn * (dy, dx)
      constexpr int DY[9] = {-1, -1, 1, 1, -1, 1, 0, 0, 0};
      constexpr int DX[9] = {-1, 1, -1, 1, 0, 0, 1, -1, 0};
      const int dd = threadIdx.x / 9, k = threadIdx.x % 9;
      const int sgn = dd ? -1 : 1;
      const int slot_i = dd * 9 + (sgn * DY[k] + 1) * 3 + sgn * DX[k] + 1;
      int e = 0;
      for (int w = 0; w < 8; w++) e += s_part[w][slot_i];
      s_fin[threadIdx.x] = e;
    }
    __syncthreads();
}

// ---- strip path ----
// Sample (y, x) of the level-l image of `slot`: polluted strips (int16), byte plane, or --
// outside the picture -- the level-0 buffer cell the reference would read.
__device__ __forceinline__ int level_cell(const SubpelParams &q, const B0View &v, int slot, int y, int x) {
  const int Yl = q.Y << q.l, Xl = q.X << q.l;
  if (y >= 0 && y < Yl && x >= 0 && x < Xl) {
    if (y < q.clean) return q.strip_top[(long long)slot * q.strip_top_stride + (long long)y * Xl + x];
    if (x < q.clean)
      return q.strip_left[(long long)slot * q.strip_left_stride + (long long)(y - q.clean) * q.clean + x];
    return q.v[(long long)slot * q.v_slot_stride + (long long)y * q.v_pitch + x];
  }
  return b0_cell(v, slot, y, x);
}

// Blocks whose windows touch the polluted strips or leave the picture.  The predicted block is
// always bytes (predicted frames are never border-filled and the block has no non-byte tile), so
// for a window sample r of any value |p - r| = |p - clamp(r, 0, 255)| + excess(r) with
// excess(r) = r - 255 above, -r below, 0 inside.  The kernel stages the clamped windows and their
// excess as byte windows (column 0 word-aligned) and runs the packed-byte SAD core on both;
// a sample with excess > 255 (the malloc size field) sends the block to the exact generator.
template <int W>
__global__ void __launch_bounds__(2 * W) k_subpel_strip(SubpelParams q, B0View v) {
  constexpr int WPR = W / 4, RPT = W / 4, NT = 2 * W;
  constexpr int RW = W + 2, TP = (W + 4) / 4;  // window rows, words per staged row (odd pitch: no bank conflicts)
  __shared__ __align__(16) unsigned sP[W * WPR];
  __shared__ unsigned sC[2][RW * TP];
  __shared__ unsigned sE[2][RW * TP];
  __shared__ int s_err[NT / 32][9];
  const int Yl = q.Y << q.l, Xl = q.X << q.l;
  const int total = *q.slow_count;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int id = q.slow_list[item];
    const int bx = id % q.BX, by = (id / q.BX) % q.BY, pair = id / (q.BX * q.BY);
    short c[4];
    subpel_centre(q, pair, by, bx, c);
    const int rs[2] = {q.slots[3 * pair], q.slots[3 * pair + 1]};
    const int ps = q.slots[3 * pair + 2];
    const int py0 = by * W, px0 = bx * W;
    const int wy[2] = {py0 + c[MV_PREV_Y] - 1, py0 + c[MV_NEXT_Y] - 1};
    const int wx[2] = {px0 + c[MV_PREV_X] - 1, px0 + c[MV_NEXT_X] - 1};
    {  // predicted block: inside the picture, 16-byte aligned rows
      const uint8_t *vp = q.v + (long long)ps * q.v_slot_stride + (long long)py0 * q.v_pitch + px0;
      for (int i = threadIdx.x; i < W * (W / 16); i += NT) {
        const int y = i / (W / 16), w = i % (W / 16);
        reinterpret_cast<uint4 *>(sP)[i] = __ldg(reinterpret_cast<const uint4 *>(vp + (long long)y * q.v_pitch) + w);
      }
    }
    int flags = 0;  // 1: some excess is non-zero, 2: some excess does not fit a byte
    for (int i = threadIdx.x; i < 2 * RW * TP; i += NT) {
      const int d = i / (RW * TP), r = i - d * (RW * TP), yy = r / TP, w = r - yy * TP;
      const int slot = rs[d], y = wy[d] + yy, x = wx[d] + 4 * w;
      unsigned cw = 0, ew = 0;
      const bool row_in = y >= 0 && y < Yl;
      if (row_in && y >= q.clean && x >= q.clean && x + 3 < Xl) {
        // byte plane: unaligned word = two aligned words + funnel shift
        const uint8_t *vb = q.v + (long long)slot * q.v_slot_stride + (long long)y * q.v_pitch + x;
        const unsigned *a4 = reinterpret_cast<const unsigned *>(vb - ((uintptr_t)vb & 3));
        const unsigned sa = 8 * (unsigned)((uintptr_t)vb & 3);
        const unsigned lo = __ldg(a4);
        cw = sa ? __funnelshift_r(lo, __ldg(a4 + 1), sa) : lo;
      } else if (row_in && x >= 0 && (y < q.clean ? x + 3 < Xl : x + 3 < q.clean)) {
        // four samples of one int16 strip row (top strip: any column; left strip: columns below `clean`)
        const short *sp = y < q.clean
                              ? q.strip_top + (long long)slot * q.strip_top_stride + (long long)y * Xl + x
                              : q.strip_left + (long long)slot * q.strip_left_stride + (long long)(y - q.clean) * q.clean + x;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int sv = sp[k];
          const int cl = min(max(sv, 0), 255);
          const int ex = abs(sv - cl);
          if (4 * w + k < RW) flags |= (ex != 0) | ((ex > 255) << 1);
          cw |= (unsigned)cl << (8 * k);
          ew |= (unsigned)(ex & 255) << (8 * k);
        }
      } else {
        // One per-sample rule for every other word (words outside the level-l image, words that
        // straddle two regions), with the row-dependent parts resolved once: inside the
        // image the int16 strips or the byte plane; outside it the reference's level-0 buffer at
        // the same coordinates (b0_cell): the compact plane where it is materialised, the malloc
        // size field at x in [-4, 0) of the un-shifted rows, zeros (never-written heap) elsewhere.
        int s4[4];
        const short *srow = nullptr;  // int16 strip row: top strip (all x) or left strip (x < clean)
        const uint8_t *vrow = nullptr;
        if (row_in) {
          if (y < q.clean) {
            srow = q.strip_top + (long long)slot * q.strip_top_stride + (long long)y * Xl;
          } else {
            srow = q.strip_left + (long long)slot * q.strip_left_stride + (long long)(y - q.clean) * q.clean;
            vrow = q.v + (long long)slot * q.v_slot_stride + (long long)y * q.v_pitch;
          }
        }
        const short *prow = (y >= -v.Bc && y < v.Y + v.Bc) ? v.p.row(slot, y) : nullptr;
        const bool size_row = y >= v.Ya - v.Ba && y < v.Ya + v.Ba;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int xk = x + k;
          int sv = 0;
          if (row_in && (unsigned)xk < (unsigned)Xl) {
            sv = (vrow == nullptr || xk < q.clean) ? (int)srow[xk] : (int)vrow[xk];
          } else if (prow != nullptr && xk >= -v.Bc && xk < v.X + v.Bc) {
            sv = prow[xk];
          } else if (size_row && xk < 0 && xk >= -4) {
            sv = (short)((v.size_field >> (16 * (xk + 4))) & 0xffff);
          }
          s4[k] = sv;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int cl = min(max(s4[k], 0), 255);
          const int ex = abs(s4[k] - cl);
          // columns >= W + 2 of the staged row are never read by the SAD core
          if (4 * w + k < RW) flags |= (ex != 0) | ((ex > 255) << 1);
          cw |= (unsigned)cl << (8 * k);
          ew |= (unsigned)(ex & 255) << (8 * k);
        }
      }
      sC[d][r] = cw;
      sE[d][r] = ew;
    }
    const int any_ex = __syncthreads_or(flags);  // (a boolean) also publishes the staged windows
    if (any_ex && __syncthreads_or(flags & 2)) {
      queue_block(q, 2, pair, by, bx);
    } else {
      const int d = threadIdx.x / W, t = threadIdx.x % W;
      const int j = t % WPR, g = t / WPR;
      const unsigned *P4 = sP + (g * RPT) * WPR + j;
      const unsigned *R4 = sC[d] + (g * RPT) * TP + j, *E4 = sE[d] + (g * RPT) * TP + j;
      unsigned acc[9];
#pragma unroll
      for (int k = 0; k < 9; k++) acc[k] = 0;
      if (any_ex) sad_rows<W, 0, true>(P4, R4, E4, TP, acc);
      else sad_rows<W, 0, false>(P4, R4, E4, TP, acc);
      sad_finish<W>(q, acc, s_err, pair, by, bx, c);
    }
    __syncthreads();
  }
}

// Level-1 image cells of the rectangle [Y0, Y1) x [X0, X1) (X0 even) straight from the level-0
// buffer with the exact lifting formulas (5_3.cpp:81-94; dwt2d.cpp:139-172: columns, then rows),
// one 32 x 64 tile per CTA in four shared-memory phases: even rows of the column pass, all rows
// of the column pass, even columns of the row pass, output.  out(y, x) -> dst[(y - Y0) * dpitch + x - X0].
__global__ void __launch_bounds__(256) k_level1_tile(B0View v, int slot0, int Y0, int Y1, int X0, int X1, short *dst,
                                                     long long dst_slot_stride, int dpitch) {
  constexpr int TR = 32, TC = 64, NC = TC / 2 + 2;  // low columns j0 .. j0+NC-2, high columns j0-1 .. j0+NC-2
  __shared__ short TE[2][TR / 2 + 2][NC + 1];  // [low | high columns][even-row index][column]
  __shared__ short TT[2][TR][NC + 1];
  __shared__ short EE[TR][NC + 1];
  const int slot = slot0 + blockIdx.z;
  const int ya = Y0 + blockIdx.y * TR, yb = min(ya + TR, Y1);
  const int xa = X0 + blockIdx.x * TC, xb = min(xa + TC, X1);
  const int X = v.X, Y = v.Y;
  const int j0 = xa >> 1;               // column jj of the low set is j0 + jj, of the high set X + j0 - 1 + jj
  const int nj = ((xb - 1) >> 1) - j0 + 2;  // low columns needed: j0 .. (xb-1)/2 + 1
  const int i0 = ya >> 1, ni = ((yb - 1) >> 1) - i0 + 2;
  auto lowc = [&](int i, int xp) -> int {  // column xp of buffer row i < Y
    return xp < X ? (int)v.p.row(slot, i)[xp] : b0_high(v, slot, i, xp);
  };
  // phase A: even rows of the column pass
  // (the loops run over the padded index space so that every index is decoded with constant divisors)
  constexpr int NI = TR / 2 + 2, NJ = NC + 1;
  for (int t = threadIdx.x; t < 2 * NI * NJ; t += 256) {
    const int set = t / (NI * NJ), r = t - set * (NI * NJ), ii = r / NJ, jj = r - ii * NJ;
    if (ii >= ni || jj > nj) continue;
    const int i = i0 + ii;
    const int xp = set ? X + j0 - 1 + jj : j0 + jj;
    int val = 0;
    if (i < Y && xp >= 0 && xp < 2 * X && (set || xp < X)) {
      const int hh = i == 0 ? tdiv2(b0_high(v, slot, Y, xp))
                            : tdiv4(b0_high(v, slot, Y + i, xp) + b0_high(v, slot, Y + i - 1, xp));
      val = (short)(lowc(i, xp) - hh);
    }
    TE[set][ii][jj] = (short)val;
  }
  __syncthreads();
  // phase B: all rows of the column pass
  for (int t = threadIdx.x; t < 2 * TR * NJ; t += 256) {
    const int set = t / (TR * NJ), r = t - set * (TR * NJ), yy = r / NJ, jj = r - yy * NJ;
    if (yy >= yb - ya || jj > nj) continue;
    const int y = ya + yy, i = y >> 1, ii = i - i0;
    const int xp = set ? X + j0 - 1 + jj : j0 + jj;
    int val = TE[set][ii][jj];
    if ((y & 1) && xp >= 0 && xp < 2 * X) {
      const int h = b0_high(v, slot, Y + i, xp);
      val = (i < Y - 1) ? (short)(h + tdiv2(val + TE[set][ii + 1][jj])) : (short)(h + val);
    }
    TT[set][yy][jj] = (short)val;
  }
  __syncthreads();
  // phase C: even columns of the row pass, j = j0 + jj
  for (int t = threadIdx.x; t < TR * NJ; t += 256) {
    const int yy = t / NJ, jj = t - yy * NJ, j = j0 + jj;
    if (yy >= yb - ya || jj >= nj) continue;
    int val = 0;
    if (j < X) {
      // high-set index of column X + j is jj + 1, of X + j - 1 is jj
      const int hh = j == 0 ? tdiv2(TT[1][yy][jj + 1]) : tdiv4(TT[1][yy][jj + 1] + TT[1][yy][jj]);
      val = (short)(TT[0][yy][jj] - hh);
    }
    EE[yy][jj] = (short)val;
  }
  __syncthreads();
  short *d = dst + (long long)slot * dst_slot_stride;
  for (int t = threadIdx.x; t < TR * TC; t += 256) {
    const int yy = t / TC, xx = t - yy * TC, x = xa + xx, j = x >> 1, jj = j - j0;
    if (yy >= yb - ya || xx >= xb - xa) continue;
    int val = EE[yy][jj];
    if (x & 1) {
      const int h = TT[1][yy][jj + 1];
      val = (j < X - 1) ? (short)(h + tdiv2(val + EE[yy][jj + 1])) : (short)(h + val);
    }
    d[(long long)(ya - Y0 + yy) * dpitch + (x - X0)] = (short)val;
  }
}

// Level-2 strips from the level-1 image (level-1 strips where polluted, V_1 bytes elsewhere):
// zero-high-band synthesis, columns first, then rows.  One CTA row per output row of the region
// rows [Y0, Y1) x columns [0, W) of the level-2 image: the row's two source rows and the kind of
// storage they live in are resolved once per CTA, a thread produces one sample.
static constexpr int STRIP2_ROWS = 16;
// One thread per level-1 column j: it produces the level-2 samples 2j and 2j + 1 of each of the CTA's
// rows from the column-pass values T(j) and T(j + 1) (the second one is the neighbour's first).
__global__ void __launch_bounds__(256) k_strip2(int Y, int X, const short *top1, long long top1_stride,
                                                const short *left1, long long left1_stride, int clean1,
                                                const uint8_t *v1, long long v1_slot_stride, int v1_pitch,
                                                short *dst, long long dst_stride, int slot0, int Y0, int Y1, int W) {
  const int slot = slot0 + blockIdx.z;
  const int X1 = 2 * X, Y2 = 4 * Y, X2 = 4 * X;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (2 * j >= W) return;
  const bool has_next = 2 * j + 1 < W;             // W odd: the region's last column stands alone
  const bool avg_next = 2 * j + 1 != X2 - 1;       // last odd column of the image = last column
  const short *t1 = top1 + slot * top1_stride, *l1 = left1 + slot * left1_stride;
  const uint8_t *b1 = v1 + slot * v1_slot_stride;
  short *out = dst + slot * dst_stride;
  // level-1 sample (row ya, column xp): int16 top strip, int16 left strip, or a byte of V_1
  const bool c0 = j < clean1, c1 = j + 1 < clean1;  // columns inside the left strip
  for (int y = Y0 + blockIdx.y * STRIP2_ROWS; y < min(Y0 + (blockIdx.y + 1) * STRIP2_ROWS, Y1); y++) {
    const int ya = y >> 1;
    const bool vavg = (y & 1) && y != Y2 - 1;
    const bool topA = ya < clean1, topB = ya + 1 < clean1;
    const short *sA = topA ? t1 + (long long)ya * X1 : l1 + (long long)(ya - clean1) * clean1;
    const short *sB = topB ? t1 + (long long)(ya + 1) * X1 : l1 + (long long)(ya + 1 - clean1) * clean1;
    const uint8_t *bA = b1 + (long long)ya * v1_pitch, *bB = bA + v1_pitch;
    int a0 = (topA || c0) ? (int)sA[j] : (int)bA[j];
    if (vavg) a0 = (short)tdiv2(a0 + ((topB || c0) ? (int)sB[j] : (int)bB[j]));
    int o1 = a0;
    if (has_next && avg_next) {
      int a1 = (topA || c1) ? (int)sA[j + 1] : (int)bA[j + 1];
      if (vavg) a1 = (short)tdiv2(a1 + ((topB || c1) ? (int)sB[j + 1] : (int)bB[j + 1]));
      o1 = (short)tdiv2(a0 + a1);
    }
    short *d = out + (long long)(y - Y0) * W + 2 * j;
    d[0] = (short)a0;
    if (has_next) d[1] = (short)o1;
  }
}

static B0View make_b0view(const SubpelParams &q) {
  B0View v;
  v.p = q.b0;
  v.Y = q.Y;
  v.X = q.X;
  v.B = q.B;
  v.Bc = q.Bc;
  v.Ya = q.Ya;
  v.Ba = q.Ba;
  v.size_field = q.size_field;
  return v;
}

void launch_strips(const Launch &L, const SubpelParams &q, int level, int slot0, int nslots, short *top, short *left,
                   const short *top1, const short *left1, long long top1_stride, long long left1_stride,
                   int clean1, const uint8_t *v1, long long v1_slot_stride, int v1_pitch) {
  if (nslots <= 0) return;
  const int Yl = q.Y << level, Xl = q.X << level;
  if (level == 1) {
    const B0View v = make_b0view(q);
    {
      dim3 grid((Xl + 63) / 64, (q.clean + 31) / 32, nslots);
      ProfScope ps_(L, KC_SEARCH_EXACT);
      k_level1_tile<<<grid, 256, 0, L.stream>>>(v, slot0, 0, q.clean, 0, Xl, top, q.strip_top_stride, Xl);
      COUNT(L);
    }
    if (Yl > q.clean) {
      dim3 grid((q.clean + 63) / 64, (Yl - q.clean + 31) / 32, nslots);
      ProfScope ps_(L, KC_SEARCH_EXACT);
      k_level1_tile<<<grid, 256, 0, L.stream>>>(v, slot0, q.clean, Yl, 0, q.clean, left, q.strip_left_stride, q.clean);
      COUNT(L);
    }
    return;
  }
  auto run = [&](short *dst, long long dst_stride, int Y0, int Y1, int W) {
    if (Y1 <= Y0 || W <= 0) return;
    ProfScope ps_(L, KC_SEARCH_EXACT);
    k_strip2<<<dim3(((W + 1) / 2 + 255) / 256, (Y1 - Y0 + STRIP2_ROWS - 1) / STRIP2_ROWS, nslots), 256, 0, L.stream>>>(
        q.Y, q.X, top1, top1_stride, left1, left1_stride, clean1, v1, v1_slot_stride, v1_pitch, dst, dst_stride,
        slot0, Y0, Y1, W);
    COUNT(L);
  };
  run(top, q.strip_top_stride, 0, q.clean, Xl);
  run(left, q.strip_left_stride, q.clean, Yl, q.clean);
}

template <int W>
__global__ void __launch_bounds__(256) k_subpel_exact(SubpelParams q, B0View v) {
  constexpr int RW = W + 2;
  extern __shared__ short sm[];
  short *Ps = sm;
  short *Rs0 = Ps + W * W;
  short *Rs1 = Rs0 + RW * RW;
  short *tmp = Rs1 + RW * RW;  // (W/2 + 4)^2 level-1 staging
  short *TL = tmp + (W / 2 + 4) * (W / 2 + 4);
  short *TH = TL + (W + 2) * (W / 2 + 4);
  short *stage = TH + (W + 2) * (W / 2 + 4);  // 4 x (W/2 + 4) x (W/2 + 4) level-0 sub-windows
  __shared__ int s_part[8][18];
  __shared__ int s_fin[18];
  const int total = *q.bad_count;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int id = q.bad_list[item];
    const int bx = id % q.BX, by = (id / q.BX) % q.BY, pair = id / (q.BX * q.BY);
    short c[4];
    subpel_centre(q, pair, by, bx, c);
    const int r0s = q.slots[3 * pair], r1s = q.slots[3 * pair + 1], ps = q.slots[3 * pair + 2];
    const int py0 = by * W, px0 = bx * W;
    gen_window<W, W>(v, ps, q.l, py0, px0, Ps, tmp, TL, TH, stage, blockDim.x);
    __syncthreads();
    gen_window<RW, RW>(v, r0s, q.l, py0 + c[MV_PREV_Y] - 1, px0 + c[MV_PREV_X] - 1, Rs0, tmp, TL, TH, stage, blockDim.x);
    __syncthreads();
    gen_window<RW, RW>(v, r1s, q.l, py0 + c[MV_NEXT_Y] - 1, px0 + c[MV_NEXT_X] - 1, Rs1, tmp, TL, TH, stage, blockDim.x);
    __syncthreads();
    sad_windows<W>(Ps, Rs0, Rs1, s_part, s_fin);
    if (threadIdx.x == 0) subpel_store(q, pair, by, bx, c, s_fin);
    __syncthreads();
  }
}

template <int W>
static void launch_subpel_w(const Launch &L, const SubpelParams &q, int npairs) {
  dim3 grid(q.BX, q.BY, npairs);
  {
    ProfScope ps_(L, KC_SEARCH);
    SubpelParams qg = q;
    qg.npairs = npairs;
    qg.pair_group = npairs < 8 ? npairs : 8;
    const dim3 ggrid(q.BX, q.BY * qg.pair_group, (npairs + qg.pair_group - 1) / qg.pair_group);
    const CUtensorMap &tp = *reinterpret_cast<const CUtensorMap *>(q.tm_p);
    const CUtensorMap &tr = *reinterpret_cast<const CUtensorMap *>(q.tm_r);
    // two word columns per thread for 64 x 64 blocks (64 threads per block: half the fixed cost per
    // block, shared window loads); env QSVC_SUBPEL_COLS=1: one column per thread as for 32 x 32
    static const int cols2 = getenv("QSVC_SUBPEL_COLS") ? atoi(getenv("QSVC_SUBPEL_COLS")) : 1;  // measured: no gain (latency per block, not instructions)
    qg.kmul[0] = 1u << 24;  // >> 8 (the FMA-pipe funnel shift of fshr<true>: measured without effect, not instantiated)
    qg.kmul[1] = 1u << 16;
    qg.kmul[2] = 1u << 8;
    if (q.use_tma && W == 64 && cols2 == 2)
      k_subpel_tma<W, 1, (W == 64 ? 2 : 1)><<<ggrid, W, 0, L.stream>>>(qg, tp, tr);
    else if (q.use_tma)
      k_subpel_tma<W, 1, 1><<<ggrid, 2 * W, 0, L.stream>>>(qg, tp, tr);
    else
      k_subpel_fast<W><<<grid, (W / 4) * (W / 8), 0, L.stream>>>(q);
    COUNT(L);
  }
  B0View v = make_b0view(q);
  const int RW = W + 2, TW = W / 2 + 4;
  {
    ProfScope ps_(L, KC_SEARCH_EXACT);
    k_subpel_strip<W><<<148 * 10, 2 * W, 0, L.stream>>>(q, v);
    COUNT(L);
  }
  size_t smem = ((size_t)W * W + 2 * (size_t)RW * RW + (size_t)TW * TW + 2 * (size_t)RW * TW + 4 * (size_t)TW * TW) *
                sizeof(short);
  static size_t s_attr = 0;
  if (smem > 48 * 1024 && smem > s_attr) {
    cudaFuncSetAttribute(k_subpel_exact<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    s_attr = smem;
  }
  {
    ProfScope ps_(L, KC_SEARCH_EXACT);
    k_subpel_exact<W><<<148 * 4, 256, smem, L.stream>>>(q, v);
    COUNT(L);
  }
}

bool subpel_supported(int W) { return W == 32 || W == 64; }

void launch_subpel(const Launch &L, const SubpelParams &q, int W, int npairs) {
  if (npairs <= 0) return;
  if (W == 32)
    launch_subpel_w<32>(L, q, npairs);
  else
    launch_subpel_w<64>(L, q, npairs);
}

// Tensor maps over the V_l planes of all slots seen as one tall 2-D u8 tensor
// (pitch bytes wide, nslots * rows_per_slot rows): box W x W for the predicted block
// and box_w x (W+2) for the windows (W + 32 when fetched from a 16-byte aligned column, W + 16
// when fetched at the window's own column).  Encoded through the driver entry point so the
// library does not link against libcuda.
bool subpel_make_tensor_maps(const uint8_t *v, int pitch, long long total_rows, int W, int box_w, void *tm_p,
                             void *tm_r) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = (encode_fn)fn;
    else
      cudaGetLastError();
  }
  if (!encode) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)total_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch};
  cuuint32_t estr[2] = {1, 1};
  cuuint32_t boxp[2] = {(cuuint32_t)W, (cuuint32_t)W};
  cuuint32_t boxr[2] = {(cuuint32_t)box_w, (cuuint32_t)(W + 2)};
  CUresult a = encode((CUtensorMap *)tm_p, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)v, gdim, gstride, boxp, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult b = encode((CUtensorMap *)tm_r, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)v, gdim, gstride, boxr, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return a == CUDA_SUCCESS && b == CUDA_SUCCESS;
}
