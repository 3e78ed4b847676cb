// kernels_mc.cu -- motion-compensated predict / update lifting steps.
//
// Reference: decorrelate.cpp:69-189 (predict, with and without block_overlaping),
// :841-848 (clip), :920-929 / :1009-1022 / :1038-1066 (residue, re-bias,
// inverse), :799-816 / :940-953 (histograms); update.cpp:71-148 (update).
#include <cstdlib>

#include "kernels.cuh"

#define COUNT(L) (++*(L).counter)

// P[c][y][x] = (R0[c][y+mv0] + R1[c][y+mv1]) / 2 over the covered area, then the
// [0,255] clip of decorrelate.cpp:841-848 fused in.  The luma vector (1/2^a pel)
// is applied unchanged to the up-sampled chroma planes (decorrelate.cpp:93-96).
__global__ void __launch_bounds__(256) k_predict(PredictParams q) {
  const int c = blockIdx.z;
  const int cy = q.BY * q.bsa, cx = q.BX * q.bsa;
  const long long plane = (long long)q.BY * q.BX;
  const short *U0 = q.ref.row(q.r0_slot * 3 + c, 0);
  const short *U1 = q.ref.row(q.r1_slot * 3 + c, 0);
  for (int y = blockIdx.y; y < cy; y += gridDim.y) {
    short *prow = q.pred.row(c, y);
    const int by = y / q.bsa;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < cx; x += gridDim.x * blockDim.x) {
      const long long b = (long long)by * q.BX + x / q.bsa;
      int y0 = y + q.mv[MV_PREV_Y * plane + b], x0 = x + q.mv[MV_PREV_X * plane + b];
      int y1 = y + q.mv[MV_NEXT_Y * plane + b], x1 = x + q.mv[MV_NEXT_X * plane + b];
      int v = ((int)bordered_ref(U0, q.ref.S, q.Ya, q.Xa, q.ba, q.padh, y0, x0) +
               (int)bordered_ref(U1, q.ref.S, q.Ya, q.Xa, q.ba, q.padh, y1, x1)) / 2;
      v = (short)v;
      prow[x] = (short)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
}

void launch_predict(const Launch &L, const PredictParams &q) {
  int cy = q.BY * q.bsa, cx = q.BX * q.bsa;
  if (cy <= 0 || cx <= 0) return;
  dim3 grid((cx + 1023) / 1024, cy < 2048 ? cy : 2048, 3);
  ProfScope ps_(L, KC_PREDICT);
  k_predict<<<grid, 256, 0, L.stream>>>(q);
  COUNT(L);
}

// predict() with block_overlaping > 0 (decorrelate.cpp:84-88, 99-172): every block is
// predicted with a margin of `ova` samples, analysed `levels` levels on its own (in-place
// Mallat layout of an N x N block, N = bsa + 2*ova), and the inner (bsa >> l)^2 part of each
// sub-band is scattered into the picture-sized Mallat layout; the caller then synthesises the
// whole picture `levels` levels.  One CTA per (block, component); the block ping-pongs between
// two shared-memory buffers (rows: A -> B, columns: B -> A), so after each level it is back in A.
__global__ void __launch_bounds__(256) k_predict_obmc(PredictParams q, int ova, int levels) {
  extern __shared__ short obmc_sm[];
  const int N = q.bsa + 2 * ova;
  short *A = obmc_sm, *B = obmc_sm + N * N;
  const int xb = blockIdx.x, yb = blockIdx.y, c = blockIdx.z;
  const long long plane = (long long)q.BY * q.BX, b = (long long)yb * q.BX + xb;
  const short *U0 = q.ref.row(q.r0_slot * 3 + c, 0);
  const short *U1 = q.ref.row(q.r1_slot * 3 + c, 0);
  const int y0 = yb * q.bsa + q.mv[MV_PREV_Y * plane + b], x0 = xb * q.bsa + q.mv[MV_PREV_X * plane + b];
  const int y1 = yb * q.bsa + q.mv[MV_NEXT_Y * plane + b], x1 = xb * q.bsa + q.mv[MV_NEXT_X * plane + b];
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
    const int y = i / N - ova, x = i % N - ova;
    const int v = (int)bordered_ref(U0, q.ref.S, q.Ya, q.Xa, q.ba, q.padh, y0 + y, x0 + x) +
                  (int)bordered_ref(U1, q.ref.S, q.Ya, q.Xa, q.ba, q.padh, y1 + y, x1 + x);
    A[i] = (short)tdiv2(v);
  }
  __syncthreads();
  int n = N;
  for (int lv = 0; lv < levels; lv++) {
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
      const int y = i / n, j = i % n;
      B[y * N + j] = l53_ana_out(A + y * N, 1, j, n);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
      const int j = i / n, x = i % n;
      A[j * N + x] = l53_ana_out(B + x, N, j, n);
    }
    __syncthreads();
    n >>= 1;
    if (n == 0) n = 1;
  }
  for (int l = 1; l <= levels; l++) {
    const int s = q.bsa >> l, o_lo = ova >> l, o_hi = (q.bsa + 3 * ova) >> l;
    const int Yl = q.Ya >> l, Xl = q.Xa >> l;
    for (int i = threadIdx.x; i < s * s; i += blockDim.x) {
      const int y = i / s, x = i % s;
      q.pred.row(c, yb * s + y)[Xl + xb * s + x] = A[(o_lo + y) * N + o_hi + x];
      q.pred.row(c, Yl + yb * s + y)[xb * s + x] = A[(o_hi + y) * N + o_lo + x];
      q.pred.row(c, Yl + yb * s + y)[Xl + xb * s + x] = A[(o_hi + y) * N + o_hi + x];
    }
  }
  {
    const int s = q.bsa >> levels, o = ova >> levels;
    for (int i = threadIdx.x; i < s * s; i += blockDim.x) {
      const int y = i / s, x = i % s;
      q.pred.row(c, yb * s + y)[xb * s + x] = A[(o + y) * N + o + x];
    }
  }
}

bool predict_obmc_supported(int bsa, int ova) {
  const long long N = bsa + 2LL * ova;
  return 2 * N * N * (long long)sizeof(short) <= 200 * 1024;
}

void launch_predict_obmc(const Launch &L, const PredictParams &q, int ova, int levels) {
  if (q.BY <= 0 || q.BX <= 0) return;
  const int N = q.bsa + 2 * ova;
  const size_t smem = 2 * (size_t)N * N * sizeof(short);
  static size_t s_attr = 0;
  if (smem > 48 * 1024 && smem > s_attr) {
    cudaFuncSetAttribute(k_predict_obmc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    s_attr = smem;
  }
  ProfScope ps_(L, KC_PREDICT);
  k_predict_obmc<<<dim3(q.BX, q.BY, 3), 256, smem, L.stream>>>(q, ova, levels);
  COUNT(L);
}

// Rows >= cy and columns >= cx of the prediction planes are never written by
// predict(); they keep what the previous pair's in-place analysis left there and
// are clipped with everything else.
__global__ void k_clip_uncovered(Plane pred, int Ya, int Xa, int cy, int cx) {
  const int c = blockIdx.z;
  for (int y = blockIdx.y; y < Ya; y += gridDim.y) {
    short *row = pred.row(c, y);
    int x_begin = y < cy ? cx : 0;
    for (int x = x_begin + blockIdx.x * blockDim.x + threadIdx.x; x < Xa;
         x += gridDim.x * blockDim.x) {
      short v = row[x];
      row[x] = v < 0 ? (short)0 : (v > 255 ? (short)255 : v);
    }
  }
}

void launch_clip_uncovered(const Launch &L, Plane pred, int Ya, int Xa, int cy, int cx) {
  if (cy >= Ya && cx >= Xa) return;
  dim3 grid(8, Ya < 1024 ? Ya : 1024, 3);
  ProfScope ps_(L, KC_PREDICT);
  k_clip_uncovered<<<grid, 256, 0, L.stream>>>(pred, Ya, Xa, cy, cx);
  COUNT(L);
}

// Analysis: r = clamp(odd - LL, -128, 127); high = clamp(r + 128, 0, 255) ('B'
// variant; 'I' frames are patched afterwards with the raw odd frame).
// Synthesis: odd = 'I' ? high : clamp(high - 128 + LL, 0, 255).
__global__ void __launch_bounds__(256) k_residue(ResidueParams q) {
  __shared__ int h_pred[256], h_res[256];
  const int c = blockIdx.z;
  const int w = c ? q.X / 2 : q.X, h = c ? q.Y / 2 : q.Y;
  const long long off = c == 0 ? 0 : (long long)q.X * q.Y + (long long)(c - 1) * (q.X / 2) * (q.Y / 2);
  const bool do_hist = q.hist && c == 0 && !q.synth;
  if (do_hist) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h_pred[i] = h_res[i] = 0;
    __syncthreads();
  }
  for (int y = blockIdx.y; y < h; y += gridDim.y) {
    const short *ll = q.pred.row(c, y);
    const uint8_t *in = q.odd + off + (long long)y * w;
    uint8_t *out = q.out + off + (long long)y * w;
    uint8_t *pout = q.prediction ? q.prediction + off + (long long)y * w : nullptr;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) {
      int p = ll[x];
      int s = in[x];
      int o;
      if (!q.synth) {
        int r = s - p;
        r = r < -128 ? -128 : (r > 127 ? 127 : r);
        o = r + 128;  // already in [0,255]
        if (do_hist) {
          atomicAdd(&h_pred[s], 1);
          atomicAdd(&h_res[o], 1);
        }
      } else if (q.is_I) {
        o = s;
      } else {
        o = s - 128 + p;
        o = o < 0 ? 0 : (o > 255 ? 255 : o);
      }
      out[x] = (uint8_t)o;
      if (pout) pout[x] = (uint8_t)p;  // truncating store (texture.cpp:139-141)
    }
  }
  if (do_hist) {
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      if (h_pred[i]) atomicAdd(&q.hist[i], h_pred[i]);
      if (h_res[i]) atomicAdd(&q.hist[256 + i], h_res[i]);
    }
  }
}

void launch_residue(const Launch &L, const ResidueParams &q) {
  dim3 grid((q.X + 1023) / 1024, q.Y < 296 ? q.Y : 296, 3);
  ProfScope ps_(L, KC_RESIDUE);
  k_residue<<<grid, 256, 0, L.stream>>>(q);
  COUNT(L);
}

// decorrelate.cpp:803-816: 256-bin histogram of mv + 128 over the 4 planes.
// Components outside [-128,127] index outside the reference's static array
// (undefined behaviour there); they are counted in hist[256] and reported.
__global__ void k_mv_hist(const short *__restrict__ mv, int n, int *hist) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int v = mv[i] + 128;
    atomicAdd(&hist[(unsigned)v < 256u ? v : 256], 1);
  }
}

void launch_mv_hist(const Launch &L, const short *mv, int n, int *hist) {
  if (n <= 0) return;
  int blocks = (n + 255) / 256;
  ProfScope ps_(L, KC_RESIDUE);
  k_mv_hist<<<blocks < 64 ? blocks : 64, 256, 0, L.stream>>>(mv, n, hist);
  COUNT(L);
}

// Sum of squared differences of two byte streams per block of `block` bytes (the distortion side of
// psnr.py:78-90, which shells out to the external `snr --type=uchar --block_size=<bytes per picture>`):
// exact 64-bit integer sums, one CTA column per block, 16 bytes per thread and step.
__global__ void __launch_bounds__(256) k_sse_u8(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
                                                long long block, unsigned long long *out) {
  const uint8_t *pa = a + (long long)blockIdx.y * block, *pb = b + (long long)blockIdx.y * block;
  unsigned long long acc = 0;
  const bool al = ((((uintptr_t)pa) | ((uintptr_t)pb)) & 15) == 0;
  const long long nv = al ? block >> 4 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const uint4 x = reinterpret_cast<const uint4 *>(pa)[i], y = reinterpret_cast<const uint4 *>(pb)[i];
    const unsigned xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const unsigned d = __vabsdiffu4(xs[k], ys[k]);  // |x - y| per byte
      s = __dp4a(d, d, s);                           // + sum of squares of the four bytes
    }
    acc += s;
  }
  for (long long i = (nv << 4) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < block;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)pa[i] - (int)pb[i];
    acc += (unsigned)(d * d);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&out[blockIdx.y], acc);
}

void launch_sse_u8(const Launch &L, const uint8_t *a, const uint8_t *b, long long block, int nblocks,
                   unsigned long long *out) {
  if (nblocks <= 0 || block <= 0) return;
  int bx = (int)((block / 16 + 255) / 256);
  bx = bx < 1 ? 1 : (bx > 64 ? 64 : bx);
  for (int k0 = 0; k0 < nblocks; k0 += 65535) {
    const int nb = nblocks - k0 < 65535 ? nblocks - k0 : 65535;
    ProfScope ps_(L, KC_IMG);
    k_sse_u8<<<dim3(bx, nb), 256, 0, L.stream>>>(a + (long long)k0 * block, b + (long long)k0 * block, block, out + k0);
    COUNT(L);
  }
}

// Motion-field (de)correlation, the step after the analysis on the motion side.
// bidirectional (bidirectional_motion_decorrelate.cpp:25-52): NEXT -= PREV (+= when inverse);
// interlevel (interlevel_motion_decorrelate.cpp:32-69, 250-297): field k of level t is predicted
// by half of field k/2 of level t+1 (C division: truncation towards zero); a reference file that
// ends early leaves its last field in the reader's buffer, a missing one reads as zeros.
// Fields are [PREV.X, PREV.Y, NEXT.X, NEXT.Y][by][bx] int16; results wrap like the reference's shorts.
__global__ void k_mv_bidirectional(const short *__restrict__ in, short *__restrict__ out, int n_fields, int plane,
                                   int inverse) {
  const long long total = (long long)n_fields * 4 * plane;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % (4LL * plane));
    int v = in[i];
    if (e >= 2 * plane) v = inverse ? v + in[i - 2 * plane] : v - in[i - 2 * plane];
    out[i] = (short)v;
  }
}

__global__ void k_mv_interlevel(const short *__restrict__ in, const short *__restrict__ ref, short *__restrict__ out,
                                int n_fields, int n_ref, int plane, int inverse) {
  const long long fsz = 4LL * plane, total = (long long)n_fields * fsz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long f = i / fsz, e = i - f * fsz;
    int r = 0;
    if (n_ref > 0) r = tdiv2(ref[(f / 2 < n_ref ? f / 2 : n_ref - 1) * fsz + e]);
    out[i] = (short)(inverse ? in[i] + r : in[i] - r);
  }
}

void launch_mv_bidirectional(const Launch &L, const short *in, short *out, int n_fields, int plane, int inverse) {
  const long long total = (long long)n_fields * 4 * plane;
  if (total <= 0) return;
  long long blocks = (total + 255) / 256;
  ProfScope ps_(L, KC_IMG);
  k_mv_bidirectional<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, L.stream>>>(in, out, n_fields, plane,
                                                                                        inverse);
  COUNT(L);
}

void launch_mv_interlevel(const Launch &L, const short *in, const short *ref, short *out, int n_fields, int n_ref,
                          int plane, int inverse) {
  const long long total = (long long)n_fields * 4 * plane;
  if (total <= 0) return;
  long long blocks = (total + 255) / 256;
  ProfScope ps_(L, KC_IMG);
  k_mv_interlevel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, L.stream>>>(in, ref, out, n_fields,
                                                                                     n_ref, plane, inverse);
  COUNT(L);
}

__global__ void k_copy_bytes(uint8_t *dst, const uint8_t *src, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
    size_t n16 = n / 16;
    const uint4 *s4 = (const uint4 *)src;
    uint4 *d4 = (uint4 *)dst;
    for (size_t k = i; k < n16; k += stride) d4[k] = s4[k];
    for (size_t k = n16 * 16 + i; k < n; k += stride) dst[k] = src[k];
  } else {
    for (size_t k = i; k < n; k += stride) dst[k] = src[k];
  }
}

void launch_copy_bytes(const Launch &L, void *dst, const void *src, size_t n) {
  if (!n) return;
  size_t blocks = (n / 16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope ps_(L, KC_IMG);
  k_copy_bytes<<<(int)blocks, 256, 0, L.stream>>>((uint8_t *)dst, (const uint8_t *)src, n);
  COUNT(L);
}

// `count` blocks of `n` bytes, block k at dst + k*dst_stride / src + k*src_stride (one launch)
__global__ void k_copy_strided(uint8_t *dst, long long dst_stride, const uint8_t *src, long long src_stride,
                               size_t n) {
  uint8_t *d = dst + (long long)blockIdx.y * dst_stride;
  const uint8_t *s = src + (long long)blockIdx.y * src_stride;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  if ((((uintptr_t)d | (uintptr_t)s) & 15) == 0) {
    size_t n16 = n / 16;
    const uint4 *s4 = (const uint4 *)s;
    uint4 *d4 = (uint4 *)d;
    for (size_t k = i; k < n16; k += stride) d4[k] = s4[k];
    for (size_t k = n16 * 16 + i; k < n; k += stride) d[k] = s[k];
  } else {
    for (size_t k = i; k < n; k += stride) d[k] = s[k];
  }
}

void launch_copy_strided(const Launch &L, void *dst, long long dst_stride, const void *src,
                         long long src_stride, size_t n, int count) {
  if (!n || count <= 0) return;
  size_t blocks = (n / 16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 4) blocks = 148 * 4;
  for (int k0 = 0; k0 < count; k0 += 65535) {
    const int m = count - k0 < 65535 ? count - k0 : 65535;
    ProfScope ps_(L, KC_IMG);
    k_copy_strided<<<dim3((unsigned)blocks, (unsigned)m), 256, 0, L.stream>>>(
        (uint8_t *)dst + (long long)k0 * dst_stride, dst_stride, (const uint8_t *)src + (long long)k0 * src_stride,
        src_stride, n);
    COUNT(L);
  }
}

__global__ void k_load_residue(Plane dst, int slot, const uint8_t *__restrict__ src, int h, int w) {
  for (int y = blockIdx.y; y < h; y += gridDim.y) {
    short *row = dst.row(slot, y);
    const uint8_t *srow = src + (long long)y * w;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x)
      row[x] = (short)((int)srow[x] - 128);
  }
}

void launch_load_residue(const Launch &L, Plane dst, int slot, const uint8_t *src, int h, int w) {
  dim3 grid((w + 255) / 256, h < 1024 ? h : 1024);
  ProfScope ps_(L, KC_IMG);
  k_load_residue<<<grid, 256, 0, L.stream>>>(dst, slot, src, h, w);
  COUNT(L);
}

// update.cpp:71-148 as a target-centric gather.  The reference scatters
// sequentially over (c, by, bx, y, x): ref[clip(src + mv)] = (short)clamp(
// float(ref) +- residue * uf).  Several sources can hit one target (overlapping
// displaced blocks, edge clipping) and every contribution is clamped and
// truncated before the next one, so the order is part of the result.  One CTA
// owns a 16x16 tile of targets: it scans the blocks in raster order, keeps those
// whose displaced (clipped) footprint touches the tile, and each thread replays
// the contributions to its own pixel in (by, bx, y, x) order.  PREV and NEXT
// touch different frames and the three components different planes, so the
// chains are independent (one launch per frame and direction).
__global__ void __launch_bounds__(256) k_update(UpdateParams q) {
  __shared__ int s_list[256];
  __shared__ int s_count;
  const int c = blockIdx.z;
  const int tx = blockIdx.x * 16 + (threadIdx.x & 15);
  const int ty = blockIdx.y * 16 + (threadIdx.x >> 4);
  const int tile_x0 = blockIdx.x * 16, tile_y0 = blockIdx.y * 16;
  const int tile_x1 = min(tile_x0 + 15, q.X - 1), tile_y1 = min(tile_y0 + 15, q.Y - 1);
  const bool active = tx < q.X && ty < q.Y;
  const long long plane = (long long)q.BY * q.BX;
  const short *mvx = q.mv + (long long)q.dir * plane;
  const short *mvy = q.mv + (long long)(q.dir + 1) * plane;
  // only blocks within `reach` (largest |vector component| of this field) of the tile can
  // touch it; candidates are enumerated in raster order, like the reference's scatter
  const int reach = *q.reach;
  const int by_lo = max(0, (tile_y0 - reach - q.bs + 1 + (q.bs - 1) * (tile_y0 - reach - q.bs + 1 > 0)) / q.bs);
  const int by_hi = min(q.BY - 1, (tile_y1 + reach) / q.bs);
  const int bx_lo = max(0, (tile_x0 - reach - q.bs + 1 + (q.bs - 1) * (tile_x0 - reach - q.bs + 1 > 0)) / q.bs);
  const int bx_hi = min(q.BX - 1, (tile_x1 + reach) / q.bs);
  const int nbw = max(bx_hi - bx_lo + 1, 0), nbh = max(by_hi - by_lo + 1, 0);
  const int nblocks = nbw * nbh;
  float aux = 0.f;
  short *target = nullptr;
  if (active) {
    target = q.ref.row(c, ty) + tx;
    aux = (float)*target;
  }
  short cur = active ? *target : (short)0;
  for (int base = 0; base < nblocks; base += 256) {
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // ordered compaction of the blocks that can reach this tile
    int b = -1;
    bool hit = false;
    if (base + threadIdx.x < nblocks) {
      const int k = base + threadIdx.x;
      const int cby = by_lo + k / nbw, cbx = bx_lo + k % nbw;
      b = cby * q.BX + cbx;
      int oy = cby * q.bs + mvy[b], ox = cbx * q.bs + mvx[b];
      int fy0 = iclamp(oy, 0, q.Y - 1), fy1 = iclamp(oy + q.bs - 1, 0, q.Y - 1);
      int fx0 = iclamp(ox, 0, q.X - 1), fx1 = iclamp(ox + q.bs - 1, 0, q.X - 1);
      hit = fy0 <= tile_y1 && fy1 >= tile_y0 && fx0 <= tile_x1 && fx1 >= tile_x0;
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    __shared__ int s_warp[8];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int prefix = 0;
    for (int w = 0; w < warp; w++) prefix += s_warp[w];
    if (hit) s_list[prefix + __popc(m & ((1u << lane) - 1))] = b;
    if (threadIdx.x == 255) s_count = prefix + __popc(m);
    __syncthreads();
    const int cnt = s_count;
    if (active) {
      for (int k = 0; k < cnt; k++) {
        int bb = s_list[k];
        int byy = bb / q.BX, bxx = bb % q.BX;
        int oy = byy * q.bs + mvy[bb], ox = bxx * q.bs + mvx[bb];
        // source rows y in [0,bs) with clip(oy + y) == ty, in increasing order
        // (edge targets collect every source that clip() folds onto them)
        int ylo = (ty == 0) ? 0 : ty - oy, yhi = (ty == q.Y - 1) ? q.bs - 1 : ty - oy;
        int xlo = (tx == 0) ? 0 : tx - ox, xhi = (tx == q.X - 1) ? q.bs - 1 : tx - ox;
        ylo = max(ylo, 0); yhi = min(yhi, q.bs - 1);
        xlo = max(xlo, 0); xhi = min(xhi, q.bs - 1);
        for (int y = ylo; y <= yhi; y++) {
          const short *rrow = q.res.row(c, byy * q.bs + y) + bxx * q.bs;
          for (int x = xlo; x <= xhi; x++) {
            float prod = __fmul_rn((float)rrow[x], q.uf);
            aux = (float)cur;
            aux = q.inverse ? __fsub_rn(aux, prod) : __fadd_rn(aux, prod);
            if (aux > 255.f) aux = 255.f;
            else if (aux < 0.f) aux = 0.f;
            cur = (short)aux;  // float -> short truncation toward zero
          }
        }
      }
    }
    __syncthreads();
  }
  if (active) *target = cur;
}

// ---- update / un_update for all frames of a level at once ----
// Frames are independent of each other: frame k first takes the NEXT contributions of pair k-1,
// then the PREV contributions of pair k (update.cpp:506-656).  Instead of scanning every block
// that could reach a tile, the blocks are binned first: k_update_bin appends each block of each
// (pair, direction) to the 16x16 target tiles its displaced, edge-clipped footprint touches
// (bounded lists; a tile whose list overflows -- blocks folded onto a picture edge -- falls back
// to the scan), and k_update_batch sorts a tile's short list back into raster order (the order
// is part of the result) and replays the contributions per target pixel.
__global__ void __launch_bounds__(256) k_update_bin(UpdateBatchParams q) {
  const int pd = blockIdx.y, pair = pd >> 1, dir = pd & 1;  // dir 0: PREV, 1: NEXT
  if (q.types[pair] != 'B') return;
  const long long plane = (long long)q.BY * q.BX;
  const short *mvx = q.mv + (long long)pair * 4 * plane + (long long)(dir ? MV_NEXT_X : MV_PREV_X) * plane;
  const short *mvy = mvx + plane;
  const int ntiles = q.tiles_x * q.tiles_y;
  int *cnt = q.cnt + (long long)pd * ntiles;
  int *list = q.list + (long long)pd * ntiles * q.cap;
  int reach = 0;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < (int)plane; b += gridDim.x * blockDim.x) {
    const int by = b / q.BX, bx = b - by * q.BX;
    const int vy = mvy[b], vx = mvx[b];
    reach = max(reach, max(abs(vy), abs(vx)));
    const int oy = by * q.bs + vy, ox = bx * q.bs + vx;
    const int fy0 = iclamp(oy, 0, q.Y - 1) >> 4, fy1 = iclamp(oy + q.bs - 1, 0, q.Y - 1) >> 4;
    const int fx0 = iclamp(ox, 0, q.X - 1) >> 4, fx1 = iclamp(ox + q.bs - 1, 0, q.X - 1) >> 4;
    for (int ty = fy0; ty <= fy1; ty++)
      for (int tx = fx0; tx <= fx1; tx++) {
        const int t = ty * q.tiles_x + tx;
        const int pos = atomicAdd(&cnt[t], 1);
        if (pos < q.cap) {
          list[(long long)t * q.cap + pos] = b;
          if (q.geo) q.geo[((long long)pd * ntiles + t) * q.cap + pos] = make_int4(oy, ox, by * q.bs, bx * q.bs);
        }
      }
  }
  reach = __reduce_max_sync(0xffffffffu, reach);
  if ((threadIdx.x & 31) == 0 && reach > 0) atomicMax(&q.reach[pd], reach);
}

__global__ void __launch_bounds__(256) k_update_batch(UpdateBatchParams q) {
  __shared__ int s_list[256];
  __shared__ int s_sorted[256];
  __shared__ int s_count;
  __shared__ int s_warp[8];
  const int frame = q.frame0 + blockIdx.z / 3, c = blockIdx.z % 3;
  const int tx = blockIdx.x * 16 + (threadIdx.x & 15);
  const int ty = blockIdx.y * 16 + (threadIdx.x >> 4);
  const int tile_x0 = blockIdx.x * 16, tile_y0 = blockIdx.y * 16;
  const int tile_x1 = min(tile_x0 + 15, q.X - 1), tile_y1 = min(tile_y0 + 15, q.Y - 1);
  const bool active = tx < q.X && ty < q.Y;
  const long long plane = (long long)q.BY * q.BX;
  const int ntiles = q.tiles_x * q.tiles_y, tile = blockIdx.y * q.tiles_x + blockIdx.x;
  short *target = active ? q.ref.row(c * q.slots_per_comp + (frame - q.frame0), ty) + tx : nullptr;
  short cur = active ? *target : (short)0;
  const int cw = c ? q.X >> 1 : q.X, ch = c ? q.Y >> 1 : q.Y;  // residue[1|2] only exists in its top-left quarter
  const long long coff = c == 0 ? 0 : (long long)q.X * q.Y + (long long)(c - 1) * (q.X / 2) * (q.Y / 2);

  for (int pass = 0; pass < 2; pass++) {
    // frame k first receives pair k-1's NEXT update, then pair k's PREV update
    const int pair = pass == 0 ? frame - 1 : frame, dir = pass == 0 ? 1 : 0;
    if (pair < 0 || pair >= q.n_pairs || q.types[pair] != 'B') continue;  // uniform per CTA
    const int pd = pair * 2 + dir;
    const short *mvx = q.mv + (long long)pair * 4 * plane + (long long)(dir ? MV_NEXT_X : MV_PREV_X) * plane;
    const short *mvy = mvx + plane;
    const uint8_t *res = q.high + (long long)pair * q.high_stride + coff;
    const int total = q.cnt[(long long)pd * ntiles + tile];
    auto replay = [&](int cnt, const int *ids) {
      if (!active) return;
      for (int k = 0; k < cnt; k++) {
        const int bb = ids[k];
        const int byy = bb / q.BX, bxx = bb - byy * q.BX;
        const int oy = byy * q.bs + mvy[bb], ox = bxx * q.bs + mvx[bb];
        // source rows y in [0,bs) with clip(oy + y) == ty, in increasing order
        // (edge targets collect every source that clip() folds onto them)
        int ylo = (ty == 0) ? 0 : ty - oy, yhi = (ty == q.Y - 1) ? q.bs - 1 : ty - oy;
        int xlo = (tx == 0) ? 0 : tx - ox, xhi = (tx == q.X - 1) ? q.bs - 1 : tx - ox;
        ylo = max(ylo, 0), yhi = min(yhi, q.bs - 1);
        xlo = max(xlo, 0), xhi = min(xhi, q.bs - 1);
        for (int y = ylo; y <= yhi; y++) {
          const int ry = byy * q.bs + y;
          for (int x = xlo; x <= xhi; x++) {
            const int rx = bxx * q.bs + x;
            const int r = (ry < ch && rx < cw) ? (int)res[(long long)ry * cw + rx] - 128 : 0;
            const float prod = __fmul_rn((float)r, q.uf);
            float aux = (float)cur;
            aux = q.inverse ? __fsub_rn(aux, prod) : __fadd_rn(aux, prod);
            if (aux > 255.f) aux = 255.f;
            else if (aux < 0.f) aux = 0.f;
            cur = (short)aux;  // float -> short truncation toward zero
          }
        }
      }
    };
    if (total <= q.cap) {
      // short list: rank sort back into raster order (ids are distinct)
      __syncthreads();
      if ((int)threadIdx.x < total) s_list[threadIdx.x] = q.list[((long long)pd * ntiles + tile) * q.cap + threadIdx.x];
      __syncthreads();
      if ((int)threadIdx.x < total) {
        const int me = s_list[threadIdx.x];
        int rank = 0;
        for (int k = 0; k < total; k++) rank += s_list[k] < me;
        s_sorted[rank] = me;
      }
      __syncthreads();
      replay(total, s_sorted);
    } else {
      // overflow: ordered scan of every block within reach of the tile (blocks folded onto an edge)
      const int reach = q.reach[pd];
      const int by_lo = max(0, (tile_y0 - reach - q.bs + 1 + (q.bs - 1) * (tile_y0 - reach - q.bs + 1 > 0)) / q.bs);
      const int by_hi = min(q.BY - 1, (tile_y1 + reach) / q.bs);
      const int bx_lo = max(0, (tile_x0 - reach - q.bs + 1 + (q.bs - 1) * (tile_x0 - reach - q.bs + 1 > 0)) / q.bs);
      const int bx_hi = min(q.BX - 1, (tile_x1 + reach) / q.bs);
      const int nbw = max(bx_hi - bx_lo + 1, 0), nbh = max(by_hi - by_lo + 1, 0);
      const int nblocks = nbw * nbh;
      for (int base = 0; base < nblocks; base += 256) {
        __syncthreads();
        int b = -1;
        bool hit = false;
        if (base + (int)threadIdx.x < nblocks) {
          const int k = base + threadIdx.x;
          const int cby = by_lo + k / nbw, cbx = bx_lo + k % nbw;
          b = cby * q.BX + cbx;
          const int oy = cby * q.bs + mvy[b], ox = cbx * q.bs + mvx[b];
          const int fy0 = iclamp(oy, 0, q.Y - 1), fy1 = iclamp(oy + q.bs - 1, 0, q.Y - 1);
          const int fx0 = iclamp(ox, 0, q.X - 1), fx1 = iclamp(ox + q.bs - 1, 0, q.X - 1);
          hit = fy0 <= tile_y1 && fy1 >= tile_y0 && fx0 <= tile_x1 && fx1 >= tile_x0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int prefix = 0;
        for (int w = 0; w < warp; w++) prefix += s_warp[w];
        if (hit) s_list[prefix + __popc(m & ((1u << lane) - 1))] = b;
        if (threadIdx.x == 255) s_count = prefix + __popc(m);
        __syncthreads();
        replay(s_count, s_list);
      }
    }
  }
  if (active) *target = cur;
}

// ---- the same for a dyadic factor (uf * 256 integral: the codec's 0.25) ----
// With uf = j / 256 every step of the reference's chain  ref = (short)clamp(float(ref) +- residue * uf, 0, 255)
// is exact in float and equals the integer saturating add  ref = clamp(ref + ((+-residue * j) >> 8), 0, 255):
// ref is integral, so truncating a non-negative sum is its floor and the floor moves onto the addend, and a
// negative sum clamps to 0 either way.  Saturating adds compose -- a run of them is x -> clamp(x + A, L, H), and
// run g after run f is (Af + Ag, clamp(Lf + Ag, Lg, Hg), clamp(Hf + Ag, Lg, Hg)) -- so the long chains of the
// targets on the picture edge (every source that clip() folds onto them, in order) are evaluated by a warp: a
// lane composes one block's contributions, an ordered tree of shuffles composes the lanes.  Targets inside the
// picture take at most one contribution per block.  One CTA = one 16x16 tile of one frame, all three components
// (they share the block lists and the geometry), both passes.
struct SatAdd {
  int A, L, H;
};
static constexpr int SAT_BIG = 1 << 24;
__device__ __forceinline__ void sat_push(SatAdd &f, int k) {  // ... then x -> clamp(x + k, 0, 255)
  f.A += k;
  f.L = min(max(f.L + k, 0), 255);
  f.H = min(max(f.H + k, 0), 255);
}
__device__ __forceinline__ SatAdd sat_then(const SatAdd &f, const SatAdd &g) {  // g after f
  SatAdd r;
  r.A = f.A + g.A;
  r.L = min(max(f.L + g.A, g.L), g.H);
  r.H = min(max(f.H + g.A, g.L), g.H);
  return r;
}

// perimeter != 0: blockIdx.x enumerates the tiles on the picture's perimeter (top row, bottom row, then left / right
// column pairs); the tiles inside go through k_update_quad
__global__ void __launch_bounds__(256, 4) k_update_dyadic(UpdateBatchParams q, int j256, int perimeter) {
  // per listed block, raster order: displaced origin (oy, ox), source origin (sy0, sx0)
  __shared__ int4 s_geo[64];    // the short lists of the two passes (cap <= 32 each)
  __shared__ int4 s_scan[256];  // one round of the overflow scan
  __shared__ int s_ids[64];
  __shared__ int s_warp[8];
  __shared__ int s_count, s_nheavy;
  __shared__ int s_hp[64];       // threads that own a target on the picture edge
  __shared__ int s_hcur[3][64];  // ... and the running values of those targets
  const int frame = q.frame0 + blockIdx.z;
  int tbx = blockIdx.x, tby = blockIdx.y;
  if (perimeter) {
    const int e = blockIdx.x;
    if (e < q.tiles_x) {
      tbx = e, tby = 0;
    } else if (e < 2 * q.tiles_x) {
      tbx = e - q.tiles_x, tby = q.tiles_y - 1;
    } else {
      const int k = e - 2 * q.tiles_x;
      tbx = (k & 1) ? q.tiles_x - 1 : 0, tby = 1 + (k >> 1);
    }
  }
  const int tile_x0 = tbx * 16, tile_y0 = tby * 16;
  const int tx = tile_x0 + (threadIdx.x & 15), ty = tile_y0 + (threadIdx.x >> 4);
  const int tile_x1 = min(tile_x0 + 15, q.X - 1), tile_y1 = min(tile_y0 + 15, q.Y - 1);
  const bool active = tx < q.X && ty < q.Y;
  const bool heavy = active && (tx == 0 || ty == 0 || tx == q.X - 1 || ty == q.Y - 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long plane = (long long)q.BY * q.BX;
  const int ntiles = q.tiles_x * q.tiles_y, tile = tby * q.tiles_x + tbx;
  const int cw = q.X >> 1, ch = q.Y >> 1;  // residue[1|2] only exists in its top-left quarter
  const long long coff1 = (long long)q.X * q.Y, coff2 = coff1 + (long long)cw * ch;
  const int jj = q.inverse ? -j256 : j256;

  short *target[3] = {nullptr, nullptr, nullptr};
  int cur[3] = {0, 0, 0};
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (c == 0 && q.luma_in) {
        cur[0] = q.luma_in[(long long)frame * q.luma_in_stride + (long long)ty * q.X + tx];
      } else {
        target[c] = q.ref.row(c * q.slots_per_comp + (frame - q.frame0), ty) + tx;
        cur[c] = *target[c];
      }
    }
  }
  // ordered list of the edge targets of this tile (at most 60); none in the 97 % of tiles inside the picture
  int my_h = -1;
  const bool edge_tile = tile_x0 == 0 || tile_y0 == 0 || tile_x1 == q.X - 1 || tile_y1 == q.Y - 1;  // CTA-uniform
  if (edge_tile) {
    const unsigned m = __ballot_sync(0xffffffffu, heavy);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int prefix = 0;
    for (int w = 0; w < warp; w++) prefix += s_warp[w];
    if (heavy) {
      my_h = prefix + __popc(m & ((1u << lane) - 1));
      s_hp[my_h] = threadIdx.x;
#pragma unroll
      for (int c = 0; c < 3; c++) s_hcur[c][my_h] = cur[c];
    }
    if (threadIdx.x == 255) s_nheavy = prefix + __popc(m);
    __syncthreads();
  }
  const int nheavy = edge_tile ? s_nheavy : 0;

  // frame k first receives pair k-1's NEXT update (pass 0), then pair k's PREV update (pass 1).  The short
  // lists of both passes are fetched, sorted and resolved to block geometry together (warp 0: pass 0,
  // warp 1: pass 1), so that the chain of dependent loads (count -> list -> vectors -> residues) is paid once.
  int total[2];
  bool valid[2];
  const short *mvx_[2];
  const uint8_t *res_[2];
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const int pair = pass == 0 ? frame - 1 : frame, dir = pass == 0 ? 1 : 0;
    valid[pass] = pair >= 0 && pair < q.n_pairs && q.types[pair] == 'B';  // uniform per CTA
    const int pc = valid[pass] ? pair : 0;
    total[pass] = valid[pass] ? q.cnt[(long long)(pc * 2 + dir) * ntiles + tile] : 0;
    mvx_[pass] = q.mv + (long long)pc * 4 * plane + (long long)(dir ? MV_NEXT_X : MV_PREV_X) * plane;
    res_[pass] = q.high + (long long)pc * q.high_stride;
  }
  auto geometry = [&](int pass, int b) -> int4 {
    const int byy = b / q.BX, bxx = b - byy * q.BX;
    return make_int4(byy * q.bs + mvx_[pass][plane + b], bxx * q.bs + mvx_[pass][b], byy * q.bs, bxx * q.bs);
  };
  if (warp < 2) {
    // the list entries are fetched before the counts are known (lanes beyond a count read unused slots), with
    // their geometry when the binning pass stored it: count and list cost one memory round trip, not three
    const int pass = warp;
    const int pair = pass == 0 ? frame - 1 : frame, dir = pass == 0 ? 1 : 0;
    int me = 0;
    int4 g = make_int4(0, 0, 0, 0);
    if (valid[pass] && lane < q.cap) {
      const long long at = ((long long)(pair * 2 + dir) * ntiles + tile) * q.cap + lane;
      if (q.geo) {
        g = q.geo[at];
        me = (g.z << 16) | g.w;  // raster order of the blocks = order of their source origins (X <= 16384)
      } else {
        me = q.list[at];
      }
    }
    const int n = total[pass];
    const bool mine = valid[pass] && n <= q.cap && lane < n;
    if (mine) s_ids[pass * 32 + lane] = me;
    __syncwarp();
    if (mine) {
      // rank sort back into raster order (keys are distinct)
      int rank = 0;
      for (int k = 0; k < n; k++) rank += s_ids[pass * 32 + k] < me;
      s_geo[pass * 32 + rank] = q.geo ? g : geometry(pass, me);
    }
  }
  __syncthreads();

  for (int pass = 0; pass < 2; pass++) {
    if (!valid[pass]) continue;
    const uint8_t *res = res_[pass];
    // addend of source sample (ry, rx) of component c
    auto addend = [&](int c, int ry, int rx) -> int {
      int r = 0;
      if (c == 0) r = (int)res[(long long)ry * q.X + rx] - 128;
      else if (ry < ch && rx < cw) r = (int)res[(c == 1 ? coff1 : coff2) + (long long)ry * cw + rx] - 128;
      return (r * jj) >> 8;
    };
    // n listed blocks (geo, raster order) onto this tile's targets
    auto apply = [&](int n, const int4 *geo) {
      if (active && !heavy) {
        for (int k = 0; k < n; k++) {
          const int4 g = geo[k];
          const int y = ty - g.x, x = tx - g.y;
          if ((unsigned)y < (unsigned)q.bs && (unsigned)x < (unsigned)q.bs) {
#pragma unroll
            for (int c = 0; c < 3; c++) cur[c] = min(max(cur[c] + addend(c, g.z + y, g.w + x), 0), 255);
          }
        }
      }
      for (int h = warp; h < nheavy; h += 8) {  // warp-uniform
        const int tid = s_hp[h];
        const int hx = tile_x0 + (tid & 15), hy = tile_y0 + (tid >> 4);
        SatAdd acc[3];
#pragma unroll
        for (int c = 0; c < 3; c++) acc[c] = SatAdd{0, -SAT_BIG, SAT_BIG};
        for (int base = 0; base < n; base += 32) {
          SatAdd f[3];
#pragma unroll
          for (int c = 0; c < 3; c++) f[c] = SatAdd{0, -SAT_BIG, SAT_BIG};
          if (base + lane < n) {
            const int4 g = geo[base + lane];
            // source rows y in [0,bs) with clip(oy + y) == hy, in increasing order (edge targets collect every
            // source that clip() folds onto them); the same for the columns
            int ylo = (hy == 0) ? 0 : hy - g.x, yhi = (hy == q.Y - 1) ? q.bs - 1 : hy - g.x;
            int xlo = (hx == 0) ? 0 : hx - g.y, xhi = (hx == q.X - 1) ? q.bs - 1 : hx - g.y;
            ylo = max(ylo, 0), yhi = min(yhi, q.bs - 1);
            xlo = max(xlo, 0), xhi = min(xhi, q.bs - 1);
            for (int y = ylo; y <= yhi; y++)
              for (int x = xlo; x <= xhi; x++) {
#pragma unroll
                for (int c = 0; c < 3; c++) sat_push(f[c], addend(c, g.z + y, g.w + x));
              }
          }
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
              SatAdd g;
              g.A = __shfl_down_sync(0xffffffffu, f[c].A, off);
              g.L = __shfl_down_sync(0xffffffffu, f[c].L, off);
              g.H = __shfl_down_sync(0xffffffffu, f[c].H, off);
              if ((lane & (2 * off - 1)) == 0) f[c] = sat_then(f[c], g);
            }
          }
#pragma unroll
          for (int c = 0; c < 3; c++) acc[c] = sat_then(acc[c], f[c]);  // lane 0 holds the 32 blocks' run
        }
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 3; c++) s_hcur[c][h] = min(max(s_hcur[c][h] + acc[c].A, acc[c].L), acc[c].H);
        }
      }
    };

    if (total[pass] <= q.cap) {
      apply(total[pass], s_geo + pass * 32);
    } else {
      // overflow: ordered scan of every block within reach of the tile (blocks folded onto an edge)
      const int pair = pass == 0 ? frame - 1 : frame, dir = pass == 0 ? 1 : 0;
      const int reach = q.reach[pair * 2 + dir];
      const int by_lo = max(0, (tile_y0 - reach - q.bs + 1 + (q.bs - 1) * (tile_y0 - reach - q.bs + 1 > 0)) / q.bs);
      const int by_hi = min(q.BY - 1, (tile_y1 + reach) / q.bs);
      const int bx_lo = max(0, (tile_x0 - reach - q.bs + 1 + (q.bs - 1) * (tile_x0 - reach - q.bs + 1 > 0)) / q.bs);
      const int bx_hi = min(q.BX - 1, (tile_x1 + reach) / q.bs);
      const int nbw = max(bx_hi - bx_lo + 1, 0), nbh = max(by_hi - by_lo + 1, 0);
      const int nblocks = nbw * nbh;
      for (int base = 0; base < nblocks; base += 256) {
        __syncthreads();
        bool hit = false;
        int4 g = make_int4(0, 0, 0, 0);
        if (base + (int)threadIdx.x < nblocks) {
          const int k = base + threadIdx.x;
          const int cby = by_lo + k / nbw, cbx = bx_lo + k % nbw;
          g = geometry(pass, cby * q.BX + cbx);
          const int fy0 = iclamp(g.x, 0, q.Y - 1), fy1 = iclamp(g.x + q.bs - 1, 0, q.Y - 1);
          const int fx0 = iclamp(g.y, 0, q.X - 1), fx1 = iclamp(g.y + q.bs - 1, 0, q.X - 1);
          hit = fy0 <= tile_y1 && fy1 >= tile_y0 && fx0 <= tile_x1 && fx1 >= tile_x0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int prefix = 0;
        for (int w = 0; w < warp; w++) prefix += s_warp[w];
        if (hit) s_scan[prefix + __popc(m & ((1u << lane) - 1))] = g;
        if (threadIdx.x == 255) s_count = prefix + __popc(m);
        __syncthreads();
        apply(s_count, s_scan);
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int v = heavy ? s_hcur[c][my_h] : cur[c];
      if (c == 0 && q.luma_in)  // a byte either way: untouched input or a clamped sum
        q.luma_out[(long long)frame * q.luma_out_stride + (long long)ty * q.X + tx] = (uint8_t)v;
      else
        *target[c] = (short)v;
    }
  }
}

// The tiles inside the picture (no target on the picture edge, always 16 x 16): 64 threads, four horizontally
// adjacent targets per thread, so that the per-tile work that is not arithmetic on targets (counts, lists, sorting,
// addresses: 70 % of k_update_dyadic's instructions) is spread over four times the targets.  Luma targets are the
// frames' bytes (q.luma_in / q.luma_out; X % 8 == 0).  grid (tiles_x - 2, tiles_y - 2, frames).
__global__ void __launch_bounds__(64) k_update_quad(UpdateBatchParams q, int j256) {
  __shared__ int4 s_geo[64];   // the short lists of the two passes (cap <= 32 each), raster order
  __shared__ int4 s_scan[64];  // one round of the overflow scan
  __shared__ int s_ids[64];
  __shared__ int s_w[2], s_count;
  const int frame = q.frame0 + blockIdx.z;
  const int tbx = blockIdx.x + 1, tby = blockIdx.y + 1;
  const int tile_x0 = tbx * 16, tile_y0 = tby * 16, tile_x1 = tile_x0 + 15, tile_y1 = tile_y0 + 15;
  const int ty = tile_y0 + (threadIdx.x >> 2), tx0 = tile_x0 + 4 * (threadIdx.x & 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long plane = (long long)q.BY * q.BX;
  const int ntiles = q.tiles_x * q.tiles_y, tile = tby * q.tiles_x + tbx;
  const int cw = q.X >> 1, ch = q.Y >> 1;
  const long long coff1 = (long long)q.X * q.Y, coff2 = coff1 + (long long)cw * ch;
  const int jj = q.inverse ? -j256 : j256;

  int cur[3][4];
  short *target[2];
  {
    const unsigned w = *reinterpret_cast<const unsigned *>(q.luma_in + (long long)frame * q.luma_in_stride +
                                                           (unsigned)(ty * q.X + tx0));
#pragma unroll
    for (int k = 0; k < 4; k++) cur[0][k] = (w >> (8 * k)) & 0xff;
#pragma unroll
    for (int c = 1; c < 3; c++) {
      target[c - 1] = q.ref.row(c * q.slots_per_comp + (frame - q.frame0), ty) + tx0;
      const uint2 v = *reinterpret_cast<const uint2 *>(target[c - 1]);
      cur[c][0] = (short)(v.x & 0xffff);
      cur[c][1] = (int)v.x >> 16;
      cur[c][2] = (short)(v.y & 0xffff);
      cur[c][3] = (int)v.y >> 16;
    }
  }
  int total[2];
  bool valid[2];
  const short *mvx_[2];
  const uint8_t *res_[2];
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const int pair = pass == 0 ? frame - 1 : frame, dir = pass == 0 ? 1 : 0;
    valid[pass] = pair >= 0 && pair < q.n_pairs && q.types[pair] == 'B';
    const int pc = valid[pass] ? pair : 0;
    total[pass] = valid[pass] ? q.cnt[(long long)(pc * 2 + dir) * ntiles + tile] : 0;
    mvx_[pass] = q.mv + (long long)pc * 4 * plane + (long long)(dir ? MV_NEXT_X : MV_PREV_X) * plane;
    res_[pass] = q.high + (long long)pc * q.high_stride;
  }
  auto geometry = [&](int pass, int b) -> int4 {
    const int byy = b / q.BX, bxx = b - byy * q.BX;
    return make_int4(byy * q.bs + mvx_[pass][plane + b], bxx * q.bs + mvx_[pass][b], byy * q.bs, bxx * q.bs);
  };
  {
    // warp 0: pass 0, warp 1: pass 1 (k_update_dyadic)
    const int pass = warp;
    const int pair = pass == 0 ? frame - 1 : frame, dir = pass == 0 ? 1 : 0;
    int me = 0;
    int4 g = make_int4(0, 0, 0, 0);
    if (valid[pass] && lane < q.cap) {
      const long long at = ((long long)(pair * 2 + dir) * ntiles + tile) * q.cap + lane;
      if (q.geo) {
        g = q.geo[at];
        me = (g.z << 16) | g.w;
      } else {
        me = q.list[at];
      }
    }
    const int n = total[pass];
    const bool mine = valid[pass] && n <= q.cap && lane < n;
    if (mine) s_ids[pass * 32 + lane] = me;
    __syncwarp();
    if (mine) {
      int rank = 0;
      for (int k = 0; k < n; k++) rank += s_ids[pass * 32 + k] < me;
      s_geo[pass * 32 + rank] = q.geo ? g : geometry(pass, me);
    }
  }
  __syncthreads();
  for (int pass = 0; pass < 2; pass++) {
    if (!valid[pass]) continue;
    const uint8_t *res = res_[pass];
    auto addend = [&](int c, int ry, int rx) -> int {
      int r = 0;
      if (c == 0) r = (int)res[(long long)ry * q.X + rx] - 128;
      else if (ry < ch && rx < cw) r = (int)res[(c == 1 ? coff1 : coff2) + (long long)ry * cw + rx] - 128;
      return (r * jj) >> 8;
    };
    auto apply = [&](int n, const int4 *geo) {
      for (int k = 0; k < n; k++) {
        const int4 g = geo[k];
        const int y = ty - g.x, xb = tx0 - g.y;
        if ((unsigned)y < (unsigned)q.bs && xb > -4 && xb < q.bs) {
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const int x = xb + kk;
            if ((unsigned)x < (unsigned)q.bs) {
#pragma unroll
              for (int c = 0; c < 3; c++) cur[c][kk] = min(max(cur[c][kk] + addend(c, g.z + y, g.w + x), 0), 255);
            }
          }
        }
      }
    };
    if (total[pass] <= q.cap) {
      apply(total[pass], s_geo + pass * 32);
    } else {
      // overflow: ordered scan of every block within reach of the tile, 64 candidates per round
      const int pair = pass == 0 ? frame - 1 : frame, dir = pass == 0 ? 1 : 0;
      const int reach = q.reach[pair * 2 + dir];
      const int by_lo = max(0, (tile_y0 - reach - q.bs + 1 + (q.bs - 1) * (tile_y0 - reach - q.bs + 1 > 0)) / q.bs);
      const int by_hi = min(q.BY - 1, (tile_y1 + reach) / q.bs);
      const int bx_lo = max(0, (tile_x0 - reach - q.bs + 1 + (q.bs - 1) * (tile_x0 - reach - q.bs + 1 > 0)) / q.bs);
      const int bx_hi = min(q.BX - 1, (tile_x1 + reach) / q.bs);
      const int nbw = max(bx_hi - bx_lo + 1, 0), nbh = max(by_hi - by_lo + 1, 0);
      const int nblocks = nbw * nbh;
      for (int base = 0; base < nblocks; base += 64) {
        __syncthreads();
        bool hit = false;
        int4 g = make_int4(0, 0, 0, 0);
        if (base + (int)threadIdx.x < nblocks) {
          const int k = base + threadIdx.x;
          const int cby = by_lo + k / nbw, cbx = bx_lo + k % nbw;
          g = geometry(pass, cby * q.BX + cbx);
          const int fy0 = iclamp(g.x, 0, q.Y - 1), fy1 = iclamp(g.x + q.bs - 1, 0, q.Y - 1);
          const int fx0 = iclamp(g.y, 0, q.X - 1), fx1 = iclamp(g.y + q.bs - 1, 0, q.X - 1);
          hit = fy0 <= tile_y1 && fy1 >= tile_y0 && fx0 <= tile_x1 && fx1 >= tile_x0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        const int prefix = warp ? s_w[0] : 0;
        if (hit) s_scan[prefix + __popc(m & ((1u << lane) - 1))] = g;
        if (threadIdx.x == 63) s_count = prefix + __popc(m);
        __syncthreads();
        apply(s_count, s_scan);
      }
    }
  }
  {
    unsigned w = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) w |= (unsigned)cur[0][k] << (8 * k);  // bytes: untouched inputs or clamped sums
    *reinterpret_cast<unsigned *>(q.luma_out + (long long)frame * q.luma_out_stride + (unsigned)(ty * q.X + tx0)) = w;
#pragma unroll
    for (int c = 1; c < 3; c++) {
      uint2 v;
      v.x = ((unsigned)cur[c][0] & 0xffffu) | ((unsigned)cur[c][1] << 16);
      v.y = ((unsigned)cur[c][2] & 0xffffu) | ((unsigned)cur[c][3] << 16);
      *reinterpret_cast<uint2 *>(target[c - 1]) = v;
    }
  }
}

// ---- chroma 4:2:0 <-> luma-sized planes around the update (update.cpp:506-656) ----
// Zero-high-band synthesis (5_3.cpp:81-94 with h = 0; dwt2d.cpp:139-172: columns, then rows) of a byte
// component: T[2i] = c[i], T[2i+1] = (c[i] + c[i+1]) / 2, last odd row = c[last]; the same along the rows.
// One thread: source rows i, i + 1 and source columns 4g .. 4g + 4 -> output rows 2i, 2i + 1, columns 8g .. 8g + 7
// (two 16-byte stores).  X % 8 == 0.
__global__ void __launch_bounds__(256) k_chroma_up_s16(Plane dst, int slot0, const uint8_t *__restrict__ src,
                                                        long long frame_stride, long long comp_off, int f0, int Y,
                                                        int X) {
  const int hw = X >> 1, hh = Y >> 1, groups = hw >> 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= hh * groups) return;
  const int i = idx / groups, g = idx - i * groups;
  const uint8_t *c = src + (long long)(f0 + blockIdx.z) * frame_stride + comp_off;
  const uint8_t *ra = c + (unsigned)(i * hw + 4 * g), *rb = i < hh - 1 ? ra + hw : ra;  // last odd row = last source row
  const unsigned a4 = *reinterpret_cast<const unsigned *>(ra), b4 = *reinterpret_cast<const unsigned *>(rb);
  const int nx = 4 * g + 4 < hw ? 4 : 3;  // last odd column = last source column
  int te[5], to[5];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    te[k] = (a4 >> (8 * k)) & 0xff;
    to[k] = (te[k] + (int)((b4 >> (8 * k)) & 0xff)) >> 1;
  }
  te[4] = ra[nx];
  to[4] = (te[4] + (int)rb[nx]) >> 1;
  unsigned we[4], wo[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    we[k] = (unsigned)te[k] | ((unsigned)((te[k] + te[k + 1]) >> 1) << 16);
    wo[k] = (unsigned)to[k] | ((unsigned)((to[k] + to[k + 1]) >> 1) << 16);
  }
  *reinterpret_cast<uint4 *>(dst.row(slot0 + blockIdx.z, 2 * i) + 8 * g) = make_uint4(we[0], we[1], we[2], we[3]);
  *reinterpret_cast<uint4 *>(dst.row(slot0 + blockIdx.z, 2 * i + 1) + 8 * g) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
}

void launch_chroma_up_s16(const Launch &L, Plane dst, int slot0, int n, const uint8_t *src, long long frame_stride,
                          long long comp_off, int f0, int Y, int X) {
  if (n <= 0) return;
  ProfScope ps_(L, KC_IMG);
  const int work = (Y >> 1) * (X >> 3);
  k_chroma_up_s16<<<dim3((work + 255) / 256, 1, n), 256, 0, L.stream>>>(dst, slot0, src, frame_stride, comp_off, f0, Y, X);
  COUNT(L);
}

// LL band of one analysis level (dwt2d.cpp:76-119: rows, then columns; 5_3.cpp:39-52, even sizes) of a
// luma-sized int16 plane, stored as bytes.  One thread owns output column gi and walks down LL1_SEG output rows:
// the low-pass sample of an input row at that column is five samples = three aligned 32-bit loads (neighbouring
// threads share two of them through L1), the column pass is a streaming lifting step (last even row, last
// high-pass value carried in registers).  No shared memory, no barriers.
static constexpr int LL1_SEG = 32;
__global__ void __launch_bounds__(128) k_ll1_store_u8(Plane src, int slot0, uint8_t *__restrict__ dst,
                                                       long long frame_stride, long long comp_off, int f0, int Y,
                                                       int X) {
  const int halfx = X >> 1, halfy = Y >> 1;
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= halfx) return;
  const int slot = slot0 + blockIdx.z;
  const int j0 = blockIdx.y * LL1_SEG, j1 = min(j0 + LL1_SEG, halfy);
  // low-pass sample gi of input row y (row pass)
  auto rowlow = [&](int y) -> int {
    const unsigned *w = reinterpret_cast<const unsigned *>(src.row(slot, y) + 2 * gi);
    const unsigned c = w[0];
    const int s0 = (short)(c & 0xffffu), s1 = (int)c >> 16;
    int h;
    if (gi == halfx - 1) {
      h = (short)(s1 - s0);
    } else {
      const int s2 = (short)(w[1] & 0xffffu);
      h = (short)(s1 - tdiv2(s0 + s2));
    }
    if (gi == 0) return (short)(s0 + tdiv2(h));
    const unsigned p = w[-1];
    const int hp = (short)(((int)p >> 16) - tdiv2((int)(short)(p & 0xffffu) + s0));
    return (short)(s0 + tdiv4(h + hp));
  };
  int e = rowlow(2 * j0), hp = 0;
  if (j0 > 0) hp = (short)(rowlow(2 * j0 - 1) - tdiv2(rowlow(2 * j0 - 2) + e));
  uint8_t *out = dst + (long long)(f0 + blockIdx.z) * frame_stride + comp_off + gi;
  for (int j = j0; j < j1; j++) {
    const int t1 = rowlow(2 * j + 1);
    int h, e2 = 0;
    if (j == halfy - 1) {
      h = (short)(t1 - e);
    } else {
      e2 = rowlow(2 * j + 2);
      h = (short)(t1 - tdiv2(e + e2));
    }
    const int l = (short)(j == 0 ? e + tdiv2(h) : e + tdiv4(h + hp));
    out[(long long)j * halfx] = (uint8_t)l;  // truncation mod 256, no clamp
    hp = h;
    e = e2;
  }
}

void launch_ll1_store_u8(const Launch &L, Plane src, int slot0, int n, uint8_t *dst, long long frame_stride,
                         long long comp_off, int f0, int Y, int X) {
  if (n <= 0) return;
  ProfScope ps_(L, KC_DWT_ROWS);
  k_ll1_store_u8<<<dim3(((X >> 1) + 127) / 128, ((Y >> 1) + LL1_SEG - 1) / LL1_SEG, n), 128, 0, L.stream>>>(
      src, slot0, dst, frame_stride, comp_off, f0, Y, X);
  COUNT(L);
}

void launch_update_bin(const Launch &L, const UpdateBatchParams &q) {
  if (q.n_pairs <= 0) return;
  ProfScope ps_(L, KC_UPDATE);
  const int plane = q.BY * q.BX;
  k_update_bin<<<dim3((plane + 255) / 256, 2 * q.n_pairs), 256, 0, L.stream>>>(q);
  COUNT(L);
}

// uf = j / 256 with a small integral j: every product and sum of the reference's float chain is exact
static bool update_dyadic(float uf, int *j256) {
  const float j = uf * 256.0f;
  if (!(j == (float)(int)j) || j > 1024.0f || j < -1024.0f) return false;
  *j256 = (int)j;
  return true;
}

bool update_is_dyadic(float uf) {
  int j256;
  static const int allow = getenv("QSVC_UPDATE_DYADIC") ? atoi(getenv("QSVC_UPDATE_DYADIC")) : 1;
  return allow && update_dyadic(uf, &j256);
}

void launch_update_batch(const Launch &L, const UpdateBatchParams &q, int nframes) {
  if (nframes <= 0) return;
  ProfScope ps_(L, KC_UPDATE);
  int j256 = 0;
  if (q.cap <= 32 && update_is_dyadic(q.uf) && update_dyadic(q.uf, &j256)) {
    static const int quad = getenv("QSVC_UPDATE_QUAD") ? atoi(getenv("QSVC_UPDATE_QUAD")) : 1;
    if (quad && q.luma_in && q.tiles_x >= 3 && q.tiles_y >= 3 && q.X % 8 == 0) {
      k_update_quad<<<dim3(q.tiles_x - 2, q.tiles_y - 2, nframes), 64, 0, L.stream>>>(q, j256);
      COUNT(L);
      k_update_dyadic<<<dim3(2 * q.tiles_x + 2 * (q.tiles_y - 2), 1, nframes), 256, 0, L.stream>>>(q, j256, 1);
      COUNT(L);
      return;
    }
    k_update_dyadic<<<dim3(q.tiles_x, q.tiles_y, nframes), 256, 0, L.stream>>>(q, j256, 0);
    COUNT(L);
    return;
  }
  k_update_batch<<<dim3(q.tiles_x, q.tiles_y, nframes * 3), 256, 0, L.stream>>>(q);
  COUNT(L);
}

// largest |component| of the two planes (x, y) of one direction of a field
__global__ void k_mv_reach(const short *__restrict__ mv, int n, int *out) {
  __shared__ int s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  int m = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = max(m, abs((int)mv[i]));
  atomicMax(&s_max, m);
  __syncthreads();
  if (threadIdx.x == 0) *out = s_max;
}

void launch_mv_reach(const Launch &L, const short *mv, int n, int *out) {
  ProfScope ps_(L, KC_UPDATE);
  k_mv_reach<<<1, 256, 0, L.stream>>>(mv, n, out);
  COUNT(L);
}

void launch_update(const Launch &L, const UpdateParams &q) {
  dim3 grid((q.X + 15) / 16, (q.Y + 15) / 16, 3);
  ProfScope ps_(L, KC_UPDATE);
  k_update<<<grid, 256, 0, L.stream>>>(q);
  COUNT(L);
}
