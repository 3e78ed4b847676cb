// kernels_dwt.cu -- in-place 2-D integer 5/3 transform in the reference's
// Mallat layout (dwt2d.cpp:76-175 over 5_3.cpp:39-115).
//
// Analysis of one level: rows first, then columns; synthesis: columns first,
// then rows.  Each line is staged in shared memory, so the in-place
// de-interleave ([lows | highs]) has no read/write hazard; a column CTA owns a
// strip of CW adjacent columns (32-byte or 64-byte row segments) over all rows.
#include "kernels.cuh"

#define COUNT(L) (++*(L).counter)

static const int kMaxDynSmem = 200 * 1024;

template <bool SYNTH>
__global__ void __launch_bounds__(256) k_dwt_rows(Plane p, int slot0, int ny, int nx) {
  extern __shared__ __align__(16) short sm[];
  const int slot = slot0 + blockIdx.z;
  const int nvec = nx >> 3;
  for (int y = blockIdx.x; y < ny; y += gridDim.x) {
    short *row = p.row(slot, y);
    const bool aligned = (((uintptr_t)row) & 15) == 0;
    if (aligned) {
      const uint4 *r4 = reinterpret_cast<const uint4 *>(row);
      uint4 *s4 = reinterpret_cast<uint4 *>(sm);
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) s4[i] = r4[i];
      for (int i = (nvec << 3) + threadIdx.x; i < nx; i += blockDim.x) sm[i] = row[i];
    } else {
      for (int i = threadIdx.x; i < nx; i += blockDim.x) sm[i] = row[i];
    }
    __syncthreads();
    if (aligned) {
      uint4 *r4 = reinterpret_cast<uint4 *>(row);
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        union { uint4 v; short h[8]; } u;
#pragma unroll
        for (int k = 0; k < 8; k++)
          u.h[k] = SYNTH ? l53_syn_out(sm, 1, 8 * i + k, nx) : l53_ana_out(sm, 1, 8 * i + k, nx);
        r4[i] = u.v;
      }
      for (int j = (nvec << 3) + threadIdx.x; j < nx; j += blockDim.x)
        row[j] = SYNTH ? l53_syn_out(sm, 1, j, nx) : l53_ana_out(sm, 1, j, nx);
    } else {
      for (int j = threadIdx.x; j < nx; j += blockDim.x)
        row[j] = SYNTH ? l53_syn_out(sm, 1, j, nx) : l53_ana_out(sm, 1, j, nx);
    }
    __syncthreads();
  }
}

// Column pass: the CTA stages a strip of CW = 2^cw_log2 columns x ny rows in
// shared memory, then every thread produces 8 horizontally adjacent outputs of
// one row (one 128-bit store).  1024 threads keep enough loads in flight.
template <bool SYNTH>
__global__ void __launch_bounds__(1024) k_dwt_cols(Plane p, int slot0, int ny, int nx, int cw_log2) {
  extern __shared__ __align__(16) short sm[];
  const int slot = slot0 + blockIdx.z;
  const int CW = 1 << cw_log2;
  const int x0 = blockIdx.x << cw_log2;
  const int cw = min(CW, nx - x0);
  if (cw_log2 >= 3) {
    const int upr_log2 = cw_log2 - 3;  // 8-column units per row
    const int units = ny << upr_log2;
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
      int y = u >> upr_log2, c0 = (u & ((1 << upr_log2) - 1)) << 3;
      if (c0 >= cw) continue;
      const short *src = p.row(slot, y) + x0 + c0;
      short *dst = sm + (y << cw_log2) + c0;
      if (c0 + 8 <= cw && (((uintptr_t)src) & 15) == 0) {
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src);
      } else {
        for (int k = 0; k < 8 && c0 + k < cw; k++) dst[k] = src[k];
      }
    }
    __syncthreads();
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
      int y = u >> upr_log2, c0 = (u & ((1 << upr_log2) - 1)) << 3;
      if (c0 >= cw) continue;
      short *dst = p.row(slot, y) + x0 + c0;
      union { uint4 v; short h[8]; } o;
#pragma unroll
      for (int k = 0; k < 8; k++)
        o.h[k] = SYNTH ? l53_syn_out(sm + c0 + k, CW, y, ny) : l53_ana_out(sm + c0 + k, CW, y, ny);
      if (c0 + 8 <= cw && (((uintptr_t)dst) & 15) == 0) {
        *reinterpret_cast<uint4 *>(dst) = o.v;
      } else {
        for (int k = 0; k < 8 && c0 + k < cw; k++) dst[k] = o.h[k];
      }
    }
  } else {
    const int total = ny << cw_log2;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      int y = idx >> cw_log2, c = idx & (CW - 1);
      if (c < cw) sm[idx] = p.row(slot, y)[x0 + c];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      int y = idx >> cw_log2, c = idx & (CW - 1);
      if (c < cw)
        p.row(slot, y)[x0 + c] =
            SYNTH ? l53_syn_out(sm + c, CW, y, ny) : l53_ana_out(sm + c, CW, y, ny);
    }
  }
}

int dwt_init_attributes() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_dwt_cols<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_dwt_cols<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_dwt_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_dwt_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  return (int)e;
}

static void launch_rows(const Launch &L, Plane p, int slot0, int nslots, int ny, int nx, bool synth) {
  dim3 grid(ny < 2048 ? ny : 2048, 1, nslots);
  size_t smem = (size_t)nx * sizeof(short);
  ProfScope ps_(L, KC_DWT_ROWS);
  if (synth)
    k_dwt_rows<true><<<grid, 256, smem, L.stream>>>(p, slot0, ny, nx);
  else
    k_dwt_rows<false><<<grid, 256, smem, L.stream>>>(p, slot0, ny, nx);
  COUNT(L);
}

static void launch_cols(const Launch &L, Plane p, int slot0, int nslots, int ny, int nx, bool synth) {
  int cw_log2 = 5;
  while (cw_log2 > 0 && ((size_t)ny << cw_log2) * sizeof(short) > (size_t)kMaxDynSmem) cw_log2--;
  // keep at least ~2 CTAs per SM worth of strips when the image is narrow
  while (cw_log2 > 3 && ((nx + (1 << cw_log2) - 1) >> cw_log2) * nslots < 296) cw_log2--;
  dim3 grid((nx + (1 << cw_log2) - 1) >> cw_log2, 1, nslots);
  size_t smem = ((size_t)ny << cw_log2) * sizeof(short);
  int threads = (ny << cw_log2) >= 8192 ? 1024 : 256;
  ProfScope ps_(L, KC_DWT_COLS);
  if (synth)
    k_dwt_cols<true><<<grid, threads, smem, L.stream>>>(p, slot0, ny, nx, cw_log2);
  else
    k_dwt_cols<false><<<grid, threads, smem, L.stream>>>(p, slot0, ny, nx, cw_log2);
  COUNT(L);
}

void launch_dwt_level(const Launch &L, Plane p, int slot0, int nslots, int ny, int nx, bool synth) {
  if (nslots <= 0) return;
  if (synth) {
    launch_cols(L, p, slot0, nslots, ny, nx, true);
    launch_rows(L, p, slot0, nslots, ny, nx, true);
  } else {
    launch_rows(L, p, slot0, nslots, ny, nx, false);
    launch_cols(L, p, slot0, nslots, ny, nx, false);
  }
}

// dwt2d.cpp:76-119
void dwt_analyze(const Launch &L, Plane p, int slot0, int nslots, int y, int x, int levels) {
  for (int lv = 0; lv < levels; lv++) {
    int nx = x, ny = y;
    x >>= 1;
    y >>= 1;
    if (y == 0) y = 1;
    if (x == 0) x = 1;
    launch_dwt_level(L, p, slot0, nslots, ny, nx, false);
  }
}

// dwt2d.cpp:128-175
void dwt_synthesize(const Launch &L, Plane p, int slot0, int nslots, int y, int x, int levels) {
  for (int lv = levels - 1; lv >= 0; lv--) {
    int nx = x >> lv, ny = y >> lv;
    if (nx == 0) nx = 1;
    if (ny == 0) ny = 1;
    launch_dwt_level(L, p, slot0, nslots, ny, nx, true);
  }
}
