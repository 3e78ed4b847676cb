// pipe_probe.cu -- issue rates of the integer / packed instructions the MCTF kernels are made
// of, alone and in pairs, on one B200 (thread-instructions per clock per SM, from clock64()).
// Two instructions that share a pipe add their times; two that do not, overlap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/pipe_probe profiles/pipe_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

enum Op { VABS4, LOP3, PRMT, SHF, IADD3, IMAD, DP4A, HFMA2, VIADD2, VMNMX2, LEAHI, IMADHI, NOPS };
static const char *names[] = {"VABSDIFF4.U8.ACC", "LOP3", "PRMT", "SHF", "IADD3", "IMAD", "IDP.4A", "HFMA2",
                              "VIADD.16x2", "VIMNMX.S16x2", "LEA.HI", "IMAD.HI.U32"};

template <int OP>
__device__ __forceinline__ void one(unsigned &d, unsigned a, unsigned b) {
  if (OP == VABS4) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b));
  if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d) : "r"(a), "r"(b));
  if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
  if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
  if (OP == IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; xor.b32 %0, t, %2;}" : "+r"(d) : "r"(a), "r"(b));
  if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
  if (OP == DP4A) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(d) : "r"(a), "r"(b));
  if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
  if (OP == VIADD2) d = __vadd2(d, a);
  if (OP == VMNMX2) d = __vmaxs2(d, a) ^ b;
  if (OP == LEAHI) d = d + (d >> 30) + a;
  if (OP == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
}

// NA instructions of kind A and NB of kind B per iteration, on 8 independent chains
template <int A, int B, int NA, int NB>
__global__ void __launch_bounds__(1024) k_probe(unsigned *out, long long *cycles, int iters, unsigned a, unsigned b) {
  unsigned r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = threadIdx.x * 2654435761u + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
      for (int k = 0; k < NA; k++) one<A>(r[i], a, b);
#pragma unroll
      for (int k = 0; k < NB; k++) one<B>(r[i], b, a);
    }
  }
  const long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= r[i];
  if (s == 0x12345u) out[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int A, int B, int NA, int NB>
static double run(unsigned *d_out, long long *d_cyc, int nsm) {
  const int iters = 2000;
  k_probe<A, B, NA, NB><<<nsm, 1024>>>(d_out, d_cyc, 10, 3, 5);
  k_probe<A, B, NA, NB><<<nsm, 1024>>>(d_out, d_cyc, iters, 3, 5);
  cudaDeviceSynchronize();
  long long cyc[256];
  cudaMemcpy(cyc, d_cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < nsm; i++) mean += (double)cyc[i];
  mean /= nsm;
  return 1024.0 * iters * 8 * (NA + NB) / mean;  // thread-instructions per clock per SM
}

#define SOLO(OPK)                                                                          \
  printf("%-18s alone          : %7.1f thread-instr/clk/SM\n", names[OPK],                \
         run<OPK, OPK, 1, 0>(d_out, d_cyc, nsm));
#define PAIR(OA, OB)                                                                       \
  printf("%-18s + %-12s : %7.1f thread-instr/clk/SM (1:1 mix)\n", names[OA], names[OB],   \
         run<OA, OB, 1, 1>(d_out, d_cyc, nsm));

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  unsigned *d_out;
  long long *d_cyc;
  cudaMalloc(&d_out, 4096);
  cudaMalloc(&d_cyc, 256 * sizeof(long long));
  printf("%s, %d SMs; one CTA of 1024 threads per SM, 8 independent chains per thread\n", p.name, nsm);
  SOLO(VABS4) SOLO(LOP3) SOLO(PRMT) SOLO(SHF) SOLO(IADD3) SOLO(IMAD) SOLO(DP4A) SOLO(HFMA2) SOLO(VIADD2)
  SOLO(VMNMX2) SOLO(LEAHI) SOLO(IMADHI)
  PAIR(IMADHI, LOP3) PAIR(IMADHI, IMAD) PAIR(IMADHI, DP4A)
  PAIR(VABS4, LOP3) PAIR(VABS4, PRMT) PAIR(VABS4, SHF) PAIR(VABS4, IMAD) PAIR(VABS4, DP4A) PAIR(VABS4, HFMA2)
  PAIR(LOP3, IMAD) PAIR(LOP3, DP4A) PAIR(LOP3, HFMA2) PAIR(IMAD, DP4A) PAIR(IMAD, HFMA2) PAIR(PRMT, SHF)
  printf("%-18s x3 + IMAD x1     : %7.1f thread-instr/clk/SM\n", names[VABS4], run<VABS4, IMAD, 3, 1>(d_out, d_cyc, nsm));
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
