// pipe_mix.cu -- which issue pipes do the packed-int16 SIMD instructions use on sm_100a?
// Two groups of independent accumulator chains run different ops; if the ops sit on different pipes
// the combined rate exceeds the single-op rate (64 thread-instr/clk/SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { VADD = 0, VMAX, LOP, PRMT, IMAD, IADD, VADDMAX, VMAX3, FFMA, VSUB, SHF };

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t o, uint32_t c)
{
  if (OP == VADD) return __vadd2(a, o);
  if (OP == VSUB) return __vsub2(a, o);
  if (OP == VMAX) return __vmaxs2(a, o);
  if (OP == LOP) return (a & o) ^ (c | a);
  if (OP == PRMT) return __byte_perm(a, o, 0x5432);
  if (OP == IMAD) return a * c + o;
  if (OP == IADD) return a + o;
  if (OP == VADDMAX) return __viaddmax_s16x2(a, o, c);
  if (OP == VMAX3) return __vimax3_s16x2(a, o, c);
  if (OP == FFMA) return __float_as_uint(fmaf(__uint_as_float(a), __uint_as_float(c), __uint_as_float(o)));
  if (OP == SHF) return __funnelshift_l(a, o, 3);
  return a;
}

template <int OPA, int OPB, int NA, int NB>
__global__ void k(uint32_t* out, int iters, uint32_t seed)
{
  uint32_t a[NA], b[NB], c = seed ^ 0x00030003u;
#pragma unroll
  for (int i = 0; i < NA; i++) a[i] = threadIdx.x * 65537u + i + seed;
#pragma unroll
  for (int i = 0; i < NB; i++) b[i] = threadIdx.x * 257u + i * 3 + seed;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < (NA > NB ? NA : NB); i++) {
        if (i < NA) a[i] = op<OPA>(a[i], a[(i + 1) % NA], c);
        if (i < NB) b[i] = op<OPB>(b[i], b[(i + 1) % NB], c);
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) r ^= a[i];
#pragma unroll
  for (int i = 0; i < NB; i++) r ^= b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OPA, int OPB, int NA, int NB>
void run(const char* name, int sms, double mhz)
{
  const int threads = 1024, blocks = sms * 2, iters = 2048;
  uint32_t* out;
  cudaMalloc(&out, blocks * threads * 4);
  k<OPA, OPB, NA, NB><<<blocks, threads>>>(out, 16, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OPA, OPB, NA, NB><<<blocks, threads>>>(out, iters, 3);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double total = (double)blocks * threads * iters * 8.0 * (NA + NB);
  printf("%-28s A:B=%d:%d  %.1f thread-instr/clk/SM (at %.0f MHz)  %.3f ms\n", name, NA, NB,
         total / (ms * 1e-3) / sms / (mhz * 1e6), mhz, ms);
  cudaFree(out);
}

int main()
{
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  run<VADD, VADD, 4, 4>("vadd only", sms, mhz);
  run<VMAX, VMAX, 4, 4>("vmax only", sms, mhz);
  run<VADD, VMAX, 4, 4>("vadd + vmax", sms, mhz);
  run<VADD, VMAX, 6, 3>("vadd + vmax", sms, mhz);
  run<VADD, VMAX, 3, 6>("vadd + vmax", sms, mhz);
  run<VADD, LOP, 4, 4>("vadd + lop3", sms, mhz);
  run<VADD, PRMT, 4, 4>("vadd + prmt", sms, mhz);
  run<VADD, IMAD, 4, 4>("vadd + imad", sms, mhz);
  run<VADD, FFMA, 4, 4>("vadd + ffma", sms, mhz);
  run<VMAX, IMAD, 4, 4>("vmax + imad", sms, mhz);
  run<VMAX, FFMA, 4, 4>("vmax + ffma", sms, mhz);
  run<VMAX, LOP, 4, 4>("vmax + lop3", sms, mhz);
  run<VMAX, PRMT, 4, 4>("vmax + prmt", sms, mhz);
  run<VADDMAX, VADDMAX, 4, 4>("vaddmax only", sms, mhz);
  run<VADDMAX, VADD, 4, 4>("vaddmax + vadd", sms, mhz);
  run<VADDMAX, VMAX, 4, 4>("vaddmax + vmax", sms, mhz);
  run<VMAX3, VADD, 4, 4>("vmax3 + vadd", sms, mhz);
  run<VMAX3, VMAX, 4, 4>("vmax3 + vmax", sms, mhz);
  run<VSUB, VMAX, 4, 4>("vsub + vmax", sms, mhz);
  run<IADD, VMAX, 4, 4>("iadd32 + vmax", sms, mhz);
  run<IADD, VADD, 4, 4>("iadd32 + vadd", sms, mhz);
  run<IMAD, FFMA, 4, 4>("imad + ffma", sms, mhz);
  run<LOP, PRMT, 4, 4>("lop3 + prmt", sms, mhz);
  run<SHF, VMAX, 4, 4>("shf + vmax", sms, mhz);
  run<SHF, VADD, 4, 4>("shf + vadd", sms, mhz);
  return 0;
}
