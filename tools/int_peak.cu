// int_peak.cu -- micro-benchmark of the sm_100a packed-int16 SIMD instruction issue rates
// (the integer-pipe roofline denominator for the turbo decoder; SURVEY.md 8d "Integer peak").
// Prints thread-instructions per clock per SM for each op (x2 = int16 lane-ops).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { VADD, VMAX, VADDMAX, VMAX3, SADD, LOP, MIX_ADD_MAX, PRMT, IADD32, IMAD32, MIX_VADD_IMAD, SHFL, MIX_MAX_LOP };

template <int OP, int ILP>
__global__ void k(uint32_t* out, int iters, unsigned long long* cyc, uint32_t seed)
{
  uint32_t a[ILP], b = seed * 0x01010101u + threadIdx.x, c = seed ^ 0x00030003u;
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 65537u + i;
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        const uint32_t o = a[(i + 1) % ILP];  // neighbour accumulator: nothing the compiler can fold
        if (OP == VADD) a[i] = __vadd2(a[i], o);
        if (OP == VMAX) a[i] = (u & 1) ? __vmaxs2(a[i], o) : __vmins2(a[i], o);
        if (OP == VADDMAX) a[i] = __viaddmax_s16x2(a[i], o, c);
        if (OP == VMAX3) a[i] = __vimax3_s16x2(a[i], o, c);
        if (OP == SADD) a[i] = __vaddss2(a[i], o);
        if (OP == LOP) a[i] = (a[i] & o) ^ (c | a[i]);
        if (OP == MIX_ADD_MAX) a[i] = (u & 1) ? __vadd2(a[i], o) : __vmaxs2(a[i], o);
        if (OP == PRMT) a[i] = __byte_perm(a[i], o, 0x5432);
        if (OP == IADD32) a[i] = a[i] + o;
        if (OP == IMAD32) a[i] = a[i] * b + o;
        if (OP == MIX_VADD_IMAD) a[i] = (u & 1) ? __vadd2(a[i], o) : (a[i] * b + o);
        if (OP == SHFL) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1) + o;
        if (OP == MIX_MAX_LOP) a[i] = (u & 1) ? __vmaxs2(a[i], o) : ((a[i] & o) ^ (c | a[i]));
      }
    }
  }
  unsigned long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int ILP>
void run(const char* name, int sms)
{
  const int threads = 1024, blocks = sms * 2, iters = 4096;
  uint32_t* out; unsigned long long* cyc;
  cudaMalloc(&out, blocks * threads * 4);
  cudaMalloc(&cyc, blocks * 8);
  k<OP, ILP><<<blocks, threads>>>(out, 16, cyc, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP, ILP><<<blocks, threads>>>(out, iters, cyc, 3);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long* h = new unsigned long long[blocks];
  cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += (double)h[i]; avg /= blocks;
  // two 1024-thread blocks share an SM: per-SM thread-instr = 2*1024*ops
  const double ops_thread = (double)iters * 8 * ILP;
  const double per_clk_sm = 2.0 * threads * ops_thread / avg;  // both resident blocks run concurrently for ~avg cycles
  const double total = (double)blocks * threads * ops_thread;
  printf("%-14s ILP=%d  %.1f thread-instr/clk/SM (clock64)   %.2f T thread-instr/s (events, %.3f ms)   eff clk %.0f MHz\n",
         name, ILP, per_clk_sm, total / (ms * 1e-3) / 1e12, ms, total / (ms * 1e-3) / (per_clk_sm * sms) / 1e6);
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main()
{
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device: %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  run<VADD, 8>("VIADD.16x2", sms);
  run<VMAX, 8>("VIMNMX.S16x2", sms);
  run<VADDMAX, 8>("VIADDMNMX", sms);
  run<VMAX3, 8>("VIMNMX3", sms);
  run<SADD, 8>("__vaddss2", sms);
  run<LOP, 8>("LOP3", sms);
  run<PRMT, 8>("PRMT", sms);
  run<IADD32, 8>("IADD32", sms);
  run<IMAD32, 8>("IMAD32", sms);
  run<MIX_ADD_MAX, 8>("mix add/max", sms);
  run<MIX_VADD_IMAD, 8>("mix vadd/imad", sms);
  run<MIX_MAX_LOP, 8>("mix max/lop", sms);
  run<SHFL, 8>("SHFL", sms);
  run<VADD, 2>("VIADD.16x2", sms);
  run<VMAX, 2>("VIMNMX.S16x2", sms);
  return 0;
}
