// alu_probe: issue-rate microbenchmarks of the non-tensor FP32/INT pipes on sm_100a.
// Pins the roofline denominator of the score kernel: what one SM sub-partition
// (SMSP) can issue per clock for the instruction mixes the DP row is made of.
// Every timed operation is inline asm volatile (or checked in SASS) so ptxas cannot drop or fuse it.
// Output: one line per mode: warp-instructions per clock per SMSP (from clock64 inside the kernel
// and from CUDA-event time x nominal clock) and "lane-ops"/s counting min3 and f32x2 as two ops.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/alu_probe tools/alu_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define FADD(d, a, b) asm volatile("add.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define FMIN(d, a, b) asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define FMIN3(d, a, b, c) asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define FADD2(x, y) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y))
#define IMIN(d, a, b) asm volatile("min.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define UMIN(d, a, b) asm volatile("min.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define IADD(d, a, b) asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define LOP(d, a, b) asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define IMAD(d, a, b, c) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))

struct ModeInfo
{
  char const *name;
  int instr; // warp instructions per inner iteration (8 chains)
  int ops;   // "algorithmic" lane ops per inner iteration (min3, f32x2 = 2)
};

__host__ __device__ constexpr ModeInfo mode_info(int m)
{
  switch (m)
  {
  case 0: return {"FADD+FMNMX 1:1", 16, 16};
  case 1: return {"FADD", 16, 16};
  case 2: return {"FMNMX", 16, 16};
  case 3: return {"FMNMX3", 16, 32};
  case 4: return {"FADD2", 16, 32};
  case 5: return {"FADD+FMNMX3 2:1", 24, 32};
  case 6: return {"FADD2+FMNMX3 1:1", 32, 64};
  case 7: return {"VIMNMX (2-input s32 min)", 16, 16};
  case 8: return {"VIMNMX3 (fused from two min.s32)", 16, 32};
  case 9: return {"FADD+VIMNMX3 2:1", 24, 32};
  case 10: return {"IADD3", 16, 16};
  case 11: return {"FADD+IADD3 1:2", 24, 24};
  case 12: return {"LOP3+IADD3 1:1", 16, 16};
  case 13: return {"IMAD", 16, 16};
  case 14: return {"FADD+IMAD 1:1", 16, 16};
  case 15: return {"FADD+LOP3+IADD3 1:1:1", 24, 24};
  case 16: return {"FADD2+VIMNMX3 1:1", 32, 64};
  case 17: return {"FADD+FMNMX+FMNMX3 4:2:1 (DP row today)", 28, 32};
  case 18: return {"FMNMX+IADD3 1:1", 32, 32};
  case 19: return {"FADD2+FMNMX 1:1", 32, 48};
  default: return {"?", 1, 1};
  }
}
constexpr int NMODES = 20;

template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, long long *cyc, int iters, float seed)
{
  float a[8], b[8], c[8];
  int ia[8], ib[8], ic[8];
  unsigned long long x[8], y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
  {
    a[j] = seed + (float)(threadIdx.x + j);
    b[j] = seed * 0.5f + (float)j;
    c[j] = seed * 0.25f + (float)(j * 3);
    ia[j] = __float_as_int(a[j]);
    ib[j] = __float_as_int(b[j]);
    ic[j] = __float_as_int(c[j]);
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x[j]) : "f"(a[j]), "f"(c[j]));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y[j]) : "f"(b[j]), "f"(b[j]));
  }
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
  {
    // every intermediate has two readers, so ptxas cannot fuse two 2-input mins into a min3
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
      constexpr int dummy = 0; (void)dummy;
      int const j1 = (j + 1) & 7, j3 = (j + 3) & 7;
      if (MODE == 0) { FADD(a[j], c[j], b[j]); FMIN(c[j], c[j], a[j]); }
      if (MODE == 1) { FADD(a[j], a[j], b[j]); FADD(c[j], c[j], b[j]); }
      if (MODE == 2) { FMIN(a[j], a[j], b[j1]); FMIN(b[j], b[j], a[j3]); }
      if (MODE == 3) { FMIN3(a[j], a[j], b[j], c[j]); FMIN3(b[j], b[j], c[j], a[j]); }
      if (MODE == 4) { FADD2(x[j], y[j]); FADD2(y[j], x[j]); }
      if (MODE == 5) { FADD(a[j], a[j], c[j]); FADD(b[j], b[j], c[j]); FMIN3(c[j], c[j], a[j], b[j]); }
      if (MODE == 6) { FADD2(x[j], y[j]); FMIN3(c[j], c[j], a[j], b[j1]); FADD2(y[j], x[j]); FMIN3(a[j], a[j], b[j], c[j1]); }
      if (MODE == 7) { IMIN(ia[j], ia[j], ib[j1]); IMIN(ib[j], ib[j], ia[j3]); }
      if (MODE == 8) { int t, u; IMIN(t, ia[j], ib[j1]); IMIN(ia[j], t, ic[j]); IMIN(u, ib[j], ic[j1]); IMIN(ib[j], u, ia[j3]); }
      if (MODE == 9) { int t; FADD(a[j], a[j], __int_as_float(ic[j])); FADD(b[j], b[j], c[j]); IMIN(t, ic[j], __float_as_int(a[j])); IMIN(ic[j], t, __float_as_int(b[j])); }
      if (MODE == 10) { IADD(ia[j], ia[j], ib[j1]); IADD(ib[j], ib[j], ia[j3]); }
      if (MODE == 11) { FADD(a[j], a[j], b[j]); IADD(ic[j], ic[j], ia[j1]); IADD(ia[j], ia[j], ic[j]); }
      if (MODE == 12) { LOP(ia[j], ia[j], ib[j1]); IADD(ib[j], ib[j], ia[j3]); }
      if (MODE == 13) { IMAD(ia[j], ia[j], ib[j], ic[j]); IMAD(ic[j], ic[j], ib[j], ia[j]); }
      if (MODE == 14) { FADD(a[j], a[j], b[j]); IMAD(ic[j], ic[j], ib[j], ia[j]); }
      if (MODE == 15) { FADD(a[j], a[j], b[j]); LOP(ic[j], ic[j], ia[j1]); IADD(ia[j], ia[j], ic[j]); }
      if (MODE == 16) { int t; FADD2(x[j], y[j]); IMIN(t, ic[j], ia[j1]); IMIN(ic[j], t, ib[j]); FADD2(y[j], x[j]); IMIN(t, ia[j], ic[j1]); IMIN(ia[j], t, ib[j]); }
      if (MODE == 18) { FMIN(a[j], a[j], b[j1]); FMIN(b[j], b[j], a[j3]); IADD(ic[j], ic[j], ia[j1]); IADD(ia[j], ia[j], ic[j]); }
      if (MODE == 19) { FADD2(x[j], y[j]); FMIN(a[j], a[j], b[j1]); FADD2(y[j], x[j]); FMIN(b[j], b[j], a[j3]); }
    }
    if (MODE == 17)
    { // 16 FADD : 8 FMNMX : 4 FMNMX3 -- the DP row's mix today
#pragma unroll
      for (int j = 0; j < 8; j += 2)
      {
        FADD(a[j], c[j], b[j]); FADD(a[j + 1], c[j + 1], b[j + 1]);
        FADD(b[j], b[j], a[j]); FADD(b[j + 1], b[j + 1], a[j + 1]);
        FMIN(c[j], c[j], a[j]); FMIN(c[j + 1], c[j + 1], a[j + 1]);
        FMIN3(c[j], c[j], b[j], b[j + 1]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
  {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[j]));
    float lo2, hi2;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo2), "=f"(hi2) : "l"(y[j]));
    s += a[j] + b[j] + c[j] + lo + hi + lo2 + hi2 + __int_as_float(ia[j] ^ ib[j] ^ ic[j]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(float *d, long long *dc, int sms, int blocks_per_sm, double ghz)
{
  int const iters = 8192, threads = 256, blocks = sms * blocks_per_sm;
  ModeInfo const mi = mode_info(MODE);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep)
  {
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(d, dc, iters, 1.0f + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  long long *hc = (long long *)malloc(sizeof(long long) * blocks);
  cudaMemcpy(hc, dc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double cyc = 0;
  for (int i = 0; i < blocks; ++i) cyc += (double)hc[i];
  cyc /= blocks;
  free(hc);
  // warps per SMSP: blocks_per_sm * 8 warps / 4 SMSPs
  double const warps_per_smsp = blocks_per_sm * (threads / 32) / 4.0;
  double const instr_per_warp = (double)iters * mi.instr;
  double const ipc_clock = instr_per_warp * warps_per_smsp / cyc; // all blocks resident at once
  double const total_instr = instr_per_warp * (double)blocks * (threads / 32);
  double const ipc_event = total_instr / (best * 1e-3) / (sms * 4.0) / (ghz * 1e9);
  double const tops = (double)iters * mi.ops * (double)blocks * threads / (best * 1e-3) / 1e12;
  printf("{\"mode\": %d, \"mix\": \"%s\", \"ms\": %.4f, \"ipc_smsp_clock64\": %.4f, \"ipc_smsp_event_at_%.3fGHz\": %.4f, "
         "\"tera_lane_ops\": %.2f, \"blocks_per_sm\": %d}\n",
         MODE, mi.name, best, ipc_clock, ghz, ipc_event, tops, blocks_per_sm);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

template <int M>
static void run_all(float *d, long long *dc, int sms, int bps, double ghz)
{
  run<M>(d, dc, sms, bps, ghz);
  if constexpr (M + 1 < NMODES) run_all<M + 1>(d, dc, sms, bps, ghz);
}

int main(int argc, char **argv)
{
  int bps = argc > 1 ? atoi(argv[1]) : 2; // resident blocks of 256 threads per SM (2 => 4 warps per SMSP)
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { fprintf(stderr, "no device\n"); return 1; }
  double ghz = p.clockRate * 1e-6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
  float *d;
  long long *dc;
  cudaMalloc(&d, sizeof(float) * 256 * p.multiProcessorCount * 8);
  cudaMalloc(&dc, sizeof(long long) * p.multiProcessorCount * 8);
  run_all<0>(d, dc, p.multiProcessorCount, bps, ghz);
  cudaFree(d);
  cudaFree(dc);
  return cudaGetLastError() != cudaSuccess;
}
