// Dependent-issue latency of the FP64 / shuffle / shared-memory instructions the sequential
// ADF pass is made of, measured on the target GPU with clock64 around unrolled dependent chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run: ./lat
#include <cstdio>
#include <cuda_runtime.h>

#define N 2048

template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, int nwarps_active) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = a;
  __syncthreads();
  double x = a + threadIdx.x * 1e-9, y = b;
  unsigned addr = (threadIdx.x & 31);
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, y, y);
    if (OP == 1) x = x + y;
    if (OP == 2) x = x * y;
    if (OP == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 4);
    if (OP == 5) { addr = (unsigned)__double_as_longlong(sm[addr & 63]) & 63; }
    if (OP == 6) { x = fma(x, y, y); y = fma(y, x, x); }   // 2 per iteration, still dependent
    if (OP == 7) { int h = __double2hiint(x); x = __hiloint2double(h + 1, __double2loint(x)); }
    if (OP == 8) x = (x > y) ? y : x + 1.0;                 // DSETP + select path
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + addr;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Throughput: W warps on one SM, each with 4 independent DFMA chains.
__global__ void tput(double* out, long long* cyc, double a, double b) {
  double x0 = a + threadIdx.x, x1 = a * 2, x2 = a * 3, x3 = a * 4;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    x0 = fma(x0, b, b); x1 = fma(x1, b, b); x2 = fma(x2, b, b); x3 = fma(x3, b, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x0 + x1 + x2 + x3;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  long long h;
  const char* names[] = {"DFMA", "DADD", "DMUL", "MUFU.RCP64H", "SHFL.BFLY (64-bit = 2 SHFL)", "LDS.64 dependent address",
                         "2 dependent DFMA", "hi-word int add round trip", "DSETP+select+DADD"};
#define RUN(OP, TH)                                                                         \
  chain<OP><<<1, TH>>>(out, cyc, 1.0000001, 0.9999999, 1); cudaDeviceSynchronize();         \
  chain<OP><<<1, TH>>>(out, cyc, 1.0000001, 0.9999999, 1); cudaDeviceSynchronize();         \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                           \
  printf("%-32s threads=%4d  %.2f cycles/iter\n", names[OP], TH, (double)h / N);
  RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32) RUN(6, 32) RUN(7, 32) RUN(8, 32)
  RUN(0, 128) RUN(0, 256) RUN(0, 384) RUN(0, 512) RUN(0, 1024)
  for (int th = 32; th <= 1024; th *= 2) {
    tput<<<1, th>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize();
    tput<<<1, th>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("throughput: %4d threads x 4 chains: %.2f cycles per 4 DFMA per warp -> %.1f DFMA lanes/cycle/SM\n", th,
           (double)h / N, 4.0 * th / ((double)h / N));
  }
  return 0;
}
