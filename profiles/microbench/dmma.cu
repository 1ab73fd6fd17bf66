// FP64 tensor-core (DMMA) latency / throughput on the target GPU, for the shapes PTX offers:
// mma.sync m8n8k4 (1 a, 1 b, 2 c registers per thread), m16n8k4, m16n8k8, m16n8k16.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu ; run: ./dmma
#include <cstdio>
#include <cuda_runtime.h>

#define N 1024

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// SHAPE 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16.  CH independent accumulator chains per warp.
template <int SHAPE, int CH>
__global__ void bench(double* out, long long* cyc, double x, double y) {
  double c[CH][4], a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = x + i * 1e-3 + threadIdx.x * 1e-6;
  for (int i = 0; i < 4; ++i) b[i] = y + i * 1e-3;
  for (int j = 0; j < CH; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < N; ++it) {
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (SHAPE == 0) mma884(c[j][0], c[j][1], a[0], b[0]);
      if (SHAPE == 1) mma1684(c[j], a, b[0]);
      if (SHAPE == 2) mma1688(c[j], a, b);
      if (SHAPE == 3) mma16816(c[j], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < CH; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int SHAPE, int CH>
void run(const char* name, int fma_per_mma, double* out, long long* cyc) {
  for (int th : {32, 128, 256, 320, 512, 1024}) {
    long long h;
    bench<SHAPE, CH><<<1, th>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize();
    bench<SHAPE, CH><<<1, th>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_iter = (double)h / N;
    printf("%-10s chains=%d threads=%4d  %.2f cycles per %d mma per warp -> %.1f FMA/cycle/SM\n", name, CH, th, per_iter, CH,
           (double)fma_per_mma * CH * (th / 32) / per_iter);
  }
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  run<0, 1>("m8n8k4", 256, out, cyc);
  run<0, 8>("m8n8k4", 256, out, cyc);
  run<1, 1>("m16n8k4", 512, out, cyc);
  run<1, 4>("m16n8k4", 512, out, cyc);
  run<2, 1>("m16n8k8", 1024, out, cyc);
  run<2, 4>("m16n8k8", 1024, out, cyc);
  run<3, 1>("m16n8k16", 2048, out, cyc);
  run<3, 4>("m16n8k16", 2048, out, cyc);
  // whole-chip sustained rate with the best shape: 148*k CTAs
  return 0;
}
