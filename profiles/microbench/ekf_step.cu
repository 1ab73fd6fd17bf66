// Stand-alone timing harness of the EKF filter kernel (csrc/ekf.cuh) on a synthetic C4-shaped model
// (D=32 two-state subbands, N=3 three-state modulators, n=73): microseconds per time step and the cycles
// between the kernel's barriers (NSAGP_EKF_PROFILE).  Iterating on the kernel here takes seconds instead of
// the minutes of a full library build.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DNSAGP_EKF_PROFILE -o ekf_step ekf_step.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../nonstationary-audio-gp_b200/csrc/ekf.cuh"

using namespace nsagp;

int main(int argc, char** argv) {
  const int D = 32, N = 3, BZ = 2, BG = 3, BM = 3, M = D + N, n = D * BZ + N * BG;
  const long long T = argc > 1 ? atoll(argv[1]) : 20000;
  const int energy = argc > 2 ? atoi(argv[2]) : 0;
  std::vector<int> off(M + 1, 0);
  for (int i = 0; i < M; ++i) off[i + 1] = off[i] + (i < D ? BZ : BG);
  std::vector<double> A(M * BM * BM, 0.0), Q(M * BM * BM, 0.0), h(M * BM, 0.0), Pinf((size_t)n * n, 0.0), W(D * N), y(T);
  srand(1);
  auto rnd = [] { return rand() / (double)RAND_MAX; };
  for (int i = 0; i < M; ++i) {
    const int b = i < D ? BZ : BG;
    if (i < D) {                       // damped rotation
      const double w = 0.05 + 1.0 * rnd(), rho = 0.995;
      A[i * 9 + 0] = rho * cos(w); A[i * 9 + 3] = -rho * sin(w); A[i * 9 + 1] = rho * sin(w); A[i * 9 + 4] = rho * cos(w);
    } else {
      for (int r = 0; r < 3; ++r) { A[i * 9 + r + r * 3] = 0.99; if (r < 2) A[i * 9 + r + (r + 1) * 3] = 0.01; }
    }
    for (int r = 0; r < b; ++r) { Q[i * 9 + r + r * 3] = 1e-3; Pinf[(off[i] + r) + (size_t)(off[i] + r) * n] = i < D ? 0.01 : 1.0; }
    h[i * BM] = 1.0;
  }
  for (auto& w : W) w = 0.1 * rnd();
  for (auto& v : y) v = 0.1 * (rnd() - 0.5);
  int* d_off; double *dA, *dQ, *dP, *dh, *dW, *dy, *dMS, *dPS, *dm, *de; int* dstat; EkfArgs* dargs;
  cudaMalloc(&d_off, off.size() * 4); cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice);
#define UP(dst, v) cudaMalloc(&dst, v.size() * 8); cudaMemcpy(dst, v.data(), v.size() * 8, cudaMemcpyHostToDevice)
  UP(dA, A); UP(dQ, Q); UP(dP, Pinf); UP(dh, h); UP(dW, W); UP(dy, y);
  cudaMalloc(&dMS, T * n * 8); const long long ps_stride = ((long long)n * n + 1) & ~1LL; cudaMalloc(&dPS, (size_t)T * ps_stride * 8); cudaMalloc(&dm, n * 8); cudaMalloc(&de, 8); cudaMalloc(&dstat, 4);
  cudaMemset(dm, 0, n * 8); cudaMemset(dstat, 0, 4);
  EkfArgs a;
  a.D = D; a.N = N; a.M = M; a.n = n; a.BM = BM; a.T = T; a.off = d_off; a.A = dA; a.Q = dQ; a.Pinf = dP; a.h = dh; a.W = dW;
  a.sigma2 = 1e-2; a.y = dy; a.MS = dMS; a.PS = dPS; a.ps_stride = ps_stride; a.m_io = dm; a.edata = de; a.status = dstat;
  cudaMalloc(&dargs, sizeof(a)); cudaMemcpy(dargs, &a, sizeof(a), cudaMemcpyHostToDevice);
  const int staged = 1;
  const size_t sm = ((staged ? ps_stride : 0) + (size_t)n * n + 2 * M * BM * BM + 2 * M * BM + D * N + 6 * n + 64 + M + 3 * N + D + 16) * 8;
  auto kf = giekf_filter2_kernel<BZ, BG, BM>;
  cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    long long z[8] = {0};
#ifdef NSAGP_EKF_PROFILE
    cudaMemcpyToSymbol(g_ekf_prof, z, sizeof(z));
#endif
    cudaEventRecord(e0);
    kf<<<1, EkfF2<BM>::TH, sm>>>(dargs, 1, energy);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%s  T=%lld  %.3f ms  %.3f us/step", cudaGetErrorString(err), T, ms, ms * 1e3 / T);
#ifdef NSAGP_EKF_PROFILE
    cudaMemcpyFromSymbol(z, g_ekf_prof, sizeof(z));
    printf("   cycles/step: A||B %.0f  C %.0f  D %.0f | mean group %.0f  first tile warp %.0f  last tile warp %.0f", (double)z[0] / T,
           (double)z[1] / T, (double)z[2] / T, (double)z[3] / T, (double)z[4] / T, (double)z[5] / T);
#endif
    printf("\n");
  }
  std::vector<double> ms_h(n);
  cudaMemcpy(ms_h.data(), dMS + (T - 1) * n, n * 8, cudaMemcpyDeviceToHost);
  double cs = 0; for (double v : ms_h) cs += v;
  std::vector<double> ps_h((size_t)n * n);
  double cp = 0;
  for (long long kk : {0LL, T / 2, T - 1}) {
    cudaMemcpy(ps_h.data(), dPS + kk * ps_stride, (size_t)n * n * 8, cudaMemcpyDeviceToHost);
    for (double v : ps_h) cp += v;
  }
  printf("checksum of the last mean: %.15g   of three covariances: %.15g\n", cs, cp);
  return 0;
}
