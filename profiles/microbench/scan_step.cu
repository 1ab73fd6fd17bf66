// Stand-alone timing harness of the dense-scan compose / apply kernels (csrc/ekfscan.cuh) on random data of the
// C4 shape (n = 73, padded to 80): ms per segment and the FP64 tensor-core rate, without a full library build.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scan_step scan_step.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../nonstationary-audio-gp_b200/csrc/ekfscan.cuh"

using namespace nsagp;

int main(int argc, char** argv) {
  const int n = 73, NP = 80, M = 35;
  const int chunk_len = argc > 1 ? atoi(argv[1]) : 64, nchunks = argc > 2 ? atoi(argv[2]) : 296;
  const long long steps = (long long)chunk_len * nchunks;
  const size_t nn = (size_t)n * n, pss = (nn + 1) & ~(size_t)1;
  std::vector<double> G(steps * nn), L(steps * pss), g(steps * n);
  srand(2);
  for (auto& v : G) v = (rand() / (double)RAND_MAX - 0.5) * 0.02;
  for (auto& v : L) v = (rand() / (double)RAND_MAX) * 1e-3;
  for (auto& v : g) v = rand() / (double)RAND_MAX;
  DsArgs a;
  memset(&a, 0, sizeof(a));
  std::vector<int> off(M + 1);
  for (int i = 0; i <= M; ++i) off[i] = i < 32 ? 2 * i : 64 + 3 * (i - 32);
  int* d_off; cudaMalloc(&d_off, off.size() * 4); cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice);
  double* dh; cudaMalloc(&dh, M * 3 * 8); cudaMemset(dh, 0, M * 3 * 8);
  a.ekf.n = n; a.ekf.M = M; a.ekf.BM = 3; a.ekf.T = steps + 1; a.ekf.ps_stride = (long long)pss; a.ekf.off = d_off; a.ekf.h = dh;
  cudaMalloc(&a.ekf.PS, (steps + 1) * pss * 8); cudaMemcpy(a.ekf.PS, L.data(), L.size() * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&a.ekf.MS, (steps + 1) * n * 8);
  cudaMalloc(&a.Gt, G.size() * 8); cudaMemcpy(a.Gt, G.data(), G.size() * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&a.gv, g.size() * 8); cudaMemcpy(a.gv, g.data(), g.size() * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&a.aggE, nchunks * nn * 8); cudaMalloc(&a.aggL, nchunks * nn * 8); cudaMalloc(&a.aggg, nchunks * n * 8);
  cudaMalloc(&a.entP, nchunks * nn * 8); cudaMalloc(&a.entm, nchunks * n * 8); cudaMalloc(&a.carryP, nn * 8); cudaMalloc(&a.carrym, n * 8);
  cudaMemset(a.entP, 0, nchunks * nn * 8); cudaMemset(a.entm, 0, nchunks * n * 8);
  cudaMalloc(&a.EV, (steps + 1) * 2 * M * 8); cudaMemset(a.EV, 0, (steps + 1) * 2 * M * 8);
  cudaMalloc(&a.maxdiff, 8);
  a.seg_k0 = 0; a.seg_k1 = steps; a.chunk_len = chunk_len;
  const size_t mat = (size_t)(NP + 4) * NP;
  const size_t sm_co = (4 * mat + 3 * NP) * 8, sm_ap = (3 * mat + 3 * NP) * 8;
  auto k2 = ds_compose_kernel<NP>; auto k4 = ds_apply_kernel<NP>;
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_co);
  cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_ap);
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  const int grid = nchunks < 148 ? nchunks : 148;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k2<<<grid, kDsThreads, sm_co>>>(a);
    cudaEventRecord(e1);
    k4<<<grid, kDsThreads, sm_ap>>>(a);
    cudaEventRecord(e2);
    cudaError_t err = cudaDeviceSynchronize();
    float m1, m2; cudaEventElapsedTime(&m1, e0, e1); cudaEventElapsedTime(&m2, e1, e2);
    const double f1 = (double)nchunks * (chunk_len - 1) * 3 * 2.0 * NP * NP * NP, f2 = (double)steps * 2 * 2.0 * NP * NP * NP;
    printf("%s  compose %.3f ms (%.1f TFLOP/s padded)   apply %.3f ms (%.1f TFLOP/s padded)\n", cudaGetErrorString(err), m1,
           f1 / m1 / 1e9, m2, f2 / m2 / 1e9);
  }
  std::vector<double> out(nn);
  cudaMemcpy(out.data(), a.aggL, nn * 8, cudaMemcpyDeviceToHost);
  double cs = 0; for (double v : out) cs += v;
  printf("checksum aggL[0]: %.15g\n", cs);
  return 0;
}
