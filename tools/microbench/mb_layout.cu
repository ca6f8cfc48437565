// Development microbenchmark (not part of the product): what does the HBM deliver for the access pattern of the convolution
// store stream (read-modify-write of ROWS rows per HRU, one thread per HRU, CTAs of 128 HRUs, one member per grid.y)
//   layout S: state[member][row][nHp]          (row stride 80 KB; what the engine uses)
//   layout T: state[member][block][row][128]   (each CTA's rows contiguous)
// with 64-bit (1 HRU/thread) or 128-bit (2 HRUs/thread) accesses, R rows in flight, at a given occupancy (dynamic smem pads it).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_layout mb_layout.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
extern __shared__ double pad_[];

template <int R, bool TILED>
__global__ void __launch_bounds__(128) rmw64(double *st, int rows, int nHp, int nblk) {
  const int e = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
  double *p; size_t stride;
  if (TILED) { p = st + (((size_t)e * nblk + b) * rows) * 128 + t; stride = 128; }
  else { p = st + (size_t)e * rows * nHp + (size_t)b * 128 + t; stride = nHp; }
  double cur[R], acc = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++) cur[r] = __ldcs(p + (size_t)r * stride);
  for (int i = 0; i < rows; i += R) {
    double nxt[R];
#pragma unroll
    for (int r = 0; r < R; r++) nxt[r] = (i + R + r < rows) ? __ldcs(p + (size_t)(i + R + r) * stride) : 0.0;
#pragma unroll
    for (int r = 0; r < R; r++) if (i + r < rows) { const double v = cur[r] * 0.999 + acc * 1e-3; acc = v; __stcs(p + (size_t)(i + r) * stride, v); }
#pragma unroll
    for (int r = 0; r < R; r++) cur[r] = nxt[r];
  }
  if (acc == 123.456) pad_[0] = acc;
}
template <int R, bool TILED>
__global__ void __launch_bounds__(64) rmw128(double *st, int rows, int nHp, int nblk) {   // 64 threads x 2 HRUs = the same 128-HRU tile
  const int e = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
  double2 *p; size_t stride;
  if (TILED) { p = reinterpret_cast<double2 *>(st + (((size_t)e * nblk + b) * rows) * 128) + t; stride = 64; }
  else { p = reinterpret_cast<double2 *>(st + (size_t)e * rows * nHp + (size_t)b * 128) + t; stride = nHp / 2; }
  double2 cur[R]; double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++) cur[r] = __ldcs(p + (size_t)r * stride);
  for (int i = 0; i < rows; i += R) {
    double2 nxt[R];
#pragma unroll
    for (int r = 0; r < R; r++) nxt[r] = (i + R + r < rows) ? __ldcs(p + (size_t)(i + R + r) * stride) : make_double2(0, 0);
#pragma unroll
    for (int r = 0; r < R; r++) if (i + r < rows) { double2 v; v.x = cur[r].x * 0.999 + a0 * 1e-3; v.y = cur[r].y * 0.999 + a1 * 1e-3; a0 = v.x; a1 = v.y; __stcs(p + (size_t)(i + r) * stride, v); }
#pragma unroll
    for (int r = 0; r < R; r++) cur[r] = nxt[r];
  }
  if (a0 + a1 == 123.456) pad_[0] = a0;
}
__global__ void copyk(const double2 *a, double2 *b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <class K> float timeit(K launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; i++) launch();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError());
  return ms / reps;
}
int main() {
  const int E = 512, nblk = 79, nHp = nblk * 128, rows = 84;
  const size_t n = (size_t)E * rows * nHp;
  double *st, *st2; CK(cudaMalloc(&st, n * 8)); CK(cudaMalloc(&st2, n * 8)); CK(cudaMemset(st, 0, n * 8));
  const double gb = 2.0 * n * 8 / 1e9;
  dim3 grid(nblk, E);
  printf("bytes moved per launch %.3f GB (read + write)\n", gb);
  { float ms = timeit([&] { copyk<<<148 * 16, 512>>>((const double2 *)st, (double2 *)st2, n / 2); }, 5); printf("%-34s %.3f ms  %.0f GB/s\n", "plain copy (double2)", ms, gb / ms * 1e3); }
#define RUN64(R, T, SM, name) { CK(cudaFuncSetAttribute(rmw64<R, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); \
    float ms = timeit([&] { rmw64<R, T><<<grid, 128, SM>>>(st, rows, nHp, nblk); }, 5); printf("%-34s %.3f ms  %.0f GB/s\n", name, ms, gb / ms * 1e3); }
#define RUN128(R, T, SM, name) { CK(cudaFuncSetAttribute(rmw128<R, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); \
    float ms = timeit([&] { rmw128<R, T><<<grid, 64, SM>>>(st, rows, nHp, nblk); }, 5); printf("%-34s %.3f ms  %.0f GB/s\n", name, ms, gb / ms * 1e3); }
  const int SM7 = 31 * 1024;   // 7 CTAs of 128 threads per SM = 28 warps, the sweep's occupancy
  RUN64(4, false, SM7, "S 64-bit R=4 7cta/SM");
  RUN64(4, true, SM7, "T 64-bit R=4 7cta/SM");
  RUN64(8, false, SM7, "S 64-bit R=8 7cta/SM");
  RUN64(8, true, SM7, "T 64-bit R=8 7cta/SM");
  RUN64(4, false, 1024, "S 64-bit R=4 full occupancy");
  RUN64(4, true, 1024, "T 64-bit R=4 full occupancy");
  RUN64(8, false, 1024, "S 64-bit R=8 full occupancy");
  RUN64(8, true, 1024, "T 64-bit R=8 full occupancy");
  RUN128(4, false, SM7, "S 128-bit R=4 7cta/SM (14 warps)");
  RUN128(4, true, SM7, "T 128-bit R=4 7cta/SM (14 warps)");
  RUN128(4, false, 15 * 1024, "S 128-bit R=4 14cta/SM (28 warps)");
  RUN128(4, true, 15 * 1024, "T 128-bit R=4 14cta/SM (28 warps)");
  RUN128(8, false, 15 * 1024, "S 128-bit R=8 14cta/SM (28 warps)");
  RUN128(8, true, 15 * 1024, "T 128-bit R=8 14cta/SM (28 warps)");
  RUN128(4, true, 1024, "T 128-bit R=4 full occupancy");
  RUN128(4, false, 1024, "S 128-bit R=4 full occupancy");
  return 0;
}
