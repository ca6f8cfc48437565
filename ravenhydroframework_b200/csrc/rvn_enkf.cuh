// rvn_enkf.cuh -- EnKF analysis on the device (reference CEnKFEnsemble, src/EnKF.cpp).
//
//   enkf_pack_kernel    AddToStateMatrix       (src/EnKF.cpp:679-747): member state -> row of the state matrix X
//   enkf_update_kernel  AssimilationCalcs      (src/EnKF.cpp:478-596), the O(M N^2) part:
//                           A = X - mean(X);  X_delta = (A . Z) * 1/(N-1);  Xa = X + X_delta
//                       The N x Nobs part (HA, P, the N solves, Z = HA'.MM) is tiny and is done by the host layer
//                       (host/rvh_enkf.cpp); Z arrives as an N x N matrix.
//   enkf_unpack_kernel  UpdateFromStateMatrix  (src/EnKF.cpp:602-673): max(Xa, 0) -> member state
//
// Layout: X is [member][row] (a member's row is contiguous), exactly the reference's _state_matrix[e][i]; on several GPUs it
// is the NCCL all-gather of every rank's local member rows, so the gathered buffer needs no transposition.
// The update kernel gives one thread one state row k (coalesced over k for every member i) and a tile of ET local members;
// the sum over the N ensemble members runs in ascending i with separate multiply and add (-fmad=false), which is the
// reference's MatMult order (src/Matrix.cpp:38-49) -> bit-identical X_delta for identical Z.  Bound: HBM/L2 streaming of X
// (N*M*8 bytes read N_local/ET times) and the FP64 pipe (2*N*M*N_local flop).
#ifndef RVN_ENKF_CUH
#define RVN_ENKF_CUH
#include "rvn_device.cuh"

struct EnkfMap {
  const rvn_enkf_row *rows;   // [M]
  long long M;
  int hist_t;                 // pushes into the basin history rings since they were last in reference order
};

RVN_DI double *enkf_locate(const DevModel &Mo, const EnkfMap &mp, const int e, const long long r) {
  const rvn_enkf_row row = mp.rows[r];
  if (row.kind == RVN_XROW_HRU_SV) return Mo.state + RVN_STATE_OFF(Mo, e, row.sub, row.index);
  const int p = row.index;
  if (row.kind == RVN_XROW_QIN_HIST) {
    const int nQ = Mo.sb_nqin[p];
    int base = nQ - (mp.hist_t % nQ); if (base == nQ) base = 0;   // same ring arithmetic as struct Ring (rvn_kernels.cuh)
    int i = base + row.sub; if (i >= nQ) i -= nQ;
    return Mo.qin + ((size_t)e * Mo.nSB + p) * Mo.max_nqin + i;
  }
  if (row.kind == RVN_XROW_QOUT) return Mo.qout + ((size_t)e * Mo.nSB + p) * RVN_MAX_RIVER_SEGS + row.sub;
  return Mo.scal + ((size_t)e * Mo.nSB + p) * 8 + 0;   // RVN_XROW_QOUT_LAST
}

__global__ void enkf_pack_kernel(const __grid_constant__ DevModel Mo, const EnkfMap mp, double *__restrict__ X) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;
  if (r >= mp.M) return;
  X[(size_t)e * mp.M + r] = *enkf_locate(Mo, mp, e, r);
}
// X points at the first LOCAL member's row
__global__ void enkf_unpack_kernel(const __grid_constant__ DevModel Mo, const EnkfMap mp, const double *__restrict__ X) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;
  if (r >= mp.M) return;
  const double v = X[(size_t)e * mp.M + r];
  *enkf_locate(Mo, mp, e, r) = (v >= 0.0) ? v : 0.0;   // max(_state_matrix[e][ii], 0.0)
}

#define RVN_ENKF_BS 128
#define RVN_ENKF_ET 16   // local members per CTA tile (accumulators per column)
#define RVN_ENKF_KT 2    // state rows (columns of X) per thread
#define RVN_ENKF_U 8     // ensemble members fetched ahead
// Xmean[k] = sum_i X[i][k]/N accumulated in member order (src/EnKF.cpp:539-546), once per column (the FP64 division costs ~25
// instructions; inside the update kernel it was repeated for every member tile and took 65 % of its stall samples).
// N a power of two: x/N == x*(1/N) exactly (one rounding of the same real number), so the division becomes a multiply.
__global__ void __launch_bounds__(256) enkf_mean_kernel(const double *__restrict__ Xall, double *__restrict__ Xmean, const long long M, const int N) {
  const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
  if (k >= M) return;
  const bool pow2 = (N & (N - 1)) == 0;
  const double rN = 1.0 / N;
  double m = 0.0;
  int i = 0;
  for (; i + 8 <= N; i += 8) {
    double x[8];
#pragma unroll
    for (int u = 0; u < 8; u++) x[u] = __ldg(Xall + (size_t)(i + u) * M + k);
#pragma unroll
    for (int u = 0; u < 8; u++) m += pow2 ? x[u] * rN : x[u] / N;
  }
  for (; i < N; i++) { const double x = __ldg(Xall + (size_t)i * M + k); m += pow2 ? x * rN : x / N; }
  Xmean[k] = m;
}
// Xall [N][M] in; Xa [nLocal][M] out (separate buffer: every CTA reads all members' rows).  Zt = Z restricted to the local
// columns, stored [N][nLocalPad] with nLocalPad a multiple of ET (padding columns zero).  Per ensemble member i a thread loads
// KT values of X (coalesced) and ET values of Z (128-bit shared-memory broadcasts) for 2*KT*ET FP64 operations.
template <bool CHECK>
RVN_DI void enkf_mac_chunk(const double *__restrict__ Xall, const double *s_z, const long long M, const int N, const int i0, const long long (&kk)[RVN_ENKF_KT],
                           const double (&Xmean)[RVN_ENKF_KT], double (&acc)[RVN_ENKF_KT][RVN_ENKF_ET]) {
  double xv[RVN_ENKF_U][RVN_ENKF_KT];   // U*KT independent loads in flight per thread
#pragma unroll
  for (int u = 0; u < RVN_ENKF_U; u++)
#pragma unroll
    for (int j = 0; j < RVN_ENKF_KT; j++) xv[u][j] = (!CHECK || i0 + u < N) ? __ldg(Xall + (size_t)(i0 + u) * M + kk[j]) : 0.0;
#pragma unroll
  for (int u = 0; u < RVN_ENKF_U; u++) {
    if (!CHECK || i0 + u < N) {
      double a[RVN_ENKF_KT];
#pragma unroll
      for (int j = 0; j < RVN_ENKF_KT; j++) a[j] = xv[u][j] - Xmean[j];   // A = X - 1/N*(X*e_N1)*e_1N
      const double2 *z2 = reinterpret_cast<const double2 *>(s_z + (i0 + u) * RVN_ENKF_ET);
#pragma unroll
      for (int c = 0; c < RVN_ENKF_ET / 2; c++) {
        const double2 z = z2[c];
#pragma unroll
        for (int j = 0; j < RVN_ENKF_KT; j++) {   // MatMult(A,Z,...): C[n][p] += A[n][m]*B[m][p], m ascending, separate multiply and add
          acc[j][2 * c] += a[j] * z.x;
          acc[j][2 * c + 1] += a[j] * z.y;
        }
      }
    }
  }
}
__global__ void __launch_bounds__(RVN_ENKF_BS) enkf_update_kernel(const double *__restrict__ Xall, const double *__restrict__ Zt, const double *__restrict__ mean,
                                                                 double *__restrict__ Xa, const long long M, const int N, const int e0, const int nLocal,
                                                                 const int nLocalPad) {
  extern __shared__ __align__(16) double s_z[];   // [N][ET]
  const int tile = blockIdx.y;
  for (int i = threadIdx.x; i < N * RVN_ENKF_ET; i += RVN_ENKF_BS) {
    const int row = i / RVN_ENKF_ET, col = i - row * RVN_ENKF_ET;
    s_z[i] = Zt[(size_t)row * nLocalPad + tile * RVN_ENKF_ET + col];
  }
  __syncthreads();
  const long long k0 = (long long)blockIdx.x * (RVN_ENKF_BS * RVN_ENKF_KT) + threadIdx.x;
  bool live[RVN_ENKF_KT];
  long long kk[RVN_ENKF_KT];
  double Xmean[RVN_ENKF_KT];
#pragma unroll
  for (int j = 0; j < RVN_ENKF_KT; j++) {
    kk[j] = k0 + (long long)j * RVN_ENKF_BS; live[j] = kk[j] < M; if (!live[j]) kk[j] = M - 1;
    Xmean[j] = __ldg(mean + kk[j]);
  }
  double acc[RVN_ENKF_KT][RVN_ENKF_ET];
#pragma unroll
  for (int j = 0; j < RVN_ENKF_KT; j++)
#pragma unroll
    for (int c = 0; c < RVN_ENKF_ET; c++) acc[j][c] = 0.0;
  int i0 = 0;
  for (; i0 + RVN_ENKF_U <= N; i0 += RVN_ENKF_U) enkf_mac_chunk<false>(Xall, s_z, M, N, i0, kk, Xmean, acc);
  if (i0 < N) enkf_mac_chunk<true>(Xall, s_z, M, N, i0, kk, Xmean, acc);
  const double scale = 1.0 / (N - 1);
#pragma unroll
  for (int j = 0; j < RVN_ENKF_KT; j++) {
    if (!live[j]) continue;
#pragma unroll
    for (int c = 0; c < RVN_ENKF_ET; c++) {
      const int el = tile * RVN_ENKF_ET + c;
      if (el < nLocal) Xa[(size_t)el * M + kk[j]] = __ldg(Xall + (size_t)(e0 + el) * M + kk[j]) + acc[j][c] * scale;   // ScalarMatMult, MatAdd
    }
  }
}
#endif
