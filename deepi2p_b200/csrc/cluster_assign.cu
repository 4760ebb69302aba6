// Point -> node clustering front-end of the point-cloud encoder (first half):
// models/networks_pc.py:60-85 without the B x N x Ma intermediates.  The reference materialises
// pc_B3NMa - node_a_B3NMa, its norm, a B x N x Ma mask (byte and float) and a B x 3 x N x Ma masked copy
// of the cloud (8 x 3 x 20480 x 128 floats = 252 MB each at the shipped configuration) to obtain
//
//   min_k_idx [B,N,k]  the k nearest nodes of every point, nearest first        (:63-64, torch.topk)
//   min_idx   [B,N]    = min_k_idx[..., 0], the `index` argument of index_max    (:65, :88-90)
//   count     [B,Ma]   points per node; mask_row_max = count > 0                 (:66-72)
//   cluster_mean [B,3,Ma] = sum of the node's points / (count + 1e-5)            (:74-76)
//   pc_centers [B,3,N] = cluster_mean gathered by min_idx;  pc_decentered = pc - pc_centers  (:78-82)
//
// Here: kernel 1 keeps the nodes of one batch item in shared memory, one thread per point scans them,
// keeps the k best in registers and adds the point to its node's count and coordinate sums; kernel 2
// turns sums into means; kernel 3 gathers and subtracts.  HBM traffic = the cloud twice + the outputs.
//
// Arithmetic contract (oracle: oracle.cluster_assign):
//   * ordering key d2 = ((dx*dx + dy*dy) + dz*dz), float32, no fma, d = point - node; ties -> lower
//     node index (torch.topk leaves tie order unspecified; sqrt is monotone so the order on d2 is a
//     valid order on the reference's norm).
//   * coordinate sums are exact fixed-point: sum of rint(x * 2^24) in int64 -- associative, so the
//     result does not depend on the order the atomics land in (the reference's float tree-sum has
//     no defined order); mean = float(sum * 2^-24) / (float(count) + 1e-5f).
#include <cfloat>

#include "common.cuh"

namespace dib {

constexpr int kCaThreads = 256;
constexpr int kCaPointsPerThread = 1;
constexpr int kCaMaxK = 8;
constexpr int kCaMaxNodes = 2048;
constexpr double kCaFixedScale = 16777216.0;        // 2^24

template <int K>
__global__ void __launch_bounds__(kCaThreads)
    cluster_assign_kernel(const float* __restrict__ pc, const float* __restrict__ node, int N, int M,
                          int32_t* __restrict__ topk, int32_t* __restrict__ min_idx, int32_t* __restrict__ count,
                          unsigned long long* __restrict__ sums) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s_sum = reinterpret_cast<unsigned long long*>(smem_raw);     // [3][M]
  float* s_node = reinterpret_cast<float*>(s_sum + 3 * (size_t)M);                 // [3][M]
  int* s_cnt = reinterpret_cast<int*>(s_node + 3 * (size_t)M);                     // [M]
  const int b = blockIdx.y, tid = threadIdx.x;
  const float* nb = node + (size_t)b * 3 * M;
  for (int i = tid; i < 3 * M; i += kCaThreads) { s_node[i] = nb[i]; s_sum[i] = 0ull; }
  for (int i = tid; i < M; i += kCaThreads) s_cnt[i] = 0;
  __syncthreads();

  const float* px = pc + (size_t)b * 3 * N;
  const int base = blockIdx.x * (kCaThreads * kCaPointsPerThread);
#pragma unroll 1
  for (int q = 0; q < kCaPointsPerThread; ++q) {
    const int n = base + q * kCaThreads + tid;
    if (n >= N) break;
    const float x = px[n], y = px[(size_t)N + n], z = px[2 * (size_t)N + n];
    float bd[K];
    int bi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { bd[j] = __int_as_float(0x7f800000); bi[j] = j; }   // +inf: never beaten by inf / NaN
#pragma unroll 4
    for (int m = 0; m < M; ++m) {
      const float dx = __fsub_rn(x, s_node[m]), dy = __fsub_rn(y, s_node[M + m]), dz = __fsub_rn(z, s_node[2 * M + m]);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (!(d2 < bd[K - 1])) continue;                  // strict: on a tie the earlier (lower) node stays
      // sorted insert from the back; all indices are compile-time, so bd / bi stay in registers
#pragma unroll
      for (int j = K - 1; j >= 1; --j) {
        const bool shift = d2 < bd[j - 1];
        const bool here = !shift && d2 < bd[j];
        bd[j] = shift ? bd[j - 1] : (here ? d2 : bd[j]);
        bi[j] = shift ? bi[j - 1] : (here ? m : bi[j]);
      }
      if (d2 < bd[0]) { bd[0] = d2; bi[0] = m; }
    }
    int32_t* o = topk + ((size_t)b * N + n) * K;
#pragma unroll
    for (int j = 0; j < K; ++j) o[j] = bi[j];
    const int m0 = bi[0];
    min_idx[(size_t)b * N + n] = m0;
    // a point with a non-finite coordinate has no fixed-point image: it keeps min_idx = 0 (every distance compares
    // false) but is left out of its node's count and sums (the reference's float sum would turn the mean into NaN)
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      atomicAdd(&s_cnt[m0], 1);
      atomicAdd(&s_sum[m0], (unsigned long long)__double2ll_rn((double)x * kCaFixedScale));
      atomicAdd(&s_sum[M + m0], (unsigned long long)__double2ll_rn((double)y * kCaFixedScale));
      atomicAdd(&s_sum[2 * M + m0], (unsigned long long)__double2ll_rn((double)z * kCaFixedScale));
    }
  }
  __syncthreads();
  for (int i = tid; i < M; i += kCaThreads) {
    const int c = s_cnt[i];
    if (c) {
      atomicAdd(&count[(size_t)b * M + i], c);
#pragma unroll
      for (int a = 0; a < 3; ++a) atomicAdd(&sums[((size_t)b * 3 + a) * M + i], s_sum[a * M + i]);
    }
  }
}

// one thread per (b, axis, node)
__global__ void cluster_mean_kernel(const unsigned long long* __restrict__ sums, const int32_t* __restrict__ count,
                                    int M, int total, float* __restrict__ mean) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int m = i % M, b = i / (3 * M);
  const float num = (float)((double)(long long)sums[i] * (1.0 / kCaFixedScale));
  mean[i] = __fdiv_rn(num, __fadd_rn((float)count[(size_t)b * M + m], 1e-5f));
}

// one thread per (b, n)
__global__ void cluster_decenter_kernel(const float* __restrict__ pc, const int32_t* __restrict__ min_idx,
                                        const float* __restrict__ mean, int N, int M, float* __restrict__ centers,
                                        float* __restrict__ decentered) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (n >= N) return;
  const int m = min_idx[(size_t)b * N + n];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const size_t i = ((size_t)b * 3 + a) * N + n;
    const float c = mean[((size_t)b * 3 + a) * M + m];
    if (centers) centers[i] = c;
    if (decentered) decentered[i] = __fsub_rn(pc[i], c);
  }
}

}  // namespace dib

extern "C" {

size_t cluster_assign_workspace_bytes(int B, int M) {
  if (B <= 0 || M <= 0) return 0;
  return (size_t)B * 3 * (size_t)M * sizeof(unsigned long long);
}

// pc [B][3][N] f32, node [B][3][M] f32 [dev]; outputs [dev]: min_k_idx [B][N][k] i32, min_idx [B][N] i32,
// count [B][M] i32, cluster_mean [B][3][M] f32, pc_centers / pc_decentered [B][3][N] f32 (either may be
// NULL).  1 <= k <= min(8, M), M <= 2048.
int cluster_assign_forward(const float* pc, const float* node, int B, int N, int M, int k, int32_t* min_k_idx,
                           int32_t* min_idx, int32_t* count, float* cluster_mean, float* pc_centers,
                           float* pc_decentered, void* workspace, size_t workspace_bytes, dib_stream_t stream_) {
  using namespace dib;
  cudaStream_t stream = (cudaStream_t)stream_;
  DIB_REQUIRE(B >= 0 && N >= 0 && M >= 1 && M <= kCaMaxNodes, "cluster_assign: need 1 <= M <= 2048");
  DIB_REQUIRE(k >= 1 && k <= kCaMaxK && k <= M, "cluster_assign: need 1 <= k <= min(8, M)");
  if (B == 0) return DIB_OK;
  DIB_REQUIRE(node && count && cluster_mean && (N == 0 || (pc && min_k_idx && min_idx)), "cluster_assign: NULL argument");
  DIB_REQUIRE(B <= 65535, "cluster_assign: B <= 65535");
  const size_t need = cluster_assign_workspace_bytes(B, M);
  DIB_REQUIRE(workspace && workspace_bytes >= need && ((uintptr_t)workspace & 7) == 0,
              "cluster_assign: workspace too small or misaligned");
  unsigned long long* sums = (unsigned long long*)workspace;
  DIB_CHECK_CUDA(cudaMemsetAsync(sums, 0, need, stream));
  DIB_CHECK_CUDA(cudaMemsetAsync(count, 0, (size_t)B * M * sizeof(int32_t), stream));
  if (N > 0) {
    const size_t smem = (size_t)M * (3 * sizeof(unsigned long long) + 3 * sizeof(float) + sizeof(int));
    const int per_cta = kCaThreads * kCaPointsPerThread;
    dim3 grid((N + per_cta - 1) / per_cta, B);
#define DIB_CA_LAUNCH(KK)                                                                                         \
  case KK:                                                                                                        \
    if (smem > 48 * 1024)                                                                                         \
      DIB_CHECK_CUDA(cudaFuncSetAttribute(cluster_assign_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)((size_t)kCaMaxNodes * 40)));                                      \
    cluster_assign_kernel<KK><<<grid, kCaThreads, smem, stream>>>(pc, node, N, M, min_k_idx, min_idx, count, sums); \
    break;
    switch (k) {
      DIB_CA_LAUNCH(1) DIB_CA_LAUNCH(2) DIB_CA_LAUNCH(3) DIB_CA_LAUNCH(4)
      DIB_CA_LAUNCH(5) DIB_CA_LAUNCH(6) DIB_CA_LAUNCH(7) DIB_CA_LAUNCH(8)
    }
#undef DIB_CA_LAUNCH
    DIB_CHECK_CUDA(cudaGetLastError());
  }
  const int total = B * 3 * M;
  cluster_mean_kernel<<<(total + 255) / 256, 256, 0, stream>>>(sums, count, M, total, cluster_mean);
  DIB_CHECK_CUDA(cudaGetLastError());
  if (N > 0 && (pc_centers || pc_decentered)) {
    dim3 grid((N + 255) / 256, B);
    cluster_decenter_kernel<<<grid, 256, 0, stream>>>(pc, min_idx, cluster_mean, N, M, pc_centers, pc_decentered);
    DIB_CHECK_CUDA(cudaGetLastError());
  }
  return DIB_OK;
}

}  // extern "C"
