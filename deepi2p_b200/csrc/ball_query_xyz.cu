// Coordinate-based radius search: the same output contract as ball_query
// (first K point indices in ascending n within `radius` of each node; none -> zeros; fewer -> cyclic
// repeat, models/ball_query_ext/ball_query_cuda.cu:11-50) but computed from coordinates through a
// uniform grid hash, so the dense B x M x N distance matrix the reference op needs (268 MB at
// BASELINE config 3) is never built or read.
//
//   points [B][3][N] f32, nodes [B][3][M] f32 (the reference's channel-first layout, networks_pc.py:47-65)
//   hit  <=>  ((dx*dx + dy*dy) + dz*dz) <= radius*radius   evaluated in float32 WITHOUT fma contraction
//             (so a numpy float32 restatement is bit-identical).
//
// Kernel 1 (one CTA per batch item): bounding box, cell = max(radius, extent/16) per axis (<= 4096
// cells), counting sort of point indices by cell into the workspace.
// Kernel 2 (one warp per node): visits the 27 neighbouring cells, sets one bit per hit in a per-warp
// shared-memory bitmap over n, then extracts the first K set bits in order -- cell order never leaks
// into the result.
#include <cfloat>

#include "common.cuh"

namespace dib {

constexpr int kGridMaxDim = 16;
constexpr int kGridMaxCells = kGridMaxDim * kGridMaxDim * kGridMaxDim;
constexpr int kGridThreads = 512;
constexpr int kGridMaxN = 65536;      // bitmap of N bits per warp must fit in shared memory

struct GridHeader {                   // per batch item, at the head of its workspace slice
  float lo[3], inv_cell[3];
  int dim[3];
  int pad;
};

__host__ __device__ inline size_t grid_slice_ints(int N) {
  return sizeof(GridHeader) / 4 + (kGridMaxCells + 1) + (size_t)((N + 3) & ~3);
}

__device__ __forceinline__ int cell_coord(float v, float lo, float inv, int dim) {
  int c = (int)floorf((v - lo) * inv);
  return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}

__global__ void __launch_bounds__(kGridThreads) bq_grid_build_kernel(const float* __restrict__ points, int N,
                                                                     float radius, int32_t* __restrict__ ws) {
  __shared__ int hist[kGridMaxCells + 1];
  __shared__ float red[6][kGridThreads / 32];
  __shared__ GridHeader hdr;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* px = points + (size_t)b * 3 * N;
  int32_t* slice = ws + (size_t)b * grid_slice_ints(N);
  // bounding box
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = tid; i < N; i += kGridThreads)
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float v = px[(size_t)c * N + i]; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
    if (lane == 0) { red[c][warp] = lo[c]; red[3 + c][warp] = hi[c]; }
  }
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < 3; ++c) {
      float l = red[c][0], h = red[3 + c][0];
      for (int w = 1; w < kGridThreads / 32; ++w) { l = fminf(l, red[c][w]); h = fmaxf(h, red[3 + c][w]); }
      const float extent = fmaxf(h - l, 0.0f);
      int dim = (radius > 0.0f && extent > 0.0f) ? (int)floorf(extent / radius) : 1;
      dim = dim < 1 ? 1 : (dim > kGridMaxDim ? kGridMaxDim : dim);
      // cell size extent/dim >= radius; inflate a hair so that rounding can only make cells larger
      const float cell = (extent > 0.0f) ? (extent / (float)dim) * 1.0001f : 1.0f;
      hdr.lo[c] = l; hdr.inv_cell[c] = 1.0f / cell; hdr.dim[c] = dim;
    }
    hdr.pad = 0;
  }
  for (int i = tid; i <= kGridMaxCells; i += kGridThreads) hist[i] = 0;
  __syncthreads();
  const int ncell = hdr.dim[0] * hdr.dim[1] * hdr.dim[2];
  // histogram
  for (int i = tid; i < N; i += kGridThreads) {
    const int cx = cell_coord(px[i], hdr.lo[0], hdr.inv_cell[0], hdr.dim[0]);
    const int cy = cell_coord(px[(size_t)N + i], hdr.lo[1], hdr.inv_cell[1], hdr.dim[1]);
    const int cz = cell_coord(px[2 * (size_t)N + i], hdr.lo[2], hdr.inv_cell[2], hdr.dim[2]);
    atomicAdd(&hist[(cx * hdr.dim[1] + cy) * hdr.dim[2] + cz], 1);
  }
  __syncthreads();
  // exclusive scan of at most 4096 counts by one warp (128 per lane)
  if (warp == 0) {
    const int per = (ncell + 31) / 32;
    int sum = 0;
    for (int k = 0; k < per; ++k) { const int j = lane * per + k; if (j < ncell) sum += hist[j]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    int run = incl - sum;
    for (int k = 0; k < per; ++k) { const int j = lane * per + k; if (j < ncell) { const int c = hist[j]; hist[j] = run; run += c; } }
    if (lane == 31) hist[ncell] = incl;
  }
  __syncthreads();
  int32_t* cell_start = slice + sizeof(GridHeader) / 4;
  int32_t* order = cell_start + (kGridMaxCells + 1);
  for (int i = tid; i <= ncell; i += kGridThreads) cell_start[i] = hist[i];
  if (tid == 0) *reinterpret_cast<GridHeader*>(slice) = hdr;
  __syncthreads();
  // scatter (cursor = hist, order inside a cell is irrelevant: the query restores ascending n)
  for (int i = tid; i < N; i += kGridThreads) {
    const int cx = cell_coord(px[i], hdr.lo[0], hdr.inv_cell[0], hdr.dim[0]);
    const int cy = cell_coord(px[(size_t)N + i], hdr.lo[1], hdr.inv_cell[1], hdr.dim[1]);
    const int cz = cell_coord(px[2 * (size_t)N + i], hdr.lo[2], hdr.inv_cell[2], hdr.dim[2]);
    const int pos = atomicAdd(&hist[(cx * hdr.dim[1] + cy) * hdr.dim[2] + cz], 1);
    order[pos] = i;
  }
}

constexpr int kQueryWarps = 4;

__global__ void __launch_bounds__(kQueryWarps * 32) bq_grid_query_kernel(const float* __restrict__ points,
                                                                         const float* __restrict__ nodes, float radius,
                                                                         const int32_t* __restrict__ ws,
                                                                         int32_t* __restrict__ out, int B, int M, int N,
                                                                         int K) {
  extern __shared__ unsigned bitmap_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long q = (long long)blockIdx.x * kQueryWarps + warp;
  if (q >= (long long)B * M) return;
  const int b = (int)(q / M), m = (int)(q % M);
  const int words = (N + 31) >> 5;
  unsigned* bitmap = bitmap_all + (size_t)warp * words;
  for (int i = lane; i < words; i += 32) bitmap[i] = 0u;
  __syncwarp();
  const int32_t* slice = ws + (size_t)b * grid_slice_ints(N);
  const GridHeader hdr = *reinterpret_cast<const GridHeader*>(slice);
  const int32_t* cell_start = slice + sizeof(GridHeader) / 4;
  const int32_t* order = cell_start + (kGridMaxCells + 1);
  const float* px = points + (size_t)b * 3 * N;
  const float nx = nodes[((size_t)b * 3 + 0) * M + m], ny = nodes[((size_t)b * 3 + 1) * M + m],
              nz = nodes[((size_t)b * 3 + 2) * M + m];
  const float r2 = __fmul_rn(radius, radius);
  const int cx = cell_coord(nx, hdr.lo[0], hdr.inv_cell[0], hdr.dim[0]);
  const int cy = cell_coord(ny, hdr.lo[1], hdr.inv_cell[1], hdr.dim[1]);
  const int cz = cell_coord(nz, hdr.lo[2], hdr.inv_cell[2], hdr.dim[2]);
  if (nx == nx && ny == ny && nz == nz && radius >= 0.0f) {
    for (int ix = max(cx - 1, 0); ix <= min(cx + 1, hdr.dim[0] - 1); ++ix)
      for (int iy = max(cy - 1, 0); iy <= min(cy + 1, hdr.dim[1] - 1); ++iy) {
        // the z-neighbours of a (x,y) column are contiguous in the cell order: one range per column
        const int c0 = (ix * hdr.dim[1] + iy) * hdr.dim[2] + max(cz - 1, 0);
        const int c1 = (ix * hdr.dim[1] + iy) * hdr.dim[2] + min(cz + 1, hdr.dim[2] - 1);
        const int j0 = cell_start[c0], j1 = cell_start[c1 + 1];
        for (int j = j0 + lane; j < j1; j += 32) {
          const int n = order[j];
          const float dx = __fsub_rn(px[n], nx), dy = __fsub_rn(px[(size_t)N + n], ny),
                      dz = __fsub_rn(px[2 * (size_t)N + n], nz);
          const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          if (d2 <= r2) atomicOr(&bitmap[n >> 5], 1u << (n & 31));
        }
      }
  }
  __syncwarp();
  // first K set bits in ascending n
  int32_t* o = out + (size_t)q * K;
  int cnt = 0;
  for (int base = 0; base < words && cnt < K; base += 32) {
    const unsigned w = (base + lane < words) ? bitmap[base + lane] : 0u;
    const int c = __popc(w);
    int incl = c;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, s); if (lane >= s) incl += t; }
    int pos = cnt + incl - c;
    unsigned bits = w;
    while (bits && pos < K) {
      const int bit = __ffs(bits) - 1;
      bits &= bits - 1;
      o[pos++] = ((base + lane) << 5) + bit;
    }
    cnt += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  if (cnt == 0) {
    for (int i = lane; i < K; i += 32) o[i] = 0;
  } else if (cnt < K) {
    for (int i = lane; i < K - cnt; i += 32) o[cnt + i] = o[i % cnt];
  }
}

}  // namespace dib

extern "C" {

size_t ball_query_xyz_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 16;
  return (size_t)B * dib::grid_slice_ints(N) * sizeof(int32_t) + 16;
}

int ball_query_xyz_forward(const float* points, const float* nodes, float radius, int32_t* out, int B, int M, int N,
                           int K, void* workspace, size_t workspace_bytes, dib_stream_t stream) {
  using namespace dib;
  DIB_REQUIRE(points && nodes && out, "NULL argument");
  DIB_REQUIRE(B >= 0 && M >= 0 && N >= 1 && K >= 1, "bad shape B=%d M=%d N=%d K=%d", B, M, N, K);
  DIB_REQUIRE(N <= kGridMaxN, "N=%d exceeds the grid search limit %d (use ball_query_forward on a distance matrix)", N,
              kGridMaxN);
  if (B == 0 || M == 0) return DIB_OK;
  if (workspace == nullptr || workspace_bytes < ball_query_xyz_workspace_bytes(B, N)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, ball_query_xyz_workspace_bytes(B, N));
    return DIB_ENOMEM;
  }
  DIB_REQUIRE(((uintptr_t)workspace % 16) == 0, "workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  bq_grid_build_kernel<<<B, kGridThreads, 0, st>>>(points, N, radius, (int32_t*)workspace);
  DIB_CHECK_CUDA(cudaGetLastError());
  const size_t smem = (size_t)kQueryWarps * ((N + 31) / 32) * sizeof(unsigned);
  DIB_CHECK_CUDA(cudaFuncSetAttribute(bq_grid_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = (long long)B * M;
  bq_grid_query_kernel<<<(unsigned)((total + kQueryWarps - 1) / kQueryWarps), kQueryWarps * 32, smem, st>>>(
      points, nodes, radius, (const int32_t*)workspace, out, B, M, N, K);
  DIB_CHECK_CUDA(cudaGetLastError());
  return DIB_OK;
}

}  // extern "C"
