/* =====================================================================================
 * TEST INFRASTRUCTURE ONLY -- CPU oracles for the two point-cloud ops on the hot path.
 * Never imported / linked / executed by the product path (deepi2p_b200/).
 *
 * index_max : restates models/index_max_ext/index_max.cpp:73-112 (forward_cpu), which is
 *             semantically identical to the CUDA kernel the model actually calls
 *             (index_max_cuda.cu:30-62).  Pinned against the reference's own forward_cpu
 *             compiled from /root/reference (oracle/build_ref.py -> tests/golden/).
 * ball_query: restates models/ball_query_ext/ball_query_cuda.cu:11-50 (the reference has no
 *             CPU path).  Pinned only on the GPU box against the reference kernel itself
 *             (oracle/_ref), see tests/test_ops_gpu.py.
 * ===================================================================================== */
#include <stdint.h>
#include <stddef.h>

/* data f32 [B,C,N], index i32 [B,N] (values in [0,K)), out i32 [B,C,K].
 * Sequential scan in ascending n with a strict '>' against a running max that starts at
 * -1000 (index_max.cpp:81,103-106): lowest n wins ties, NaN never wins, untouched -> 0. */
int index_max_oracle(const float* data, const int32_t* index, int32_t* out,
                     int64_t B, int64_t C, int64_t N, int64_t K, float* scratch_best /* B*C*K */) {
  for (int64_t i = 0; i < B * C * K; ++i) { out[i] = 0; scratch_best[i] = -1000.0f; }
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c) {
      const float* row = data + (b * C + c) * N;
      const int32_t* idx = index + b * N;
      float* best = scratch_best + (b * C + c) * K;
      int32_t* o = out + (b * C + c) * K;
      for (int64_t n = 0; n < N; ++n) {
        const int32_t k = idx[n];
        if (k < 0 || k >= K) return -1;   /* UB in the reference; the oracle refuses */
        const float v = row[n];
        if (v > best[k]) { best[k] = v; o[k] = (int32_t)n; }
      }
    }
  return 0;
}

/* dist f32 [B,M,N], out i32 [B,M,K]: first K indices (ascending n) with dist <= radius;
 * none -> zeros; fewer than K -> cyclic repetition of the hits (ball_query_cuda.cu:23-47). */
int ball_query_oracle(const float* dist, float radius, int32_t* out,
                      int64_t B, int64_t M, int64_t N, int64_t K) {
  for (int64_t r = 0; r < B * M; ++r) {
    const float* row = dist + r * N;
    int32_t* o = out + r * K;
    int64_t cnt = 0;
    for (int64_t n = 0; n < N && cnt < K; ++n)
      if (row[n] <= radius) o[cnt++] = (int32_t)n;
    if (cnt == 0) { for (int64_t i = 0; i < K; ++i) o[i] = 0; }
    else if (cnt < K) { for (int64_t i = 0; i < K - cnt; ++i) o[cnt + i] = o[i % cnt]; }
  }
  return 0;
}
