// Batched evaluation-side ops next to the solver:
//   frustum_inside_mask_f32 : the ground-truth / prediction-check label rule
//        0 <= u <= W-1  and  0 <= v <= H-1  and  z > 0.1      with  [u v 1]^T ~ K (P p)
//        (evaluation/registration_lsq.py:67-84, models/multimodal_classifier.py:136-148)
//   pose_error_batch        : get_P_diff (evaluation/registration_lsq.py:87-95): P_diff = P_pred^-1 P_gt,
//        translation error |P_diff[:3,3]|, rotation error = sum |euler 'xzy' angles| in degrees, and the
//        authors' success flag t < 2 m and r < 5 deg (evaluation/registration_result_analysis.py:37-38).
#include <cmath>

#include "common.cuh"

namespace dib {

__global__ void inside_mask_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ n_pts, int n_stride,
                                   const double* __restrict__ P16, const double* __restrict__ K9, double H, double W,
                                   int8_t* __restrict__ mask) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_stride) return;
  const int n = n_pts ? n_pts[s] : n_stride;
  int8_t out = -1;
  if (i < n) {
    const double* P = P16 + (size_t)s * 16;
    const double* K = K9 + (size_t)s * 9;
    const float* b = xyz + (size_t)s * 3 * n_stride;
    const double x = b[i], y = b[n_stride + i], z = b[2 * (size_t)n_stride + i];
    // P_points = (P [x y z 1]^T)[0:3]
    const double X = P[0] * x + P[1] * y + P[2] * z + P[3];
    const double Y = P[4] * x + P[5] * y + P[6] * z + P[7];
    const double Z = P[8] * x + P[9] * y + P[10] * z + P[11];
    // K_pc = K P_points ; pxpy = K_pc[0:2] / K_pc[2]
    const double kx = K[0] * X + K[1] * Y + K[2] * Z;
    const double ky = K[3] * X + K[4] * Y + K[5] * Z;
    const double kz = K[6] * X + K[7] * Y + K[8] * Z;
    const double u = kx / kz, v = ky / kz;
    out = (u >= 0.0 && u <= W - 1.0 && v >= 0.0 && v <= H - 1.0 && Z > 0.1) ? 1 : 0;
  }
  mask[(size_t)s * n_stride + i] = out;
}

__global__ void pose_error_kernel(const double* __restrict__ Pp, const double* __restrict__ Pg, int S,
                                  double t_thresh, double r_thresh, double* __restrict__ t_err,
                                  double* __restrict__ r_err_deg, int32_t* __restrict__ success) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const double* A = Pp + (size_t)s * 16;   // predicted pose (rigid): inverse = [R^T | -R^T t]
  const double* B = Pg + (size_t)s * 16;
  double R[9], t[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += A[4 * k + i] * B[4 * k + j];      // (A_R^T B_R)[i][j]
      R[3 * i + j] = acc;
    }
    double acc = 0.0;
    for (int k = 0; k < 3; ++k) acc += A[4 * k + i] * (B[4 * k + 3] - A[4 * k + 3]);   // A_R^T (t_B - t_A)
    t[i] = acc;
  }
  const double te = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  // extrinsic x-z-y Euler angles of R = Ry(c) Rz(b) Rx(a):  b = asin(R10), a = atan2(-R12, R11), c = atan2(-R20, R00)
  const double r10 = fmin(1.0, fmax(-1.0, R[3]));
  const double kDeg = 57.29577951308232087680;
  const double a = atan2(-R[5], R[4]), b = asin(r10), c = atan2(-R[6], R[0]);
  const double re = (fabs(a) + fabs(b) + fabs(c)) * kDeg;
  t_err[s] = te;
  r_err_deg[s] = re;
  if (success) success[s] = (te < t_thresh && re < r_thresh) ? 1 : 0;
}

}  // namespace dib

extern "C" {

int frustum_inside_mask_f32(const float* xyz, const int32_t* n_pts, int n_stride, const double* P16, const double* K9,
                            double H, double W, int S, int8_t* mask_out, dib_stream_t stream) {
  using namespace dib;
  DIB_REQUIRE(xyz && P16 && K9 && mask_out, "NULL argument");
  DIB_REQUIRE(S >= 0 && n_stride >= 0 && S <= 65535, "bad sizes");
  if (S == 0 || n_stride == 0) return DIB_OK;
  dim3 grid((n_stride + 255) / 256, S);
  inside_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, n_pts, n_stride, P16, K9, H, W, mask_out);
  DIB_CHECK_CUDA(cudaGetLastError());
  return DIB_OK;
}

int pose_error_batch(const double* P_pred16, const double* P_gt16, int S, double t_thresh_m, double r_thresh_deg,
                     double* t_err, double* r_err_deg, int32_t* success, dib_stream_t stream) {
  using namespace dib;
  DIB_REQUIRE(P_pred16 && P_gt16 && t_err && r_err_deg, "NULL argument");
  DIB_REQUIRE(S >= 0, "bad sizes");
  if (S == 0) return DIB_OK;
  pose_error_kernel<<<(S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P_pred16, P_gt16, S, t_thresh_m, r_thresh_deg,
                                                                         t_err, r_err_deg, success);
  DIB_CHECK_CUDA(cudaGetLastError());
  return DIB_OK;
}

}  // extern "C"
