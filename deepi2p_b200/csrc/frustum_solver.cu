// Batched robust bounded Levenberg-Marquardt solver for the inverse-camera-projection
// ("frustum") registration problem, written for sm_100a.
//
// Replaces FrustumRegistration.solvePGivenK (evaluation/frustum_reg/src/registration.cpp:9-186)
// and the multi-start loop around it (evaluation/registration_lsq.py:127-186).  One thread block
// owns one (cloud, labels, intrinsics, init pose) problem from its first evaluation to its final
// pose: per pass the block culls whole 32-point groups against the frustum with a per-launch box table,
// classifies the points of undecided groups in fp32 with a conservative margin, evaluates the
// maybe-active ones exactly in fp64 (residual, analytic Jacobian, Cauchy corrector), reduces
// cost, J^T r and J^T J in a fixed order, and thread 0 runs the trust-region control flow
// (Jacobi scaling, LM damping, Cholesky of the damped normal matrix, model cost change, box
// projection, projected Armijo line search with cubic interpolation, step acceptance and the
// tolerance tests) without ever returning to the host.  A persistent grid pulls problems from an
// atomic queue; a second tiny kernel takes the per-sample arg-min over inits and builds the 4x4.
//
// Residual definitions: registration_3d.hpp:34-68,105-127 / registration_2d.hpp:34-69,106-129.
// Algorithm text: SURVEY.md Appendix A; oracle/frustum_oracle.cpp is the CPU checker.
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdlib>

#include "common.cuh"

namespace dib {

#ifndef DIB_WIDE_TU
#define DIB_WIDE_TU 0                         // 1: this file is being compiled a second time by frustum_solver_wide.cu (128-thread
#endif                                        //    CTAs, own namespace); the C ABI and the error buffer live in the primary TU only
#if !DIB_WIDE_TU
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
#endif

#ifndef DIB_THREADS
#define DIB_THREADS 64
#endif
#ifndef DIB_MINBLOCKS4
#define DIB_MINBLOCKS4 10                         // ptxas settles on 96 registers (~60 B of spills) for both 9 and 10;
                                              // 8 (124 registers, no spills, 16 warps/SM) measured slower
#endif
#ifndef DIB_MINBLOCKS6
#define DIB_MINBLOCKS6 8                          // 128 registers (6 -> 168 registers measured 2 % slower)
#endif
constexpr int kThreads = DIB_THREADS;   // threads per problem (CTA)
constexpr int kWarps = kThreads / 32;


struct Cam {
  double fx, fy, cx, cy, W1, H1, hW, hH;
};

// Per-evaluation pose constants, computed once by thread 0 and broadcast through shared memory.
struct PoseConst {
  double R[9];    // rotation applied to the point (first-order I + [a]x when |a|^2 <= DBL_EPSILON)
  double t[3];
  double Jl[9];   // 6-DoF: left Jacobian of SO(3); d(R p)/da = -[R p]x Jl
  double cD, sD;  // 4-DoF: cos/sin used in d q / d ry
  int small;      // first-order branch taken
};

template <int P>
struct NAcc {
  static constexpr int NA = P * (P + 1) / 2;
  static constexpr int N = 1 + P + NA;   // cost, g, packed upper triangle of J^T J
};

__host__ __device__ constexpr int tri(int P, int j, int k) {   // j <= k
  return j * P - j * (j - 1) / 2 + (k - j);
}

// ------------------------------------------------------------------------------------------
// Pose constants.
// ------------------------------------------------------------------------------------------
template <int P>
__device__ __noinline__ void make_pose(const double* x, PoseConst* pc) {
  if (P == 4) {
    const double ry = x[0];
    pc->t[0] = x[1]; pc->t[1] = x[2]; pc->t[2] = x[3];
    double c, s;
    if (ry * ry > DBL_EPSILON) { sincos(ry, &s, &c); pc->cD = c; pc->sD = s; pc->small = 0; }
    else { c = 1.0; s = ry; pc->cD = 1.0; pc->sD = 0.0; pc->small = 1; }
    pc->R[0] = c; pc->R[1] = 0; pc->R[2] = s;
    pc->R[3] = 0; pc->R[4] = 1; pc->R[5] = 0;
    pc->R[6] = -s; pc->R[7] = 0; pc->R[8] = c;
  } else {
    const double ax = x[0], ay = x[1], az = x[2];
    pc->t[0] = x[3]; pc->t[1] = x[4]; pc->t[2] = x[5];
    const double th2 = ax * ax + ay * ay + az * az;
    if (th2 > DBL_EPSILON) {
      const double th = sqrt(th2);
      double s, c;
      sincos(th, &s, &c);
      const double wx = ax / th, wy = ay / th, wz = az / th;
      const double omc = 1.0 - c;
      pc->R[0] = c + wx * wx * omc;      pc->R[1] = wx * wy * omc - wz * s; pc->R[2] = wy * s + wx * wz * omc;
      pc->R[3] = wz * s + wx * wy * omc; pc->R[4] = c + wy * wy * omc;      pc->R[5] = -wx * s + wy * wz * omc;
      pc->R[6] = -wy * s + wx * wz * omc; pc->R[7] = wx * s + wy * wz * omc; pc->R[8] = c + wz * wz * omc;
      // Jl = I + A [a]x + B [a]x^2,  A = (1-cos)/th^2,  B = (th - sin)/th^3   (cancellation-free forms)
      double sh, ch;
      sincos(0.5 * th, &sh, &ch);
      const double q = sh / (0.5 * th);
      const double A = 0.5 * q * q;
      double B;
      if (th < 0.5) {
        const double t2 = th2;
        B = 1.0 / 6.0 + t2 * (-1.0 / 120.0 + t2 * (1.0 / 5040.0 + t2 * (-1.0 / 362880.0 +
            t2 * (1.0 / 39916800.0 + t2 * (-1.0 / 6227020800.0 + t2 * (1.0 / 1307674368000.0))))));
      } else {
        B = (th - s) / (th2 * th);
      }
      // [a]x^2 = a a^T - th2 I
      pc->Jl[0] = 1.0 + B * (ax * ax - th2); pc->Jl[1] = -A * az + B * ax * ay;     pc->Jl[2] = A * ay + B * ax * az;
      pc->Jl[3] = A * az + B * ax * ay;      pc->Jl[4] = 1.0 + B * (ay * ay - th2); pc->Jl[5] = -A * ax + B * ay * az;
      pc->Jl[6] = -A * ay + B * ax * az;     pc->Jl[7] = A * ax + B * ay * az;      pc->Jl[8] = 1.0 + B * (az * az - th2);
      pc->small = 0;
    } else {
      pc->R[0] = 1;   pc->R[1] = -az; pc->R[2] = ay;
      pc->R[3] = az;  pc->R[4] = 1;   pc->R[5] = -ax;
      pc->R[6] = -ay; pc->R[7] = ax;  pc->R[8] = 1;
      for (int i = 0; i < 9; ++i) pc->Jl[i] = (i % 4 == 0) ? 1.0 : 0.0;
      pc->small = 1;
    }
    pc->cD = 0; pc->sD = 0;
  }
}

// ------------------------------------------------------------------------------------------
// Per-point evaluation.  acc = [cost, g[P], packed upper J^T J]; everything weighted by the
// Cauchy corrector w = rho'(s) = 1/(1+s) (CauchyLoss(1.0), registration.cpp:103,121).
// ------------------------------------------------------------------------------------------
// Accumulator view over shared memory: acc[j] is column `lane` of row j ([N][32] doubles per warp).
struct SmemAcc {
  double* base;
  __device__ __forceinline__ double& operator[](int j) const { return base[j * 32]; }
};

template <int P, typename ACC>
__device__ __forceinline__ void rank1(ACC acc, const double* J, double w, double r) {
  const double wr = w * r;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    acc[1 + j] = fma(wr, J[j], acc[1 + j]);
    const double wj = w * J[j];
#pragma unroll
    for (int k = j; k < P; ++k) acc[1 + P + tri(P, j, k)] = fma(wj, J[k], acc[1 + P + tri(P, j, k)]);
  }
}

// Rows of the block Jacobian before robustification.  Returns the number of rows (1 or 3) and
// fills r[] and J[][P]; inactive rows are exactly zero.
template <int P>
__device__ __forceinline__ bool point_rows(double px, double py, double pz, int lab, const Cam& cam,
                                           const PoseConst& pc, double r[3], double J[3][P], int* nrows) {
  constexpr int NR = P - 3;
  const double rx = fma(pc.R[0], px, fma(pc.R[1], py, pc.R[2] * pz));
  const double ry_ = fma(pc.R[3], px, fma(pc.R[4], py, pc.R[5] * pz));
  const double rz = fma(pc.R[6], px, fma(pc.R[7], py, pc.R[8] * pz));
  const double X = rx + pc.t[0], Y = ry_ + pc.t[1], Z = rz + pc.t[2];
  const double iz = 1.0 / Z;
  const double u = fma(cam.fx * X, iz, cam.cx);
  const double v = fma(cam.fy * Y, iz, cam.cy);

  bool any = false;
  double su = 0.0, sv = 0.0, sz = 0.0;   // d r / d u, d r / d v, d r / d Z selectors per row
  if (lab == 0) {
    *nrows = 1;
    const double du_ = u - cam.hW, dv_ = v - cam.hH;
    const double xd = cam.hW - fabs(du_);
    const double yd = cam.hH - fabs(dv_);
    r[0] = 0.0;
    if (Z > 0.0 && xd > 0.0 && yd > 0.0) {
      r[0] = xd + yd;
      su = (du_ < 0.0) ? 1.0 : -1.0;     // -sgn(u - W1/2), sgn(0) = +1
      sv = (dv_ < 0.0) ? 1.0 : -1.0;
      any = true;
    }
  } else {
    *nrows = 3;
    const double a0 = -u, b0 = u - cam.W1;
    const double a1 = -v, b1 = v - cam.H1;
    r[0] = (a0 < 0.0 ? 0.0 : a0) + (b0 < 0.0 ? 0.0 : b0);
    r[1] = (a1 < 0.0 ? 0.0 : a1) + (b1 < 0.0 ? 0.0 : b1);
    r[2] = (-Z < 0.0 ? 0.0 : -Z) * 100.0;
    su = (a0 < 0.0 ? 0.0 : -1.0) + (b0 < 0.0 ? 0.0 : 1.0);
    sv = (a1 < 0.0 ? 0.0 : -1.0) + (b1 < 0.0 ? 0.0 : 1.0);
    sz = (-Z < 0.0) ? 0.0 : -100.0;
    any = (su != 0.0) || (sv != 0.0) || (sz != 0.0);
  }
  if (!any) return false;

  // d q / d rot (NR columns).
  double dX[NR], dY[NR], dZ[NR];
  if (P == 4) {
    dX[0] = fma(-pc.sD, px, pc.cD * pz);
    dY[0] = 0.0;
    dZ[0] = -fma(pc.cD, px, pc.sD * pz);
  } else {
    const double a = pc.small ? px : rx, b = pc.small ? py : ry_, c = pc.small ? pz : rz;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      dX[k] = c * pc.Jl[3 + k] - b * pc.Jl[6 + k];
      dY[k] = a * pc.Jl[6 + k] - c * pc.Jl[k];
      dZ[k] = b * pc.Jl[k] - a * pc.Jl[3 + k];
    }
  }
  const double au = cam.fx * iz, bu = (u - cam.cx) * iz;   // du = au dX - bu dZ
  const double av = cam.fy * iz, bv = (v - cam.cy) * iz;   // dv = av dY - bv dZ
  double dU[P], dV[P];
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    dU[k] = au * dX[k] - bu * dZ[k];
    dV[k] = av * dY[k] - bv * dZ[k];
  }
  dU[NR] = au; dU[NR + 1] = 0.0; dU[NR + 2] = -bu;
  dV[NR] = 0.0; dV[NR + 1] = av; dV[NR + 2] = -bv;
  if (lab == 0) {
#pragma unroll
    for (int j = 0; j < P; ++j) J[0][j] = su * dU[j] + sv * dV[j];
  } else {
#pragma unroll
    for (int j = 0; j < P; ++j) { J[0][j] = su * dU[j]; J[1][j] = sv * dV[j]; }
#pragma unroll
    for (int k = 0; k < NR; ++k) J[2][k] = sz * dZ[k];
    J[2][NR] = 0.0; J[2][NR + 1] = 0.0; J[2][NR + 2] = sz;
  }
  return true;
}

template <int P>
__device__ __forceinline__ void point_accumulate(double px, double py, double pz, int lab, const Cam& cam,
                                                 const PoseConst& pc, double* acc) {
  double r[3], J[3][P];
  int nrows;
  if (!point_rows<P>(px, py, pz, lab, cam, pc, r, J, &nrows)) return;
  if (lab == 0) {
    const double s = r[0] * r[0];
    const double sum = 1.0 + s;
    const double w = 1.0 / sum;
    acc[0] += 0.5 * log(sum);
    rank1<P, double*>(acc, J[0], w, r[0]);
  } else {
    const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    const double sum = 1.0 + s;
    const double w = 1.0 / sum;
    acc[0] += 0.5 * log(sum);
    // rows with a zero selector are exactly zero and contribute nothing
    if (r[0] != 0.0 || J[0][P - 3] != 0.0) rank1<P, double*>(acc, J[0], w, r[0]);
    if (r[1] != 0.0 || J[1][P - 2] != 0.0) rank1<P, double*>(acc, J[1], w, r[1]);
    if (J[2][P - 1] != 0.0) rank1<P, double*>(acc, J[2], w, r[2]);
  }
}

// ------------------------------------------------------------------------------------------
// Branch-free evaluators used by the solver's hot loop.  Batches are homogeneous in the label and
// every lane evaluates two points, so that two independent dependency chains are in flight per
// lane.  `valid == false` (padding lane) and exact-inactive points contribute exactly zero.
// ------------------------------------------------------------------------------------------
template <int P>
struct Proj {
  double X, Y, Z, iz, u, v;
  double dX[P - 3], dY[P - 3], dZ[P - 3];
};

template <int P>
__device__ __forceinline__ void project_point(double px, double py, double pz, bool valid, const Cam& cam,
                                              const PoseConst& pc, Proj<P>& o) {
  constexpr int NR = P - 3;
  if (!valid) { px = 0.0; py = 0.0; pz = 0.0; }
  const double rx = fma(pc.R[0], px, fma(pc.R[1], py, pc.R[2] * pz));
  const double ry_ = fma(pc.R[3], px, fma(pc.R[4], py, pc.R[5] * pz));
  const double rz = fma(pc.R[6], px, fma(pc.R[7], py, pc.R[8] * pz));
  o.X = rx + pc.t[0]; o.Y = ry_ + pc.t[1]; o.Z = rz + pc.t[2];
  const double zs = (valid && o.Z != 0.0) ? o.Z : 1.0;     // keeps masked lanes finite
  o.iz = 1.0 / zs;
  o.u = fma(cam.fx * o.X, o.iz, cam.cx);
  o.v = fma(cam.fy * o.Y, o.iz, cam.cy);
  if (P == 4) {
    o.dX[0] = fma(-pc.sD, px, pc.cD * pz);
    o.dY[0] = 0.0;
    o.dZ[0] = -fma(pc.cD, px, pc.sD * pz);
  } else {
    const double a = pc.small ? px : rx, b = pc.small ? py : ry_, c = pc.small ? pz : rz;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      o.dX[k] = c * pc.Jl[3 + k] - b * pc.Jl[6 + k];
      o.dY[k] = a * pc.Jl[6 + k] - c * pc.Jl[k];
      o.dZ[k] = b * pc.Jl[k] - a * pc.Jl[3 + k];
    }
  }
}

// label 0 ("should be outside", registration_3d.hpp:34-68): one dense row.
template <int P>
struct Out0 {
  double w, r, s1;      // corrector weight (0 if inactive), residual, 1 + s
  double J[P];
};

template <int P>
__device__ __forceinline__ void eval_outside(double px, double py, double pz, bool valid, const Cam& cam,
                                             const PoseConst& pc, Out0<P>& o) {
  constexpr int NR = P - 3;
  Proj<P> q;
  project_point<P>(px, py, pz, valid, cam, pc, q);
  const double du_ = q.u - cam.hW, dv_ = q.v - cam.hH;
  const double xd = cam.hW - fabs(du_), yd = cam.hH - fabs(dv_);
  const bool act = valid && q.Z > 0.0 && xd > 0.0 && yd > 0.0;
  const double su = (du_ < 0.0) ? 1.0 : -1.0;      // -sgn(u - W1/2), sgn(0) = +1
  const double sv = (dv_ < 0.0) ? 1.0 : -1.0;
  o.r = act ? xd + yd : 0.0;
  o.s1 = fma(o.r, o.r, 1.0);
  o.w = act ? 1.0 / o.s1 : 0.0;
  const double au = su * (cam.fx * q.iz), bu = su * ((q.u - cam.cx) * q.iz);
  const double av = sv * (cam.fy * q.iz), bv = sv * ((q.v - cam.cy) * q.iz);
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    double j = au * q.dX[k] - (bu + bv) * q.dZ[k];
    if (P != 4) j = fma(av, q.dY[k], j);
    o.J[k] = j;
  }
  o.J[NR] = au; o.J[NR + 1] = av; o.J[NR + 2] = -(bu + bv);
}

// label 1 ("should be inside", registration_3d.hpp:105-127): three sparse rows
//   JU = su dU (no ty entry), JV = sv dV (no tx entry), JZ = sz dZ (rot and tz entries only).
template <int P>
struct Out1 {
  double w, r0, r1, r2, s1;
  double JU[P], JV[P], JZ[P];
};

template <int P>
__device__ __forceinline__ void eval_inside(double px, double py, double pz, bool valid, const Cam& cam,
                                            const PoseConst& pc, Out1<P>& o) {
  constexpr int NR = P - 3;
  Proj<P> q;
  project_point<P>(px, py, pz, valid, cam, pc, q);
  const double a0 = -q.u, b0 = q.u - cam.W1;
  const double a1 = -q.v, b1 = q.v - cam.H1;
  double su = (a0 < 0.0 ? 0.0 : -1.0) + (b0 < 0.0 ? 0.0 : 1.0);
  double sv = (a1 < 0.0 ? 0.0 : -1.0) + (b1 < 0.0 ? 0.0 : 1.0);
  double sz = (-q.Z < 0.0) ? 0.0 : -100.0;
  o.r0 = (a0 < 0.0 ? 0.0 : a0) + (b0 < 0.0 ? 0.0 : b0);
  o.r1 = (a1 < 0.0 ? 0.0 : a1) + (b1 < 0.0 ? 0.0 : b1);
  o.r2 = (-q.Z < 0.0 ? 0.0 : -q.Z) * 100.0;
  if (!valid) { su = 0.0; sv = 0.0; sz = 0.0; o.r0 = 0.0; o.r1 = 0.0; o.r2 = 0.0; }
  o.s1 = fma(o.r0, o.r0, fma(o.r1, o.r1, fma(o.r2, o.r2, 1.0)));
  o.w = 1.0 / o.s1;
  const double au = su * (cam.fx * q.iz), bu = su * ((q.u - cam.cx) * q.iz);
  const double av = sv * (cam.fy * q.iz), bv = sv * ((q.v - cam.cy) * q.iz);
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    o.JU[k] = au * q.dX[k] - bu * q.dZ[k];
    o.JV[k] = (P == 4) ? -bv * q.dZ[k] : av * q.dY[k] - bv * q.dZ[k];
    o.JZ[k] = sz * q.dZ[k];
  }
  o.JU[NR] = au;  o.JU[NR + 1] = 0.0; o.JU[NR + 2] = -bu;
  o.JV[NR] = 0.0; o.JV[NR + 1] = av;  o.JV[NR + 2] = -bv;
  o.JZ[NR] = 0.0; o.JZ[NR + 1] = 0.0; o.JZ[NR + 2] = sz;
}

// rank-1 update restricted to the entries of J that can be non-zero (compile-time MASK).
template <int P, unsigned MASK, typename ACC>
__device__ __forceinline__ void rank1_masked(ACC acc, const double* J, double w, double r) {
  const double wr = w * r;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    if ((MASK >> j) & 1u) {
      acc[1 + j] = fma(wr, J[j], acc[1 + j]);
      const double wj = w * J[j];
#pragma unroll
      for (int k = j; k < P; ++k)
        if ((MASK >> k) & 1u) acc[1 + P + tri(P, j, k)] = fma(wj, J[k], acc[1 + P + tri(P, j, k)]);
    }
  }
}

template <int P, typename ACC>
__device__ __forceinline__ void accumulate_inside(ACC acc, const Out1<P>& o) {
  constexpr int NR = P - 3;
  constexpr unsigned ROT = (1u << NR) - 1u;
  rank1_masked<P, ROT | (1u << NR) | (1u << (NR + 2)), ACC>(acc, o.JU, o.w, o.r0);
  rank1_masked<P, ROT | (1u << (NR + 1)) | (1u << (NR + 2)), ACC>(acc, o.JV, o.w, o.r1);
  rank1_masked<P, ROT | (1u << (NR + 2)), ACC>(acc, o.JZ, o.w, o.r2);
}

// Running product of (1 + s) kept as mantissa in [1,2) x 2^expo: sum log(1+s) = log(prod) + expo ln 2.
__device__ __forceinline__ void renorm_product(double& prod, int& expo) {
  const int hi = __double2hiint(prod);
  const int e = ((hi >> 20) & 0x7ff) - 1023;
  if (e != 1024) {                       // leave inf / NaN alone so that they propagate
    expo += e;
    prod = __hiloint2double(hi - (e << 20), __double2loint(prod));
  }
}

// ------------------------------------------------------------------------------------------
// Conservative fp32 activity test.  Most points contribute exactly zero to cost, gradient and
// J^T J at a given pose (an "outside" point that projects outside, an "inside" point that
// projects inside).  Each test  u > 0, u < W1, v > 0, v < H1, Z > 0  is, for Z > 0, the sign of
// a linear form in (x, y, z, 1) whose coefficients depend only on the pose and intrinsics:
//     u > 0   <=>  fx X + cx Z        > 0        u < W1  <=>  fx X + (cx - W1) Z < 0
//     v > 0   <=>  fy Y + cy Z        > 0        v < H1  <=>  fy Y + (cy - H1) Z < 0
// The forms are evaluated in fp32 with a per-point error margin m >= 1.6 x the worst-case fp32
// evaluation error (coefficients rounded from fp64, two/three FMAs, inputs possibly rounded from
// fp64).  A point is skipped only if its state is decided by more than m; everything else is
// re-evaluated exactly in fp64, so the sums are those of an all-fp64 evaluation.
// ------------------------------------------------------------------------------------------
struct ClassConst {
  float zc[4], al[4], ah[4], bl[4], bh[4];
  float G, G0;
  int enabled;
};

__device__ __noinline__ void make_class(const PoseConst& pc, const Cam& cam, ClassConst* cc) {
  const double gamma = 8.0 / 16777216.0;    // 8 * 2^-24
  // the five forms kx X + ky Y + kz Z:  Z;  fx X + cx Z;  fx X + (cx - W1) Z;  fy Y + cy Z;  fy Y + (cy - H1) Z
  const double kx[5] = {0.0, cam.fx, cam.fx, 0.0, 0.0};
  const double ky[5] = {0.0, 0.0, 0.0, cam.fy, cam.fy};
  const double kz[5] = {1.0, cam.cx, cam.cx - cam.W1, cam.cy, cam.cy - cam.H1};
  double R[9], t[3];
#pragma unroll
  for (int j = 0; j < 9; ++j) R[j] = pc.R[j];
#pragma unroll
  for (int j = 0; j < 3; ++j) t[j] = pc.t[j];
  float* out = cc->zc;                       // zc, al, ah, bl, bh are contiguous float[4]
  double amax = 0.0, cmax = 0.0;
  float fsum = 0.0f;                         // stays finite iff every coefficient is
#pragma unroll
  for (int f = 0; f < 5; ++f) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double c = fma(kx[f], R[j], fma(ky[f], R[3 + j], kz[f] * R[6 + j]));
      amax = fmax(amax, fabs(c));
      const float cf = (float)c;
      fsum += fabsf(cf);
      out[4 * f + j] = cf;
    }
    const double c3 = fma(kx[f], t[0], fma(ky[f], t[1], kz[f] * t[2]));
    cmax = fmax(cmax, fabs(c3));
    const float c3f = (float)c3;
    fsum += fabsf(c3f);
    out[4 * f + 3] = c3f;
  }
  cc->G = (float)(gamma * amax * 1.0000002);
  cc->G0 = (float)(gamma * cmax * 1.0000002) + 1e-30f;
  cc->enabled = (isfinite(fsum) && isfinite(cc->G) && isfinite(cc->G0)) ? 1 : 0;
}

template <int P>
__device__ __forceinline__ bool maybe_active(float x, float y, float z, int lab, const ClassConst& cc) {
  // branch-free on purpose: the DIB_GPS groups of a step are then classified as one straight-line block that
  // shares the loads of `cc` and interleaves the independent FMA chains
  const bool lab_ok = (unsigned)lab <= 1u;             // labels other than 0/1 carry no residual block
  const float m = fmaf(cc.G, fabsf(x) + fabsf(y) + fabsf(z), cc.G0);
  float Z, al, ah;
  if (P == 4) {                                        // R = Ry: no y terms in X and Z
    Z = fmaf(cc.zc[0], x, fmaf(cc.zc[2], z, cc.zc[3]));
    al = fmaf(cc.al[0], x, fmaf(cc.al[2], z, cc.al[3]));
    ah = fmaf(cc.ah[0], x, fmaf(cc.ah[2], z, cc.ah[3]));
  } else {
    Z = fmaf(cc.zc[0], x, fmaf(cc.zc[1], y, fmaf(cc.zc[2], z, cc.zc[3])));
    al = fmaf(cc.al[0], x, fmaf(cc.al[1], y, fmaf(cc.al[2], z, cc.al[3])));
    ah = fmaf(cc.ah[0], x, fmaf(cc.ah[1], y, fmaf(cc.ah[2], z, cc.ah[3])));
  }
  const float bl = fmaf(cc.bl[0], x, fmaf(cc.bl[1], y, fmaf(cc.bl[2], z, cc.bl[3])));
  const float bh = fmaf(cc.bh[0], x, fmaf(cc.bh[1], y, fmaf(cc.bh[2], z, cc.bh[3])));
  // al > m, ah < -m, bl > m, bh < -m  <=>  lo4 > m ;   al < -m or ah > m or bl < -m or bh > m  <=>  lo4 < -m
  const float lo4 = fminf(fminf(al, -ah), fminf(bl, -bh));
  const bool front = Z > m;
  const bool inside = front && lo4 > m;                    // surely in the image
  const bool outside = (Z < -m) || (front && lo4 < -m);    // surely not
  const bool act = lab == 1 ? !inside : !outside;
  return lab_ok && (cc.enabled ? act : true);
}

// ------------------------------------------------------------------------------------------
// Box table.  The cloud of a sample is static over all of its (inits x evaluations) passes, so
// each group of 32 consecutive points gets an axis-aligned bounding box once per launch
// (frustum_boxes_kernel).  A linear form over a box ranges over  f(centre) +- sum |coef| half,
// so one lane decides a whole group against the five forms; only undecided groups are ever
// loaded point by point.  (frustum_prepare_batch sorts points by (label, Morton cell) so that
// groups are spatially compact and label-pure; unsorted clouds still work, they just cull less.)
//
// Global layout per sample: rounds x [8 fields][kThreads] floats, group id of slot (round r, warp
// w, lane l) = r * kThreads + l * kWarps + w  -- neighbouring groups are dealt round-robin to the
// warps (load balance) while each warp reads consecutive words (no bank conflicts).
// Fields: cx cy cz hx hy hz flags(bit0: has label 0, bit1: has label 1) pad.
// ------------------------------------------------------------------------------------------
constexpr int kBoxFields = 8;
constexpr int kBoxRoundFloats = kBoxFields * kThreads;
#ifndef DIB_BOX_ROUNDS
#define DIB_BOX_ROUNDS 8
#endif
#ifndef DIB_ACC_SMEM
#define DIB_ACC_SMEM 1                        // 1: the per-lane accumulators live in shared memory ([N][32] per warp) instead of
#endif                                        //    registers: fewer spills at 96 registers (-1.8 % kernel time); going on to 12 or
                                              //    16 CTAs/SM (80 / 64 registers) measured slower again
#ifndef DIB_RING
#define DIB_RING 128                          // pending-ring entries per warp and label (power of two)
#endif
#ifndef DIB_PACKED
#define DIB_PACKED 1                          // 1: frustum_boxes_kernel also writes a packed {x, y, z, label} record per point into the
#endif                                        //    workspace and the solver loads an undecided group's points with ONE 16-byte load per lane
                                              //    instead of four (the x/y/z/label arrays of the canonical record are then read once per launch)
#ifndef DIB_GROUP_PIPE
#define DIB_GROUP_PIPE 0                      // 1: software-pipeline the undecided-group loads one step ahead (registers)
#endif
#ifndef DIB_GROUP_TMA
#define DIB_GROUP_TMA 0                       // 1: undecided groups are staged by cp.async.bulk (TMA engine) into a per-warp
#endif                                        //    double buffer one step ahead; 0: plain coalesced loads into registers.
                                              //    Measured on B200, same call (profiles/r01_sweep_build_params.jsonl):
                                              //    TMA 133 ms vs loads 103 ms per 512x60 problems -- 128-byte bulk copies
                                              //    cost more than the 8 coalesced loads they replace.
#ifndef DIB_GPS
#define DIB_GPS 2                             // undecided groups fetched + classified per step
#endif
#ifndef DIB_EXACT_ILP
#define DIB_EXACT_ILP 1                       // exact-path entries per lane per batch (1 or 2); 1 fits 96 registers
                                              // -> 10 CTAs/SM, measured faster than ILP 2 at 8 CTAs/SM
#endif
#ifndef DIB_BOX_SMEM
#define DIB_BOX_SMEM 0                        // 1: table bulk-copied (TMA engine) into shared memory per problem; 0: read via L1/L2
                                              // (measured faster on B200: profiles/r01_sweep_build_params.jsonl)
#endif
constexpr int kBoxRounds = DIB_BOX_SMEM ? DIB_BOX_ROUNDS : 1;   // rounds resident in shared memory (x kThreads x 32 points)

__host__ __device__ inline int box_rounds(int n) { return (((n + 31) >> 5) + kThreads - 1) / kThreads; }

// Self-contained record of a point: the element of the packed per-launch copy (DIB_PACKED) and of the
// per-warp rings of maybe-active points.
template <typename CT> struct Entry;
template <> struct alignas(16) Entry<float> { float x, y, z; int lab; };
template <> struct alignas(16) Entry<double> { double x, y, z; long long lab; };

#ifndef DIB_PK_LDG
#define DIB_PK_LDG 0                          // 1: packed records are read with ld.global.nc (LDG) instead of generic loads
#endif
template <typename CT> __device__ __forceinline__ Entry<CT> load_entry(const Entry<CT>* p);
template <> __device__ __forceinline__ Entry<float> load_entry<float>(const Entry<float>* p) {
#if DIB_PK_LDG
  const int4 v = __ldg(reinterpret_cast<const int4*>(p));
  Entry<float> e;
  e.x = __int_as_float(v.x); e.y = __int_as_float(v.y); e.z = __int_as_float(v.z); e.lab = v.w;
  return e;
#else
  return *p;
#endif
}
template <> __device__ __forceinline__ Entry<double> load_entry<double>(const Entry<double>* p) {
#if DIB_PK_LDG
  const int4 a = __ldg(reinterpret_cast<const int4*>(p)), b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  Entry<double> e;
  e.x = __hiloint2double(a.y, a.x); e.y = __hiloint2double(a.w, a.z); e.z = __hiloint2double(b.y, b.x);
  e.lab = (long long)(((unsigned long long)(unsigned)b.w << 32) | (unsigned)b.z);
  return e;
#else
  return *p;
#endif
}

template <typename CT>
__global__ void __launch_bounds__(256) frustum_boxes_kernel(const CT* __restrict__ xyz,
                                                            const int8_t* __restrict__ label,
                                                            const int32_t* __restrict__ n_pts, int n_stride,
                                                            int rounds_max, float* __restrict__ table,
                                                            Entry<CT>* __restrict__ packed) {
  const int s = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int gid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gid >= rounds_max * kThreads) return;
  const int n = n_pts ? n_pts[s] : n_stride;
  const int i = gid * 32 + lane;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  int lab = -1;
  Entry<CT> e;
  e.x = 0; e.y = 0; e.z = 0; e.lab = -1;
  if (i < n) {
    lab = label[(size_t)s * n_stride + i];
    if (lab == 0 || lab == 1) {
      e.x = xyz[((size_t)s * 3 + 0) * n_stride + i];
      e.y = xyz[((size_t)s * 3 + 1) * n_stride + i];
      e.z = xyz[((size_t)s * 3 + 2) * n_stride + i];
      e.lab = lab;
      lo[0] = hi[0] = (double)e.x; lo[1] = hi[1] = (double)e.y; lo[2] = hi[2] = (double)e.z;
    }
  }
  // packed copy: every slot of the sample's rounds x kThreads x 32 grid is written (padding and ignored labels
  // as label -1), so the solver needs no bounds test
  if (packed) packed[(size_t)s * rounds_max * (kThreads * 32) + i] = e;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
  const unsigned m0 = __ballot_sync(0xffffffffu, lab == 0), m1 = __ballot_sync(0xffffffffu, lab == 1);
  if (lane == 0) {
    const int r = gid / kThreads, q = gid % kThreads;
    const int w = q % kWarps, l = q / kWarps;
    float* rec = table + ((size_t)s * rounds_max + r) * kBoxRoundFloats + (w * 32 + l);
    const int flags = (m0 ? 1 : 0) | (m1 ? 2 : 0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float cf = 0.f, hf = 0.f;
      if (flags) {
        const double cd = 0.5 * (lo[c] + hi[c]);
        cf = (float)cd;
        // half extent measured from the ROUNDED centre, inflated so the fp32 box contains every point
        const double hd = fmax(hi[c] - (double)cf, (double)cf - lo[c]);
        hf = (float)(hd * 1.000001 + (fabs(cd) + hd) * 1.3e-7 + 1e-30);
      }
      rec[c * kThreads] = cf;
      rec[(3 + c) * kThreads] = hf;
    }
    rec[6 * kThreads] = __int_as_float(flags);
    rec[7 * kThreads] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// Shared-memory layout of one CTA.
// ------------------------------------------------------------------------------------------

struct LsSample {
  double x, value, gradient;
  int value_valid, gradient_valid;
};

template <int P>
struct LMState {
  static constexpr int NA = NAcc<P>::NA;
  double x[P], x_norm, cost, g[P], A[NA], grad_max;
  double scale[P], diag[P];
  double radius, dec;
  double step[P], delta[P], mcc;
  double lb[P], ub[P];
  double gd, dmax;
  LsSample lower, prev, cur;
  double xt[P];          // point of the pending evaluation
  int reuse_diag, step_ok, invalid, iteration, max_iter;
  int phase;             // 0 initial, 1 line-search sample, 2 candidate after failed line search
  int ls_iter, evals, ls_steps, term;
  double scratch[72];    // line-search interpolation workspace (shared memory, see interpolating_min_step)
};

constexpr int kBatch = 32 * DIB_EXACT_ILP;   // entries evaluated per exact-path batch
constexpr int kRing = DIB_RING;          // pending ring per label: < kBatch carried + at most DIB_GPS x 32 appended per step
static_assert(kBatch + 32 * DIB_GPS <= kRing, "ring too small");

// Per-warp staging buffer of one step's undecided groups (filled by bulk copies).
template <typename CT>
struct alignas(16) GroupStage {
  CT x[DIB_GPS][32];
  CT y[DIB_GPS][32];
  CT z[DIB_GPS][32];
  int8_t lab[DIB_GPS][32];
};

template <typename CT, int P>
struct Smem {
  // Staging buffers of the optional TMA variants shrink to stubs when those variants are compiled out: with
  // them the CTA needed 23.2 KB and shared memory capped the SM at 9 CTAs although registers allow 10.
  alignas(16) float box[kBoxRounds][DIB_BOX_SMEM ? kBoxRoundFloats : 4];   // bulk-copied (TMA engine) once per problem
  Entry<CT> list[kWarps][2][kRing];                     // [label 0 | label 1] pending rings
#if DIB_ACC_SMEM
  double accs[kWarps][NAcc<P>::N][32];                  // per-lane accumulators (column = lane: conflict-free)
#endif
#if DIB_GROUP_TMA
  GroupStage<CT> gstage[kWarps][2];                     // double-buffered group staging, private to each warp
#endif
  alignas(8) uint64_t gbar[kWarps][2];                  // their mbarriers
  alignas(8) uint64_t full;                             // mbarrier of the box-table copy
  double red[kWarps][NAcc<P>::N];
  double tot[NAcc<P>::N];
  PoseConst pose;
  ClassConst cls;
  Cam cam;
  LMState<P> lm;
  int problem;       // current problem id (broadcast)
  int go;            // 1 = another evaluation requested
};

// Box test of one group by one lane: true if the group needs per-point work.
// One box record as seven registers (cx cy cz hx hy hz flags).
struct BoxRec {
  float v[7];
};

template <int SMEM>
__device__ __forceinline__ void box_load(const float* f, int slot, BoxRec& b) {
#pragma unroll
  for (int k = 0; k < 7; ++k) b.v[k] = SMEM ? f[k * kThreads + slot] : __ldg(f + k * kThreads + slot);
}

// Box test of one group by one lane: true if the group needs per-point work.
__device__ __forceinline__ bool box_undecided(const BoxRec& b, const ClassConst& cc) {
  const int flags = __float_as_int(b.v[6]);
  if (flags == 0) return false;                       // no point with a residual block
  if (!cc.enabled) return true;
  const float cx = b.v[0], cy = b.v[1], cz = b.v[2], hx = b.v[3], hy = b.v[4], hz = b.v[5];
  const float m = 2.0f * fmaf(cc.G, (fabsf(cx) + hx) + (fabsf(cy) + hy) + (fabsf(cz) + hz), cc.G0);
  float lo[5], hi[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float* c = cc.zc + 4 * k;                   // zc, al, ah, bl, bh are contiguous float[4]
    const float mid = fmaf(c[0], cx, fmaf(c[1], cy, fmaf(c[2], cz, c[3])));
    const float rad = fmaf(fabsf(c[0]), hx, fmaf(fabsf(c[1]), hy, fabsf(c[2]) * hz));
    lo[k] = mid - rad; hi[k] = mid + rad;
  }
  const bool front = lo[0] > m;
  const bool all_out = (hi[0] < -m) || (front && fminf(fminf(hi[1], -lo[2]), fminf(hi[3], -lo[4])) < -m);
  const bool all_in = front && fminf(fminf(lo[1], -hi[2]), fminf(lo[3], -hi[4])) > m;
  const bool skip = (!(flags & 1) || all_out) && (!(flags & 2) || all_in);
  return !skip;
}

// ------------------------------------------------------------------------------------------
// One pass over a cloud (all threads must call): box tests -> coalesced loads of the undecided
// groups -> per-point fp32 culling -> ordered compaction into the warp's pending rings -> exact
// fp64 evaluation of 64 pending points at a time (two per lane, label-homogeneous) -> fixed-order
// block reduction into sm.tot.  Warps never synchronise with each other inside the pass.
// `boxes_resident` says the sample's whole table already sits in sm.box.
// ------------------------------------------------------------------------------------------
template <typename CT, int P>
__device__ void evaluate_cloud(Smem<CT, P>& sm, const CT* xyz_s, const int8_t* lab_s, const Entry<CT>* pk_s,
                               int n_stride, int n, const float* box_s, bool boxes_resident, uint32_t& box_phase, uint32_t& gphase) {
  constexpr int N = NAcc<P>::N;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int rounds = box_rounds(n);
#if DIB_ACC_SMEM
  SmemAcc acc{&sm.accs[warp][0][lane]};
#else
  double acc[N];
#endif
#pragma unroll
  for (int j = 0; j < N; ++j) acc[j] = 0.0;
#if DIB_ACC_SMEM
  const SmemAcc acc_view = acc;
#else
  double* const acc_view = acc;
#endif
  const Cam& cam = sm.cam;
  const PoseConst& pc = sm.pose;
  const ClassConst& cc = sm.cls;
  Entry<CT>* ring0 = sm.list[warp][0];
  Entry<CT>* ring1 = sm.list[warp][1];
  int head0 = 0, pend0 = 0, head1 = 0, pend1 = 0;   // warp-uniform
  double prod = 1.0;                                  // per-lane product of (1 + s), renormalised
  int expo = 0;
  BoxRec box_cur, box_nxt;
#pragma unroll
  for (int k = 0; k < 7; ++k) { box_cur.v[k] = 0.f; box_nxt.v[k] = 0.f; }
#if !DIB_BOX_SMEM
  if (rounds > 0) box_load<0>(box_s, warp * 32 + lane, box_nxt);
#endif

  // Rounds r = 0 .. rounds-1 test one box per lane; the extra last round only drains what is left,
  // so each exact evaluator has ONE code instance.
#pragma unroll 1
  for (int r = 0; r <= rounds; ++r) {
    unsigned mask = 0;
    int threshold = 1;
    if (r < rounds) {
      threshold = kBatch;
      if (DIB_BOX_SMEM && !boxes_resident && (r % kBoxRounds) == 0) {
        // clouds larger than the resident window: stream the table chunk by chunk, every pass
        __syncthreads();
        if (tid == 0) {
          int nr = rounds - r;
          if (nr > kBoxRounds) nr = kBoxRounds;
          const uint32_t bytes = (uint32_t)nr * kBoxRoundFloats * sizeof(float);
          mbar_expect_tx(&sm.full, bytes);
          bulk_g2s(&sm.box[0][0], box_s + (size_t)r * kBoxRoundFloats, bytes, &sm.full);
        }
        mbar_wait(&sm.full, box_phase & 1);
        ++box_phase;
      }
#if DIB_BOX_SMEM
      box_load<1>(sm.box[r % kBoxRounds], warp * 32 + lane, box_cur);
#else
      box_cur = box_nxt;                              // loaded while the previous round was processed
      if (r + 1 < rounds) box_load<0>(box_s + (size_t)(r + 1) * kBoxRoundFloats, warp * 32 + lane, box_nxt);
#endif
      mask = __ballot_sync(0xffffffffu, box_undecided(box_cur, cc));
    }
#if DIB_GROUP_TMA
    // Undecided groups are taken DIB_GPS at a time and STAGED BY THE TMA ENGINE: lane 0 issues one
    // cp.async.bulk per coordinate array and group (128 B of x, y, z and 32 B of labels) into this warp's
    // double buffer, one step ahead of its use, completion signalled on the slot's mbarrier.  While the
    // copies fly the warp drains pending exact-path batches; then it classifies the staged groups
    // (independent instruction streams) and appends.  Ring bound: < kBatch carried + 32 x DIB_GPS new.
    int cur_cnt = 0, cur_g[DIB_GPS], slot = 0;
    auto take_and_issue = [&](int which, int* g_out) -> int {
      int cnt = 0;
#pragma unroll
      for (int u = 0; u < DIB_GPS; ++u) {
        g_out[u] = -1;
        if (mask) {
          const int b = __ffs(mask) - 1;
          mask &= mask - 1;
          g_out[u] = r * kThreads + b * kWarps + warp;
          ++cnt;
        }
      }
      if (cnt && lane == 0) {
        GroupStage<CT>& gs = sm.gstage[warp][which];
        uint64_t* bar = &sm.gbar[warp][which];
        uint32_t bytes = 0;
#pragma unroll
        for (int u = 0; u < DIB_GPS; ++u)
          if (g_out[u] >= 0) {
            int c = n_stride - g_out[u] * 32;           // points of this group inside the row (multiple of 16)
            c = c > 32 ? 32 : c;
            bytes += (uint32_t)c * (3u * sizeof(CT) + 1u);
          }
        mbar_expect_tx(bar, bytes);
#pragma unroll
        for (int u = 0; u < DIB_GPS; ++u)
          if (g_out[u] >= 0) {
            const int base = g_out[u] * 32;
            int c = n_stride - base;
            c = c > 32 ? 32 : c;
            bulk_g2s(gs.x[u], xyz_s + base, (uint32_t)c * sizeof(CT), bar);
            bulk_g2s(gs.y[u], xyz_s + n_stride + base, (uint32_t)c * sizeof(CT), bar);
            bulk_g2s(gs.z[u], xyz_s + 2 * (size_t)n_stride + base, (uint32_t)c * sizeof(CT), bar);
            bulk_g2s(gs.lab[u], lab_s + base, (uint32_t)c, bar);
          }
      }
      return cnt;
    };
    if (mask) cur_cnt = take_and_issue(slot, cur_g);
#pragma unroll 1
    do {
      int nxt_cnt = 0, nxt_g[DIB_GPS];
#pragma unroll
      for (int u = 0; u < DIB_GPS; ++u) nxt_g[u] = -1;
      if (mask) nxt_cnt = take_and_issue(slot ^ 1, nxt_g);
#pragma unroll 1
      while (pend0 >= threshold) {
        __syncwarp();
        const int take = pend0 < kBatch ? pend0 : kBatch;
        const Entry<CT> ea = ring0[(head0 + lane) & (kRing - 1)];
        Out0<P> oa;
        eval_outside<P>((double)ea.x, (double)ea.y, (double)ea.z, lane < take, cam, pc, oa);
#if DIB_EXACT_ILP == 2
        const Entry<CT> eb = ring0[(head0 + 32 + lane) & (kRing - 1)];
        Out0<P> ob;
        eval_outside<P>((double)eb.x, (double)eb.y, (double)eb.z, lane + 32 < take, cam, pc, ob);
        prod *= oa.s1 * ob.s1;
#else
        prod *= oa.s1;
#endif
        renorm_product(prod, expo);
        rank1<P, decltype(acc_view)>(acc_view, oa.J, oa.w, oa.r);
#if DIB_EXACT_ILP == 2
        rank1<P, decltype(acc_view)>(acc_view, ob.J, ob.w, ob.r);
#endif
        head0 = (head0 + take) & (kRing - 1);
        pend0 -= take;
      }
#pragma unroll 1
      while (pend1 >= threshold) {
        __syncwarp();
        const int take = pend1 < kBatch ? pend1 : kBatch;
        const Entry<CT> ea = ring1[(head1 + lane) & (kRing - 1)];
        Out1<P> oa;
        eval_inside<P>((double)ea.x, (double)ea.y, (double)ea.z, lane < take, cam, pc, oa);
#if DIB_EXACT_ILP == 2
        const Entry<CT> eb = ring1[(head1 + 32 + lane) & (kRing - 1)];
        Out1<P> ob;
        eval_inside<P>((double)eb.x, (double)eb.y, (double)eb.z, lane + 32 < take, cam, pc, ob);
        prod *= oa.s1 * ob.s1;
#else
        prod *= oa.s1;
#endif
        renorm_product(prod, expo);
        accumulate_inside<P, decltype(acc_view)>(acc_view, oa);
#if DIB_EXACT_ILP == 2
        accumulate_inside<P, decltype(acc_view)>(acc_view, ob);
#endif
        head1 = (head1 + take) & (kRing - 1);
        pend1 -= take;
      }
      if (cur_cnt) {
        mbar_wait(&sm.gbar[warp][slot], (gphase >> slot) & 1u);
        gphase ^= 1u << slot;
        const GroupStage<CT>& gs = sm.gstage[warp][slot];
        CT gx[DIB_GPS], gy[DIB_GPS], gz[DIB_GPS];
        int glab[DIB_GPS];
        bool mb[DIB_GPS];
#pragma unroll
        for (int u = 0; u < DIB_GPS; ++u) {
          const bool in = cur_g[u] >= 0 && cur_g[u] * 32 + lane < n;
          gx[u] = in ? gs.x[u][lane] : (CT)0; gy[u] = in ? gs.y[u][lane] : (CT)0; gz[u] = in ? gs.z[u][lane] : (CT)0;
          glab[u] = in ? (int)gs.lab[u][lane] : -1;
        }
#pragma unroll
        for (int u = 0; u < DIB_GPS; ++u) mb[u] = maybe_active<P>((float)gx[u], (float)gy[u], (float)gz[u], glab[u], cc);
#pragma unroll
        for (int u = 0; u < DIB_GPS; ++u) {
          const unsigned m1 = __ballot_sync(0xffffffffu, mb[u] && glab[u] == 1);
          const unsigned m0 = __ballot_sync(0xffffffffu, mb[u] && glab[u] == 0);
          if (mb[u]) {
            Entry<CT> e;
            e.x = gx[u]; e.y = gy[u]; e.z = gz[u]; e.lab = glab[u];
            if (glab[u]) ring1[(head1 + pend1 + __popc(m1 & lt_mask)) & (kRing - 1)] = e;
            else         ring0[(head0 + pend0 + __popc(m0 & lt_mask)) & (kRing - 1)] = e;
          }
          pend0 += __popc(m0);
          pend1 += __popc(m1);
        }
      }
      cur_cnt = nxt_cnt;
#pragma unroll
      for (int u = 0; u < DIB_GPS; ++u) cur_g[u] = nxt_g[u];
      slot ^= 1;
    } while (cur_cnt);
#else
    // Undecided groups are taken DIB_GPS at a time.  Their loads are issued first, then the pending
    // exact-path batches are drained WHILE THE LOADS ARE IN FLIGHT, then the groups are classified
    // (independent instruction streams) and appended.  Ring bound: < kBatch carried + 32 x DIB_GPS new.
    auto load_groups = [&](CT* lx, CT* ly, CT* lz, int* ll) -> bool {
      const bool any = mask != 0;
      if (any) {
#pragma unroll
        for (int u = 0; u < DIB_GPS; ++u) {
          ll[u] = -1; lx[u] = 0; ly[u] = 0; lz[u] = 0;
          if (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const int i = (r * kThreads + b * kWarps + warp) * 32 + lane;      // this lane's point
#if DIB_PACKED
            const Entry<CT> e = load_entry<CT>(pk_s + i);
            ll[u] = (int)e.lab; lx[u] = e.x; ly[u] = e.y; lz[u] = e.z;
#else
            if (i < n) {
              ll[u] = lab_s[i];
              lx[u] = xyz_s[i]; ly[u] = xyz_s[n_stride + i]; lz[u] = xyz_s[2 * (size_t)n_stride + i];
            }
#endif
          }
        }
      }
      return any;
    };
#if DIB_GROUP_PIPE
    // register software pipeline: the NEXT DIB_GPS groups are loaded before the current ones are classified,
    // so the L2 latency of a group is covered by the classification of the previous one as well
    CT gx[DIB_GPS], gy[DIB_GPS], gz[DIB_GPS];
    int glab[DIB_GPS];
    bool have = load_groups(gx, gy, gz, glab);
#endif
#pragma unroll 1
    do {
#if DIB_GROUP_PIPE
      CT nx[DIB_GPS], ny[DIB_GPS], nz[DIB_GPS];
      int nlab[DIB_GPS];
      const bool nhave = load_groups(nx, ny, nz, nlab);
#else
      CT gx[DIB_GPS], gy[DIB_GPS], gz[DIB_GPS];
      int glab[DIB_GPS];
      const bool have = load_groups(gx, gy, gz, glab);
#endif
#pragma unroll 1
      while (pend0 >= threshold) {
        __syncwarp();
        const int take = pend0 < kBatch ? pend0 : kBatch;
        const Entry<CT> ea = ring0[(head0 + lane) & (kRing - 1)];
        Out0<P> oa;
        eval_outside<P>((double)ea.x, (double)ea.y, (double)ea.z, lane < take, cam, pc, oa);
#if DIB_EXACT_ILP == 2
        const Entry<CT> eb = ring0[(head0 + 32 + lane) & (kRing - 1)];
        Out0<P> ob;
        eval_outside<P>((double)eb.x, (double)eb.y, (double)eb.z, lane + 32 < take, cam, pc, ob);
        prod *= oa.s1 * ob.s1;
#else
        prod *= oa.s1;
#endif
        renorm_product(prod, expo);
        rank1<P, decltype(acc_view)>(acc_view, oa.J, oa.w, oa.r);
#if DIB_EXACT_ILP == 2
        rank1<P, decltype(acc_view)>(acc_view, ob.J, ob.w, ob.r);
#endif
        head0 = (head0 + take) & (kRing - 1);
        pend0 -= take;
      }
#pragma unroll 1
      while (pend1 >= threshold) {
        __syncwarp();
        const int take = pend1 < kBatch ? pend1 : kBatch;
        const Entry<CT> ea = ring1[(head1 + lane) & (kRing - 1)];
        Out1<P> oa;
        eval_inside<P>((double)ea.x, (double)ea.y, (double)ea.z, lane < take, cam, pc, oa);
#if DIB_EXACT_ILP == 2
        const Entry<CT> eb = ring1[(head1 + 32 + lane) & (kRing - 1)];
        Out1<P> ob;
        eval_inside<P>((double)eb.x, (double)eb.y, (double)eb.z, lane + 32 < take, cam, pc, ob);
        prod *= oa.s1 * ob.s1;
#else
        prod *= oa.s1;
#endif
        renorm_product(prod, expo);
        accumulate_inside<P, decltype(acc_view)>(acc_view, oa);
#if DIB_EXACT_ILP == 2
        accumulate_inside<P, decltype(acc_view)>(acc_view, ob);
#endif
        head1 = (head1 + take) & (kRing - 1);
        pend1 -= take;
      }
      if (have) {
        bool mb[DIB_GPS];
#pragma unroll
        for (int u = 0; u < DIB_GPS; ++u) mb[u] = maybe_active<P>((float)gx[u], (float)gy[u], (float)gz[u], glab[u], cc);
#pragma unroll
        for (int u = 0; u < DIB_GPS; ++u) {
          const unsigned m1 = __ballot_sync(0xffffffffu, mb[u] && glab[u] == 1);
          const unsigned m0 = __ballot_sync(0xffffffffu, mb[u] && glab[u] == 0);
          if (mb[u]) {
            // one select + one store (the label picks ring, base and mask) instead of two predicated copies
            const bool l1 = glab[u] != 0;
            Entry<CT>* ring = l1 ? ring1 : ring0;
            const int at = (l1 ? head1 + pend1 : head0 + pend0) + __popc((l1 ? m1 : m0) & lt_mask);
            Entry<CT> e;
            e.x = gx[u]; e.y = gy[u]; e.z = gz[u]; e.lab = glab[u];
            ring[at & (kRing - 1)] = e;
          }
          pend0 += __popc(m0);
          pend1 += __popc(m1);
        }
      }
#if DIB_GROUP_PIPE
      if (nhave) {
#pragma unroll
        for (int u = 0; u < DIB_GPS; ++u) { gx[u] = nx[u]; gy[u] = ny[u]; gz[u] = nz[u]; glab[u] = nlab[u]; }
      }
      have = nhave;
    } while (have);
#else
    } while (mask);
#endif
#endif
  }
  acc[0] = 0.5 * (log(prod) + (double)expo * 0.6931471805599453094);

  // Fixed-order reduction.  The pending rings are empty now, so each warp reuses its ring memory as a
  // [N][33] scratch: lane j sums accumulator j over lanes 0..31 in order, then thread j sums the
  // warps in order.  (A loop over shared memory instead of ~300 unrolled shuffles: this epilogue
  // runs once per pass and its code size matters more than its speed.)
#if DIB_ACC_SMEM
  __syncwarp();
  if (lane < N) {                                   // the accumulators already sit in shared memory as [N][32]
    double v = 0.0;
#pragma unroll 4
    for (int l = 0; l < 32; ++l) v += sm.accs[warp][lane][l];
    sm.red[warp][lane] = v;
  }
#else
  double* scratch = reinterpret_cast<double*>(sm.list[warp]);
  static_assert(sizeof(Entry<CT>) * 2 * kRing >= sizeof(double) * N * 33, "ring too small for the reduction scratch");
  __syncwarp();
#pragma unroll
  for (int j = 0; j < N; ++j) scratch[j * 33 + lane] = acc[j];
  __syncwarp();
  if (lane < N) {
    double v = 0.0;
#pragma unroll 4
    for (int l = 0; l < 32; ++l) v += scratch[lane * 33 + l];
    sm.red[warp][lane] = v;
  }
#endif
  __syncthreads();
  if (tid < N) {
    double v = sm.red[0][tid];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) v += sm.red[w][tid];
    sm.tot[tid] = v;
  }
  __syncthreads();
}

// Loads a sample's whole box table into shared memory if it fits the resident window (all
// threads call; returns true if resident).
template <typename CT, int P>
__device__ bool load_boxes(Smem<CT, P>& sm, const float* box_s, int n, uint32_t& box_phase) {
  const int rounds = box_rounds(n);
  if (!DIB_BOX_SMEM) return true;
  if (rounds > kBoxRounds) return false;
  if (rounds > 0) {
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)rounds * kBoxRoundFloats * sizeof(float);
      mbar_expect_tx(&sm.full, bytes);
      bulk_g2s(&sm.box[0][0], box_s, bytes, &sm.full);
    }
    mbar_wait(&sm.full, box_phase & 1);
    ++box_phase;
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// Trust-region control flow (thread 0 only).
// ------------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void project_plus(const LMState<P>& st, const double* x, const double* d, double a, double* out) {
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double v = fma(a, d[j], x[j]);
    v = fmax(v, st.lb[j]);
    v = fmin(v, st.ub[j]);
    out[j] = v;
  }
}

template <int P>
__device__ __forceinline__ double grad_max_norm(const LMState<P>& st, const double* x, const double* g) {
  double mx = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double v = x[j] - g[j];
    v = fmax(v, st.lb[j]);
    v = fmin(v, st.ub[j]);
    mx = fmax(mx, fabs(x[j] - v));
  }
  return mx;
}

// Solve (As + diag(d2)) y = gs by Cholesky, As full symmetric P x P in registers.  One rsqrt per
// column and no divisions (the control code is latency-bound on one thread).  false if not SPD.
template <int P>
__device__ __forceinline__ bool chol_solve(const double (&As)[P][P], const double* d2, const double* gs, double* y) {
  double L[P][P], inv[P];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    double s = As[j][j] + d2[j];
#pragma unroll
    for (int k = 0; k < j; ++k) s = fma(-L[j][k], L[j][k], s);
    if (!(s > 0.0)) ok = false;
    inv[j] = rsqrt(s);
    L[j][j] = s * inv[j];
#pragma unroll
    for (int i = j + 1; i < P; ++i) {
      double t = As[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) t = fma(-L[i][k], L[j][k], t);
      L[i][j] = t * inv[j];
    }
  }
  double z[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    double t = gs[i];
#pragma unroll
    for (int k = 0; k < i; ++k) t = fma(-L[i][k], z[k], t);
    z[i] = t * inv[i];
  }
#pragma unroll
  for (int i = P - 1; i >= 0; --i) {
    double t = z[i];
#pragma unroll
    for (int k = i + 1; k < P; ++k) t = fma(-L[k][i], y[k], t);
    y[i] = t * inv[i];
  }
  return ok;
}

__device__ double poly_eval(const double* p, int n, double x) {   // n coefficients, highest first
  double v = 0.0;
  #pragma unroll 1
  for (int i = 0; i < n; ++i) v = v * x + p[i];
  return v;
}

__device__ double ipow(double x, int k) {
  double v = 1.0;
  #pragma unroll 1
  for (int i = 0; i < k; ++i) v *= x;
  return v;
}

// Durand-Kerner iteration for a polynomial of exact degree DEG (coefficients highest first, p[0] != 0).
// Everything lives in registers (fully unrolled); returns DEG real parts.
template <int DEG>
__device__ __noinline__ int durand_kerner(const double* p, double* roots) {
  double c[DEG + 1], zr[DEG], zi[DEG];
  const double ip0 = 1.0 / p[0];
#pragma unroll
  for (int i = 0; i <= DEG; ++i) c[i] = p[i] * ip0;
  // Fujiwara's bound on the root moduli: 2 max_k |c_k|^(1/k) (the last coefficient halved)
  double radius = 0.0;
#pragma unroll
  for (int i = 1; i <= DEG; ++i) {
    const double a = fabs(c[i]) * (i == DEG ? 0.5 : 1.0);
    radius = fmax(radius, a > 0.0 ? exp(log(a) / (double)i) : 0.0);
  }
  radius = 2.0 * radius + 1e-300;
  // start points on the circle of half that radius, fixed phases (cos/sin of 2 pi i / DEG + 0.4)
#pragma unroll
  for (int i = 0; i < DEG; ++i) {
    double s, co;
    sincos(2.0 * 3.14159265358979323846 * i / DEG + 0.4, &s, &co);
    zr[i] = 0.5 * radius * co; zi[i] = 0.5 * radius * s;
  }
#pragma unroll 1
  for (int it = 0; it < 100; ++it) {
    double change = 0.0;
#pragma unroll
    for (int i = 0; i < DEG; ++i) {
      double nr = 0.0, ni = 0.0;
#pragma unroll
      for (int k = 0; k <= DEG; ++k) {
        const double tr = nr * zr[i] - ni * zi[i] + c[k];
        const double ti = nr * zi[i] + ni * zr[i];
        nr = tr; ni = ti;
      }
      double dr = 1.0, di = 0.0;
#pragma unroll
      for (int j = 0; j < DEG; ++j) if (j != i) {
        const double er = zr[i] - zr[j], ei = zi[i] - zi[j];
        const double tr = dr * er - di * ei, ti = dr * ei + di * er;
        dr = tr; di = ti;
      }
      double den = dr * dr + di * di;
      if (den == 0.0) { dr = 1e-300; di = 0.0; den = dr * dr; if (den == 0.0) den = 1e-300; }
      const double iden = 1.0 / den;
      const double qr = (nr * dr + ni * di) * iden, qi = (ni * dr - nr * di) * iden;
      zr[i] -= qr; zi[i] -= qi;
      change = fmax(change, sqrt(qr * qr + qi * qi));
    }
    if (change < 1e-14 * radius) break;
  }
#pragma unroll
  for (int i = 0; i < DEG; ++i) roots[i] = zr[i];
  return DEG;
}

// Real parts of all roots of p (n coefficients, highest first; n - 1 <= 4).
__device__ __noinline__ int poly_roots_real(const double* pin, int n, double* roots) {
  double p[6];
  int lead = 0;
  #pragma unroll 1
  while (lead < n && pin[lead] == 0.0) ++lead;
  const int m = n - lead;
  #pragma unroll 1
  for (int i = 0; i < m; ++i) p[i] = pin[lead + i];
  const int deg = m - 1;
  if (deg <= 0) return 0;
  if (deg == 1) { roots[0] = -p[1] / p[0]; return 1; }
  if (deg == 2) {
    const double a = p[0], b = p[1], c = p[2];
    const double D = b * b - 4 * a * c;
    const double sD = sqrt(fabs(D));
    if (D >= 0) {
      if (b >= 0) { roots[0] = (-b - sD) / (2.0 * a); roots[1] = (2.0 * c) / (-b - sD); }
      else { roots[0] = (2.0 * c) / (-b + sD); roots[1] = (-b + sD) / (2.0 * a); }
    } else { roots[0] = -b / (2.0 * a); roots[1] = roots[0]; }
    return 2;
  }
  // Durand-Kerner on the monic polynomial (all roots, complex), register-resident for the
  // quartic that the 3-sample interpolation produces; DEG is a compile-time bound.
  if (deg == 4) return durand_kerner<4>(p, roots);
  return durand_kerner<3>(p, roots);
}

// Step size minimising the polynomial that interpolates the line-search samples over
// [xmin, xmax] (cubic interpolation: values and gradients of lower / current / previous).
// `scratch` (>= 72 doubles, shared memory): the dynamically indexed 6x6 system must not live in
// local memory, whose lines get evicted from L1 by the streaming loads (long-scoreboard stalls).
__device__ __noinline__ double interpolating_min_step(const LsSample& lower, const LsSample& previous, const LsSample& current,
                                         double xmin, double xmax, double* scratch) {
  if (!current.value_valid) return fmin(fmax(current.x * 0.5, xmin), xmax);
  const LsSample* s[3] = {&lower, &current, &previous};
  const int ns = previous.value_valid ? 3 : 2;
  int nc = 0;
  #pragma unroll 1
  for (int i = 0; i < ns; ++i) { if (s[i]->value_valid) ++nc; if (s[i]->gradient_valid) ++nc; }
  const int deg = nc - 1;
  double (*M)[6] = reinterpret_cast<double (*)[6]>(scratch);   // [6][6]
  double* rhs = scratch + 36;
  double* poly = scratch + 42;
  double* z = scratch + 48;
  double* der = scratch + 54;
  double* roots = scratch + 60;
  double* pw = scratch + 66;
  #pragma unroll 1
  for (int i = 0; i < 6; ++i) { rhs[i] = 0; poly[i] = 0; for (int j = 0; j < 6; ++j) M[i][j] = 0; }
  int row = 0;
  #pragma unroll 1
  for (int i = 0; i < ns; ++i) {
    pw[0] = 1.0;                        // pw[k] = x^k by repeated multiplication
    #pragma unroll 1
    for (int k = 1; k <= deg; ++k) pw[k] = pw[k - 1] * s[i]->x;
    if (s[i]->value_valid) {
      #pragma unroll 1
      for (int j = 0; j <= deg; ++j) M[row][j] = pw[deg - j];
      rhs[row] = s[i]->value; ++row;
    }
    if (s[i]->gradient_valid) {
      #pragma unroll 1
      for (int j = 0; j < deg; ++j) M[row][j] = (deg - j) * pw[deg - j - 1];
      rhs[row] = s[i]->gradient; ++row;
    }
  }
  // full-pivot elimination (column permutation packed in one register, 4 bits per entry)
  unsigned perm = 0x543210u;
  #pragma unroll 1
  for (int k = 0; k < nc; ++k) {
    int pr = k, pcv = k; double best = -1.0;
    #pragma unroll 1
    for (int i = k; i < nc; ++i) for (int j = k; j < nc; ++j)
      if (fabs(M[i][j]) > best) { best = fabs(M[i][j]); pr = i; pcv = j; }
    if (best == 0.0) { for (int i = k; i < nc; ++i) rhs[i] = 0.0; break; }
    if (pr != k) { for (int j = 0; j < nc; ++j) { double t = M[pr][j]; M[pr][j] = M[k][j]; M[k][j] = t; } double t = rhs[pr]; rhs[pr] = rhs[k]; rhs[k] = t; }
    if (pcv != k) { for (int i = 0; i < nc; ++i) { double t = M[i][pcv]; M[i][pcv] = M[i][k]; M[i][k] = t; } const unsigned pa = (perm >> (4 * pcv)) & 15u, pb = (perm >> (4 * k)) & 15u;
                    perm = (perm & ~((15u << (4 * pcv)) | (15u << (4 * k)))) | (pb << (4 * pcv)) | (pa << (4 * k)); }
    #pragma unroll 1
    for (int i = k + 1; i < nc; ++i) {
      const double f = M[i][k] / M[k][k];
      #pragma unroll 1
      for (int j = k; j < nc; ++j) M[i][j] -= f * M[k][j];
      rhs[i] -= f * rhs[k];
    }
  }
  #pragma unroll 1
  for (int k = nc - 1; k >= 0; --k) {
    if (M[k][k] == 0.0) { z[k] = 0.0; continue; }
    double sacc = rhs[k];
    #pragma unroll 1
    for (int j = k + 1; j < nc; ++j) sacc -= M[k][j] * z[j];
    z[k] = sacc / M[k][k];
  }
  #pragma unroll 1
  for (int k = 0; k < nc; ++k) poly[(perm >> (4 * k)) & 15u] = z[k];

  double best_x = (xmin + xmax) / 2.0;
  double best_v = poly_eval(poly, nc, best_x);
  const double vmin = poly_eval(poly, nc, xmin);
  if (vmin < best_v) { best_v = vmin; best_x = xmin; }
  const double vmax = poly_eval(poly, nc, xmax);
  if (vmax < best_v) { best_v = vmax; best_x = xmax; }
  if (nc <= 2) return best_x;
  #pragma unroll 1
  for (int j = 0; j < deg; ++j) der[j] = (deg - j) * poly[j];
  const int nr = poly_roots_real(der, deg, roots);
  #pragma unroll 1
  for (int i = 0; i < nr; ++i) {
    if (roots[i] < xmin || roots[i] > xmax) continue;
    const double v = poly_eval(poly, nc, roots[i]);
    if (v < best_v) { best_v = v; best_x = roots[i]; }
  }
  return best_x;
}

enum { LM_DONE = 0, LM_EVAL = 1 };

// Starts the next trust-region iteration(s) until an evaluation is needed or the solve ends.
template <int P>
__device__ __noinline__ int lm_next_step(LMState<P>& st) {
  for (;;) {
    if (st.iteration >= st.max_iter) { st.term = 3; return LM_DONE; }
    if (st.step_ok && st.grad_max <= 1e-10) { st.term = 0; return LM_DONE; }
    if (st.radius <= 1e-32) { st.term = 4; return LM_DONE; }
    ++st.iteration;
    st.step_ok = 0;

    double As[P][P], gs[P], d2[P], y[P], sc[P];
#pragma unroll
    for (int j = 0; j < P; ++j) { sc[j] = st.scale[j]; gs[j] = st.g[j] * sc[j]; }
    {
      int idx = 0;
#pragma unroll
      for (int j = 0; j < P; ++j)
#pragma unroll
        for (int k = j; k < P; ++k) {
          const double v = st.A[idx++] * sc[j] * sc[k];
          As[j][k] = v; As[k][j] = v;
        }
    }
    if (!st.reuse_diag) {
#pragma unroll
      for (int j = 0; j < P; ++j) st.diag[j] = fmin(fmax(As[j][j], 1e-6), 1e32);
    }
    st.reuse_diag = 1;
    {
      const double inv_radius = 1.0 / st.radius;      // D^2 = diag / radius (LM damping)
#pragma unroll
      for (int j = 0; j < P; ++j) d2[j] = st.diag[j] * inv_radius;
    }
    bool ok = chol_solve<P>(As, d2, gs, y);
    double mcc = 0.0;
    if (ok) {
      // model cost change = -step^T gs - 1/2 step^T As step, step = -y
      double quad = 0.0, lin = 0.0;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        st.step[j] = -y[j];
        if (!isfinite(y[j])) ok = false;
      }
#pragma unroll
      for (int j = 0; j < P; ++j) {
        lin = fma(-y[j], gs[j], lin);
        double rowsum = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) rowsum = fma(As[j][k], -y[k], rowsum);
        quad = fma(-y[j], rowsum, quad);
      }
      mcc = -lin - 0.5 * quad;
    }
    if (!ok || !(mcc > 0.0)) {
      if (++st.invalid >= 5) { st.term = 5; return LM_DONE; }
      st.radius *= 0.5;
      st.reuse_diag = 1;
      continue;
    }
    st.invalid = 0;
    st.mcc = mcc;
    double gd = 0.0, dmax = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      st.delta[j] = st.step[j] * st.scale[j];
      gd += st.g[j] * st.delta[j];
      dmax = fmax(dmax, fabs(st.delta[j]));
    }
    st.gd = gd; st.dmax = dmax;
    st.lower.x = 0.0; st.lower.value = st.cost; st.lower.gradient = gd; st.lower.value_valid = 1; st.lower.gradient_valid = 1;
    st.prev.value_valid = 0; st.prev.gradient_valid = 0; st.prev.x = 0; st.prev.value = 0; st.prev.gradient = 0;
    st.cur.x = 1.0;
    st.ls_iter = 0;
    project_plus<P>(st, st.x, st.delta, 1.0, st.xt);
    st.phase = 1;
    return LM_EVAL;
  }
}

template <int P>
__device__ __noinline__ int lm_after_candidate(LMState<P>& st, const double* tot) {
  constexpr int NA = NAcc<P>::NA;
  const double ccost = tot[0];
  double sn = 0.0;
#pragma unroll
  for (int j = 0; j < P; ++j) sn += (st.x[j] - st.xt[j]) * (st.x[j] - st.xt[j]);
  if (sqrt(sn) <= 1e-8 * (st.x_norm + 1e-8)) { st.term = 1; return LM_DONE; }
  if (fabs(st.cost - ccost) <= 1e-6 * st.cost) { st.term = 2; return LM_DONE; }
  const double rho = (st.cost - ccost) / st.mcc;
  if (rho > 1e-3) {
    double xn = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) { st.x[j] = st.xt[j]; xn += st.x[j] * st.x[j]; st.g[j] = tot[1 + j]; }
    st.x_norm = sqrt(xn);
    st.cost = ccost;
#pragma unroll
    for (int j = 0; j < NA; ++j) st.A[j] = tot[1 + P + j];
    st.grad_max = grad_max_norm<P>(st, st.x, st.g);
    const double t = 2.0 * rho - 1.0;
    st.radius = st.radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
    st.radius = fmin(1e16, st.radius);
    st.dec = 2.0;
    st.reuse_diag = 0;
    st.step_ok = 1;
  } else {
    st.radius = st.radius / st.dec;
    st.dec *= 2.0;
    st.reuse_diag = 1;
  }
  return lm_next_step<P>(st);
}

// Consumes the evaluation at st.xt (totals in tot).  Returns LM_EVAL with a new st.xt or LM_DONE.
template <int P>
__device__ __noinline__ int lm_consume(LMState<P>& st, const double* tot) {
  constexpr int NA = NAcc<P>::NA;
  ++st.evals;
  if (st.phase == 0) {
    st.cost = tot[0];
    #pragma unroll 1
    for (int j = 0; j < P; ++j) st.g[j] = tot[1 + j];
    #pragma unroll 1
    for (int j = 0; j < NA; ++j) st.A[j] = tot[1 + P + j];
    #pragma unroll 1
    for (int j = 0; j < P; ++j) st.scale[j] = 1.0 / (1.0 + sqrt(st.A[tri(P, j, j)]));
    st.grad_max = grad_max_norm<P>(st, st.x, st.g);
    st.step_ok = 1;
    return lm_next_step<P>(st);
  }
  if (st.phase == 1) {
    LsSample& cur = st.cur;
    cur.value = tot[0];
    cur.value_valid = isfinite(cur.value) ? 1 : 0;
    double gr = 0.0;
    #pragma unroll 1
    for (int j = 0; j < P; ++j) gr += st.delta[j] * tot[1 + j];
    cur.gradient = gr;
    cur.gradient_valid = (cur.value_valid && isfinite(gr)) ? 1 : 0;
    const bool armijo_ok = cur.value_valid && !(cur.value > st.cost + 1e-4 * st.gd * cur.x);
    if (armijo_ok) {
      #pragma unroll 1
      for (int j = 0; j < P; ++j) st.delta[j] *= cur.x;
      return lm_after_candidate<P>(st, tot);      // candidate == this sample, bit for bit
    }
    ++st.ls_iter;
    ++st.ls_steps;
    bool fail = st.ls_iter >= 20;
    double a = 0.0;
    if (!fail) {
      a = interpolating_min_step(st.lower, st.prev, cur, 1e-3 * cur.x, 0.6 * cur.x, st.scratch);
      if (a * st.dmax < 1e-9) fail = true;
    }
    if (fail) {
      // line search failed: the full step is the candidate (delta untouched)
      project_plus<P>(st, st.x, st.delta, 1.0, st.xt);
      st.phase = 2;
      return LM_EVAL;
    }
    st.prev = cur;
    cur.x = a;
    {
      double sd[P];
      #pragma unroll 1
      for (int j = 0; j < P; ++j) sd[j] = a * st.delta[j];
      project_plus<P>(st, st.x, sd, 1.0, st.xt);
    }
    return LM_EVAL;
  }
  return lm_after_candidate<P>(st, tot);
}

// Returns LM_EVAL (st.xt set) or LM_DONE (infeasible start).
template <int P>
__device__ __noinline__ int lm_begin(LMState<P>& st, const double* init4, const double* lb3, const double* ub3, int max_iter) {
  constexpr int toff = P - 3;
  #pragma unroll 1
  for (int j = 0; j < P; ++j) { st.lb[j] = -DBL_MAX; st.ub[j] = DBL_MAX; st.x[j] = 0.0; }
  #pragma unroll 1
  for (int k = 0; k < 3; ++k) { st.lb[toff + k] = lb3[k]; st.ub[toff + k] = ub3[k]; st.x[toff + k] = init4[1 + k]; }
  st.x[P == 4 ? 0 : 1] = init4[0];        // registration.cpp:34-50
  st.max_iter = max_iter;
  st.iteration = 0; st.evals = 0; st.ls_steps = 0; st.term = -1; st.invalid = 0;
  st.radius = 1e4; st.dec = 2.0; st.reuse_diag = 0; st.step_ok = 1; st.phase = 0;
  st.cost = 0.0; st.grad_max = 0.0;
  #pragma unroll 1
  for (int j = 0; j < P; ++j) st.g[j] = 0.0;
  #pragma unroll 1
  for (int j = 0; j < P; ++j)
    if (st.x[j] < st.lb[j] || st.x[j] > st.ub[j]) { st.term = 6; return LM_DONE; }
  double xn = 0.0;
  #pragma unroll 1
  for (int j = 0; j < P; ++j) { st.xt[j] = st.x[j]; xn += st.x[j] * st.x[j]; }
  st.x_norm = sqrt(xn);
  return LM_EVAL;
}

__device__ void make_cam(const double* K9, double H, double W, Cam* cam) {
  cam->fx = K9[0]; cam->fy = K9[4]; cam->cx = K9[2]; cam->cy = K9[5];   // registration.cpp:79-82
  cam->W1 = W - 1.0; cam->H1 = H - 1.0;                                 // registration.cpp:21-22
  cam->hW = cam->W1 * 0.5; cam->hH = cam->H1 * 0.5;
}

struct SolveArgs {
  const void* xyz;
  const int8_t* label;
  const int32_t* n_pts;
  int n_stride;
  const double* K9;
  const double* init;
  double lb[3], ub[3];
  double H, W;
  int max_iter, S, I;
  double* params_all;   // [S*I*6]
  double* cost_all;     // [S*I]
  int32_t* stats_all;   // [S*I*4]
  unsigned int* queue;  // problem counter
  const float* boxes;   // [S][rounds_max][8][kThreads] box table
  const void* packed;   // [S][rounds_max * kThreads * 32] Entry<CT> (DIB_PACKED) or NULL
  int rounds_max;
  const int32_t* perm;  // [S][I] inits of each sample, longest-predicted first
  int chunk;            // samples per scheduling chunk
};

// Scheduling order.  Solve length correlates with how far an init's heading is from the centre of its
// sample's inits (rank correlation ~0.6 with the number of evaluations), so each sample's inits are
// ranked by that distance, longest-predicted first, and the queue walks chunks of samples rank-major:
// the expensive solves start early and the end-of-kernel tail (idle SMs waiting for the last
// long solves) shrinks, while concurrently running problems still share an L2-sized set of clouds.
// perm [S][I]: perm[s][r] = init index of rank r.  Results do not depend on the order.
__global__ void frustum_order_kernel(const double* __restrict__ init, int I, int32_t* __restrict__ perm) {
  extern __shared__ double key[];
  const int s = blockIdx.x;
  const double* in = init + (size_t)s * I * 4;
  double mean = 0.0;
  for (int i = 0; i < I; ++i) mean += in[(size_t)i * 4];       // every thread: same fixed-order sum
  mean /= (double)I;
  for (int i = threadIdx.x; i < I; i += blockDim.x) key[i] = fabs(in[(size_t)i * 4] - mean);
  __syncthreads();
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    const double k = key[i];
    int rank = 0;
    for (int j = 0; j < I; ++j) {
      const double kj = key[j];
      rank += (kj > k || (kj == k && j < i)) ? 1 : 0;       // NaN keys compare false: ties by index
    }
    perm[(size_t)s * I + rank] = i;
  }
}

template <typename CT, int P>
__global__ void __launch_bounds__(kThreads, (P == 4) ? DIB_MINBLOCKS4 : DIB_MINBLOCKS6) frustum_solve_kernel(SolveArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<CT, P>& sm = *reinterpret_cast<Smem<CT, P>*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&sm.full, 1);
    for (int w = 0; w < kWarps; ++w) { mbar_init(&sm.gbar[w][0], 1); mbar_init(&sm.gbar[w][1], 1); }
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t box_phase = 0, gphase = 0;      // gphase: bit s = parity of this warp's staging slot s
  const int total = a.S * a.I;

  for (;;) {
    if (tid == 0) {
      const int q = (int)atomicAdd(a.queue, 1u);
      int prob = q;
      if (q < total) {
        const int per_chunk = a.chunk * a.I;
        const int c = q / per_chunk, within = q - c * per_chunk;
        const int s0 = c * a.chunk;
        const int gc = min(a.chunk, a.S - s0);             // samples in this chunk
        const int r = within / gc, sc = s0 + (within - r * gc);
        prob = sc * a.I + a.perm[(size_t)sc * a.I + r];
      }
      sm.problem = prob;
    }
    __syncthreads();
    const int prob = sm.problem;
    if (prob >= total) break;
    const int s = prob / a.I;
    const CT* xyz_s = reinterpret_cast<const CT*>(a.xyz) + (size_t)s * 3 * a.n_stride;
    const int8_t* lab_s = a.label + (size_t)s * a.n_stride;
    const Entry<CT>* pk_s = reinterpret_cast<const Entry<CT>*>(a.packed) + (size_t)s * a.rounds_max * (kThreads * 32);
    const float* box_s = a.boxes + (size_t)s * a.rounds_max * kBoxRoundFloats;
    const int n = a.n_pts ? a.n_pts[s] : a.n_stride;
    if (tid == 0) {
      make_cam(a.K9 + (size_t)s * 9, a.H, a.W, &sm.cam);
      const int rc = lm_begin<P>(sm.lm, a.init + (size_t)prob * 4, a.lb, a.ub, a.max_iter);
      sm.go = rc;
      if (rc == LM_EVAL) { make_pose<P>(sm.lm.xt, &sm.pose); make_class(sm.pose, sm.cam, &sm.cls); }
    }
    const bool resident = load_boxes<CT, P>(sm, box_s, n, box_phase);   // overlaps thread 0's set-up
    __syncthreads();
    while (sm.go == LM_EVAL) {
      evaluate_cloud<CT, P>(sm, xyz_s, lab_s, pk_s, a.n_stride, n, box_s, resident, box_phase, gphase);
      if (tid == 0) {
        const int rc = lm_consume<P>(sm.lm, sm.tot);
        sm.go = rc;
        if (rc == LM_EVAL) { make_pose<P>(sm.lm.xt, &sm.pose); make_class(sm.pose, sm.cam, &sm.cls); }
      }
      __syncthreads();
    }
    if (tid == 0) {
      const LMState<P>& st = sm.lm;
      double* po = a.params_all + (size_t)prob * 6;
      for (int j = 0; j < 6; ++j) po[j] = (j < P) ? st.x[j] : 0.0;
      a.cost_all[prob] = st.cost;
      int32_t* so = a.stats_all + (size_t)prob * 4;
      so[0] = st.iteration; so[1] = st.evals; so[2] = st.ls_steps; so[3] = st.term;
    }
    __syncthreads();
  }
}

// Pose matrix from parameters (registration.cpp:161-185).
__device__ void pose_matrix(const double* x, int P, double* M) {
  double aa[3];
  const double* t;
  if (P == 4) { aa[0] = 0; aa[1] = x[0]; aa[2] = 0; t = x + 1; } else { aa[0] = x[0]; aa[1] = x[1]; aa[2] = x[2]; t = x + 3; }
  double R[9];
  const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (th2 > DBL_EPSILON) {
    const double th = sqrt(th2);
    const double wx = aa[0] / th, wy = aa[1] / th, wz = aa[2] / th;
    double s, c;
    sincos(th, &s, &c);
    const double omc = 1.0 - c;
    R[0] = c + wx * wx * omc;       R[1] = wx * wy * omc - wz * s;  R[2] = wy * s + wx * wz * omc;
    R[3] = wz * s + wx * wy * omc;  R[4] = c + wy * wy * omc;       R[5] = -wx * s + wy * wz * omc;
    R[6] = -wy * s + wx * wz * omc; R[7] = wx * s + wy * wz * omc;  R[8] = c + wz * wz * omc;
  } else {
    R[0] = 1; R[1] = -aa[2]; R[2] = aa[1];
    R[3] = aa[2]; R[4] = 1; R[5] = -aa[0];
    R[6] = -aa[1]; R[7] = aa[0]; R[8] = 1;
  }
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) M[4 * i + j] = R[3 * i + j]; M[4 * i + 3] = t[i]; }
  M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
}

// Per-sample arg-min over inits (lowest index wins ties; NaN never wins) + pose matrix.
__global__ void frustum_finalize_kernel(const double* params_all, const double* cost_all, int S, int I, int P,
                                        double* P16_out, double* cost_out, int32_t* best_out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  int best = 0;
  double bc = cost_all[(size_t)s * I];
  if (!(bc == bc)) bc = INFINITY;
  for (int i = 1; i < I; ++i) {
    const double c = cost_all[(size_t)s * I + i];
    if (c < bc) { bc = c; best = i; }
  }
  pose_matrix(params_all + ((size_t)s * I + best) * 6, P, P16_out + (size_t)s * 16);
  cost_out[s] = cost_all[(size_t)s * I + best];
  if (best_out) best_out[s] = best;
}

template <typename CT, int P>
__global__ void __launch_bounds__(kThreads) frustum_evaluate_kernel(const CT* xyz, const int8_t* label,
                                                                     const int32_t* n_pts, int n_stride,
                                                                     const double* K9, const double* x, double H,
                                                                     double W, const float* boxes,
                                                                     const Entry<CT>* packed, int rounds_max,
                                                                     double* cost_out, double* grad_out,
                                                                     double* JtJ_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<CT, P>& sm = *reinterpret_cast<Smem<CT, P>*>(smem_raw);
  const int tid = threadIdx.x;
  const int s = blockIdx.x;
  if (tid == 0) {
    mbar_init(&sm.full, 1);
    for (int w = 0; w < kWarps; ++w) { mbar_init(&sm.gbar[w][0], 1); mbar_init(&sm.gbar[w][1], 1); }
    mbar_fence_init();
    make_cam(K9 + (size_t)s * 9, H, W, &sm.cam);
    make_pose<P>(x + (size_t)s * 6, &sm.pose);
    make_class(sm.pose, sm.cam, &sm.cls);
  }
  __syncthreads();
  uint32_t box_phase = 0, gphase = 0;
  const int n = n_pts ? n_pts[s] : n_stride;
  const float* box_s = boxes + (size_t)s * rounds_max * kBoxRoundFloats;
  const bool resident = load_boxes<CT, P>(sm, box_s, n, box_phase);
  __syncthreads();
  evaluate_cloud<CT, P>(sm, xyz + (size_t)s * 3 * n_stride, label + (size_t)s * n_stride,
                        packed + (size_t)s * rounds_max * (kThreads * 32), n_stride, n, box_s, resident, box_phase, gphase);
  if (tid == 0) {
    cost_out[s] = sm.tot[0];
    for (int j = 0; j < 6; ++j) grad_out[(size_t)s * 6 + j] = (j < P) ? sm.tot[1 + j] : 0.0;
    for (int j = 0; j < 36; ++j) JtJ_out[(size_t)s * 36 + j] = 0.0;
    for (int j = 0; j < P; ++j)
      for (int k = 0; k < P; ++k)
        JtJ_out[(size_t)s * 36 + j * P + k] = sm.tot[1 + P + (j <= k ? tri(P, j, k) : tri(P, k, j))];
  }
}

template <typename CT, int P>
__global__ void frustum_residuals_kernel(const CT* xyz, const int8_t* label, int n, int n_stride, const double* K9,
                                         const double* x, double H, double W, const int32_t* row_offset,
                                         double* residuals) {
  __shared__ PoseConst pc;
  __shared__ Cam cam;
  if (threadIdx.x == 0) { make_cam(K9, H, W, &cam); make_pose<P>(x, &pc); }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int lab = label[i];
  if (lab != 0 && lab != 1) return;
  double r[3] = {0, 0, 0}, J[3][P];
  int nrows = (lab == 1) ? 3 : 1;
  const bool any = point_rows<P>((double)xyz[i], (double)xyz[n_stride + i], (double)xyz[2 * (size_t)n_stride + i], lab,
                                 cam, pc, r, J, &nrows);
  double s = 0.0;
  for (int k = 0; k < nrows; ++k) s += r[k] * r[k];
  const double sq = sqrt(1.0 / (1.0 + s));
  const int ro = row_offset[i];
  for (int k = 0; k < nrows; ++k) residuals[ro + k] = any ? r[k] * sq : 0.0;
}

// ------------------------------------------------------------------------------------------
// Host side of the C ABI.
// ------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static size_t box_table_bytes(int S, int n_stride) {
  return align_up((size_t)(S > 0 ? S : 0) * box_rounds(n_stride) * kBoxRoundFloats * sizeof(float), 256);
}
// packed per-launch copy of the clouds; sized for the wider (f64) record so one workspace serves both ABIs
static size_t packed_bytes(int S, int n_stride) {
#if DIB_PACKED
  return align_up((size_t)(S > 0 ? S : 0) * box_rounds(n_stride) * (kThreads * 32) * sizeof(Entry<double>), 256);
#else
  (void)S; (void)n_stride;
  return 0;
#endif
}

template <typename CT>
static int launch_boxes(const CT* xyz, const int8_t* label, const int32_t* n_pts, int n_stride, int S, float* table,
                        Entry<CT>* packed, cudaStream_t st) {
  const int rounds_max = box_rounds(n_stride);
  if (rounds_max == 0 || S == 0) return DIB_OK;
  DIB_REQUIRE(S <= 65535, "S (%d) exceeds grid.y; split the batch", S);
  const int groups = rounds_max * kThreads;
  dim3 grid((groups + 7) / 8, S);
  frustum_boxes_kernel<CT><<<grid, 256, 0, st>>>(xyz, label, n_pts, n_stride, rounds_max, table, packed);
  DIB_CHECK_CUDA(cudaGetLastError());
  return DIB_OK;
}

template <typename CT>
static int check_cloud_args(const CT* xyz, const int8_t* label, int n_stride, int S) {
  DIB_REQUIRE(xyz != nullptr && label != nullptr, "xyz/label must not be NULL");
  DIB_REQUIRE(S >= 0 && n_stride >= 0, "negative size");
  DIB_REQUIRE(n_stride % 16 == 0, "n_stride (%d) must be a multiple of 16", n_stride);
  DIB_REQUIRE(((uintptr_t)xyz % 16) == 0 && ((uintptr_t)label % 16) == 0, "xyz/label must be 16-byte aligned");
  return DIB_OK;
}

template <typename CT, int P>
static int launch_solve(const SolveArgs& a_in, cudaStream_t st) {
  SolveArgs a = a_in;
  auto kern = frustum_solve_kernel<CT, P>;
  const size_t smem = sizeof(Smem<CT, P>);
  DIB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0, per_sm = 0;
  DIB_CHECK_CUDA(cudaGetDevice(&dev));
  DIB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DIB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)sms * per_sm;
  const long long total = (long long)a.S * a.I;
  if (grid > total) grid = total;
  if (grid < 1) return DIB_OK;
  // scheduling chunk: as many samples as keep the concurrently touched clouds around 32 MB (a quarter of
  // L2), and at least ~2x the resident problems.  Measured on B200 (512 x 60 problems): 1 or 50 samples
  // 103.3 ms, 128 samples 97.7 ms, 256 samples 99.8 ms, 512 samples (no chunking) ~100 ms.
  long long chunk = (2 * grid + a.I - 1) / a.I;
  const long long bytes_per_sample =
      DIB_PACKED ? (long long)a.rounds_max * (kThreads * 32) * (long long)sizeof(Entry<CT>)
                 : (long long)a.n_stride * (3 * (long long)sizeof(CT) + 1);
  if (bytes_per_sample > 0 && chunk < (32ll << 20) / bytes_per_sample) chunk = (32ll << 20) / bytes_per_sample;
  if (const char* e = getenv("DIB_CHUNK_SAMPLES")) chunk = atoll(e);   // tuning knob
  if (chunk < 1) chunk = 1;
  if (chunk > a.S) chunk = a.S;
  {
    // equal-sized chunks: a short last chunk would bring the tail back
    const long long nchunks = (a.S + chunk - 1) / chunk;
    chunk = (a.S + nchunks - 1) / nchunks;
  }
  a.chunk = (int)chunk;
  // (Queueing the top 4 / 10 longest-predicted inits of EVERY sample ahead of the chunked walk measured
  // 1 % / 3 % slower: the tail comes from mispredicted long solves, not from the predicted ones.)
  kern<<<(unsigned)grid, kThreads, smem, st>>>(a);
  DIB_CHECK_CUDA(cudaGetLastError());
  return DIB_OK;
}

template <typename CT>
static int solve_batch(const CT* xyz, const int8_t* label, const int32_t* n_pts, int n_stride, const double* K9,
                       const double* init, const double* lb3, const double* ub3, double H, double W, int max_iter,
                       int is_2d, int S, int I, double* P16_out, double* cost_out, int32_t* best_out,
                       double* params_all, double* cost_all, int32_t* stats_all, void* workspace,
                       size_t workspace_bytes, dib_stream_t stream) {
  int rc = check_cloud_args<CT>(xyz, label, n_stride, S);
  if (rc != DIB_OK) return rc;
  DIB_REQUIRE(I >= 1, "I must be >= 1");
  DIB_REQUIRE(K9 && init && lb3 && ub3 && P16_out && cost_out, "NULL argument");
  DIB_REQUIRE((long long)S * I < (1ll << 31), "S*I too large");
  if (S == 0) return DIB_OK;
  if (workspace_bytes < frustum_solve_workspace_bytes(S, I, n_stride) || workspace == nullptr) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, frustum_solve_workspace_bytes(S, I, n_stride));
    return DIB_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)workspace;
  const size_t n = (size_t)S * I;
  SolveArgs a;
  a.queue = (unsigned int*)ws;
  size_t off = 256;
  a.params_all = params_all ? params_all : (double*)(ws + off);
  off += align_up(n * 6 * sizeof(double), 256);
  a.cost_all = cost_all ? cost_all : (double*)(ws + off);
  off += align_up(n * sizeof(double), 256);
  a.stats_all = stats_all ? stats_all : (int32_t*)(ws + off);
  off += align_up(n * 4 * sizeof(int32_t), 256);
  float* table = (float*)(ws + off);
  off += box_table_bytes(S, n_stride);
  Entry<CT>* packed = DIB_PACKED ? (Entry<CT>*)(ws + off) : nullptr;
  off += packed_bytes(S, n_stride);
  int32_t* perm = (int32_t*)(ws + off);
  a.boxes = table;
  a.packed = packed;
  a.rounds_max = box_rounds(n_stride);
  a.perm = perm;
  a.xyz = xyz; a.label = label; a.n_pts = n_pts; a.n_stride = n_stride; a.K9 = K9; a.init = init;
  for (int k = 0; k < 3; ++k) { a.lb[k] = lb3[k]; a.ub[k] = ub3[k]; }
  a.H = H; a.W = W; a.max_iter = max_iter; a.S = S; a.I = I;
  DIB_CHECK_CUDA(cudaMemsetAsync(a.queue, 0, 256, st));
  rc = launch_boxes<CT>(xyz, label, n_pts, n_stride, S, table, packed, st);
  if (rc != DIB_OK) return rc;
  {
    const int threads = I < 256 ? ((I + 31) / 32) * 32 : 256;
    DIB_REQUIRE((size_t)I * sizeof(double) <= 48 * 1024, "I=%d too large for the ordering kernel", I);
    frustum_order_kernel<<<S, threads, (size_t)I * sizeof(double), st>>>(init, I, perm);
    DIB_CHECK_CUDA(cudaGetLastError());
  }
  rc = is_2d ? launch_solve<CT, 4>(a, st) : launch_solve<CT, 6>(a, st);
  if (rc != DIB_OK) return rc;
  frustum_finalize_kernel<<<(S + 127) / 128, 128, 0, st>>>(a.params_all, a.cost_all, S, I, is_2d ? 4 : 6, P16_out,
                                                          cost_out, best_out);
  DIB_CHECK_CUDA(cudaGetLastError());
  return DIB_OK;
}

template <typename CT>
static int evaluate_batch(const CT* xyz, const int8_t* label, const int32_t* n_pts, int n_stride, const double* K9,
                          const double* x, double H, double W, int is_2d, int S, double* cost_out, double* grad_out,
                          double* JtJ_out, void* workspace, size_t workspace_bytes, dib_stream_t stream) {
  int rc = check_cloud_args<CT>(xyz, label, n_stride, S);
  if (rc != DIB_OK) return rc;
  DIB_REQUIRE(K9 && x && cost_out && grad_out && JtJ_out, "NULL argument");
  if (S == 0) return DIB_OK;
  if (workspace_bytes < frustum_evaluate_workspace_bytes(S, n_stride) || workspace == nullptr) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, frustum_evaluate_workspace_bytes(S, n_stride));
    return DIB_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* table = (float*)workspace;
  Entry<CT>* packed = DIB_PACKED ? (Entry<CT>*)((unsigned char*)workspace + box_table_bytes(S, n_stride)) : nullptr;
  const int rounds_max = box_rounds(n_stride);
  rc = launch_boxes<CT>(xyz, label, n_pts, n_stride, S, table, packed, st);
  if (rc != DIB_OK) return rc;
  if (is_2d) {
    auto kern = frustum_evaluate_kernel<CT, 4>;
    const size_t smem = sizeof(Smem<CT, 4>);
    DIB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<S, kThreads, smem, st>>>(xyz, label, n_pts, n_stride, K9, x, H, W, table, packed, rounds_max, cost_out,
                                    grad_out, JtJ_out);
  } else {
    auto kern = frustum_evaluate_kernel<CT, 6>;
    const size_t smem = sizeof(Smem<CT, 6>);
    DIB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<S, kThreads, smem, st>>>(xyz, label, n_pts, n_stride, K9, x, H, W, table, packed, rounds_max, cost_out,
                                    grad_out, JtJ_out);
  }
  DIB_CHECK_CUDA(cudaGetLastError());
  return DIB_OK;
}

template <typename CT>
static int residuals_single(const CT* xyz, const int8_t* label, int n, int n_stride, const double* K9,
                            const double* x, double H, double W, int is_2d, const int32_t* row_offset,
                            double* residuals, dib_stream_t stream) {
  DIB_REQUIRE(xyz && label && K9 && x && row_offset && residuals, "NULL argument");
  DIB_REQUIRE(n >= 0 && n <= n_stride, "bad n");
  if (n == 0) return DIB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (n + 255) / 256;
  if (is_2d)
    frustum_residuals_kernel<CT, 4><<<blocks, 256, 0, st>>>(xyz, label, n, n_stride, K9, x, H, W, row_offset, residuals);
  else
    frustum_residuals_kernel<CT, 6><<<blocks, 256, 0, st>>>(xyz, label, n, n_stride, K9, x, H, W, row_offset, residuals);
  DIB_CHECK_CUDA(cudaGetLastError());
  return DIB_OK;
}

static size_t solve_workspace_bytes_impl(int S, int I, int n_stride) {
  const size_t n = (size_t)(S > 0 ? S : 0) * (size_t)(I > 0 ? I : 0);
  return 256 + align_up(n * 6 * sizeof(double), 256) + align_up(n * sizeof(double), 256) +
         align_up(n * 4 * sizeof(int32_t), 256) + box_table_bytes(S, n_stride > 0 ? n_stride : 0) +
         packed_bytes(S, n_stride > 0 ? n_stride : 0) + align_up(n * sizeof(int32_t), 256);
}

#if DIB_WIDE_TU
// Entry points of the wide (128-thread CTA) build of this file; called by the primary TU's C ABI for
// batches with few problems (a problem then finishes ~1.8x sooner and the persistent grid balances better).
size_t wide_solve_workspace_bytes(int S, int I, int n_stride) { return solve_workspace_bytes_impl(S, I, n_stride); }
int wide_solve_f32(const float* xyz, const int8_t* label, const int32_t* n_pts, int n_stride, const double* K9,
                   const double* init, const double* lb3, const double* ub3, double H, double W, int max_iter, int is_2d,
                   int S, int I, double* P16_out, double* cost_out, int32_t* best_out, double* params_all,
                   double* cost_all, int32_t* stats_all, void* workspace, size_t workspace_bytes, dib_stream_t stream) {
  return solve_batch<float>(xyz, label, n_pts, n_stride, K9, init, lb3, ub3, H, W, max_iter, is_2d, S, I, P16_out,
                            cost_out, best_out, params_all, cost_all, stats_all, workspace, workspace_bytes, stream);
}
int wide_solve_f64(const double* xyz, const int8_t* label, const int32_t* n_pts, int n_stride, const double* K9,
                   const double* init, const double* lb3, const double* ub3, double H, double W, int max_iter, int is_2d,
                   int S, int I, double* P16_out, double* cost_out, int32_t* best_out, double* params_all,
                   double* cost_all, int32_t* stats_all, void* workspace, size_t workspace_bytes, dib_stream_t stream) {
  return solve_batch<double>(xyz, label, n_pts, n_stride, K9, init, lb3, ub3, H, W, max_iter, is_2d, S, I, P16_out,
                             cost_out, best_out, params_all, cost_all, stats_all, workspace, workspace_bytes, stream);
}
#endif

}  // namespace dib

#if !DIB_WIDE_TU
#ifndef DIB_HAVE_WIDE
#define DIB_HAVE_WIDE 0                       // set by the build when frustum_solver_wide.cu is linked in
#endif
#if DIB_HAVE_WIDE
namespace dib_w128 {
size_t wide_solve_workspace_bytes(int S, int I, int n_stride);
int wide_solve_f32(const float*, const int8_t*, const int32_t*, int, const double*, const double*, const double*,
                   const double*, double, double, int, int, int, int, double*, double*, int32_t*, double*, double*,
                   int32_t*, void*, size_t, dib_stream_t);
int wide_solve_f64(const double*, const int8_t*, const int32_t*, int, const double*, const double*, const double*,
                   const double*, double, double, int, int, int, int, double*, double*, int32_t*, double*, double*,
                   int32_t*, void*, size_t, dib_stream_t);
}  // namespace dib_w128
#endif
// Batches with fewer than DIB_WIDE_BELOW problems (S x I) go to the 128-thread build.  Experimental: the
// default 0 never does (not yet measured on a B200; scripts/round2_sweep.sh).
static bool use_wide(int S, int I) {
#if DIB_HAVE_WIDE
  static const long long below = getenv("DIB_WIDE_BELOW") ? atoll(getenv("DIB_WIDE_BELOW")) : 0;
  return (long long)S * I < below;
#else
  (void)S; (void)I;
  return false;
#endif
}

extern "C" {

int dib_abi_version(void) { return 2; }
const char* dib_last_error(void) { return dib::g_err; }

int dib_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DIB_ENODEV;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return DIB_ENODEV;
  return sms;
}

size_t frustum_solve_workspace_bytes(int S, int I, int n_stride) {
  size_t need = dib::solve_workspace_bytes_impl(S, I, n_stride);
#if DIB_HAVE_WIDE
  const size_t wide = dib_w128::wide_solve_workspace_bytes(S, I, n_stride);
  if (wide > need) need = wide;
#endif
  return need;
}

size_t frustum_evaluate_workspace_bytes(int S, int n_stride) {
  return dib::box_table_bytes(S, n_stride > 0 ? n_stride : 0) + dib::packed_bytes(S, n_stride > 0 ? n_stride : 0) + 256;
}

int frustum_solve_batch_f32(const float* xyz, const int8_t* label, const int32_t* n_pts, int n_stride,
                            const double* K9, const double* init, const double* lb3, const double* ub3, double H,
                            double W, int max_iter, int is_2d, int S, int I, double* P16_out, double* cost_out,
                            int32_t* best_out, double* params_all, double* cost_all, int32_t* stats_all,
                            void* workspace, size_t workspace_bytes, dib_stream_t stream) {
#if DIB_HAVE_WIDE
  if (use_wide(S, I))
    return dib_w128::wide_solve_f32(xyz, label, n_pts, n_stride, K9, init, lb3, ub3, H, W, max_iter, is_2d, S, I, P16_out,
                                    cost_out, best_out, params_all, cost_all, stats_all, workspace, workspace_bytes, stream);
#endif
  return dib::solve_batch<float>(xyz, label, n_pts, n_stride, K9, init, lb3, ub3, H, W, max_iter, is_2d, S, I,
                                 P16_out, cost_out, best_out, params_all, cost_all, stats_all, workspace,
                                 workspace_bytes, stream);
}

int frustum_solve_batch_f64(const double* xyz, const int8_t* label, const int32_t* n_pts, int n_stride,
                            const double* K9, const double* init, const double* lb3, const double* ub3, double H,
                            double W, int max_iter, int is_2d, int S, int I, double* P16_out, double* cost_out,
                            int32_t* best_out, double* params_all, double* cost_all, int32_t* stats_all,
                            void* workspace, size_t workspace_bytes, dib_stream_t stream) {
#if DIB_HAVE_WIDE
  if (use_wide(S, I))
    return dib_w128::wide_solve_f64(xyz, label, n_pts, n_stride, K9, init, lb3, ub3, H, W, max_iter, is_2d, S, I, P16_out,
                                    cost_out, best_out, params_all, cost_all, stats_all, workspace, workspace_bytes, stream);
#endif
  return dib::solve_batch<double>(xyz, label, n_pts, n_stride, K9, init, lb3, ub3, H, W, max_iter, is_2d, S, I,
                                  P16_out, cost_out, best_out, params_all, cost_all, stats_all, workspace,
                                  workspace_bytes, stream);
}

int frustum_evaluate_f32(const float* xyz, const int8_t* label, const int32_t* n_pts, int n_stride, const double* K9,
                         const double* x, double H, double W, int is_2d, int S, double* cost_out, double* grad_out,
                         double* JtJ_out, void* workspace, size_t workspace_bytes, dib_stream_t stream) {
  return dib::evaluate_batch<float>(xyz, label, n_pts, n_stride, K9, x, H, W, is_2d, S, cost_out, grad_out, JtJ_out,
                                    workspace, workspace_bytes, stream);
}
int frustum_evaluate_f64(const double* xyz, const int8_t* label, const int32_t* n_pts, int n_stride,
                         const double* K9, const double* x, double H, double W, int is_2d, int S, double* cost_out,
                         double* grad_out, double* JtJ_out, void* workspace, size_t workspace_bytes,
                         dib_stream_t stream) {
  return dib::evaluate_batch<double>(xyz, label, n_pts, n_stride, K9, x, H, W, is_2d, S, cost_out, grad_out, JtJ_out,
                                     workspace, workspace_bytes, stream);
}

int frustum_residuals_f32(const float* xyz, const int8_t* label, int n, int n_stride, const double* K9,
                          const double* x, double H, double W, int is_2d, const int32_t* row_offset,
                          double* residuals, dib_stream_t stream) {
  return dib::residuals_single<float>(xyz, label, n, n_stride, K9, x, H, W, is_2d, row_offset, residuals, stream);
}
int frustum_residuals_f64(const double* xyz, const int8_t* label, int n, int n_stride, const double* K9,
                          const double* x, double H, double W, int is_2d, const int32_t* row_offset,
                          double* residuals, dib_stream_t stream) {
  return dib::residuals_single<double>(xyz, label, n, n_stride, K9, x, H, W, is_2d, row_offset, residuals, stream);
}

}  // extern "C"
#endif  // !DIB_WIDE_TU
