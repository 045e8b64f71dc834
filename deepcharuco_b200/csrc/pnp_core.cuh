// (host + device core of pnp.cu; tests/pnp_host_check.cu runs the same code on the CPU against cv2.solvePnP)
// Batched board pose from the engine's corner lists: the step that follows the hot path in the reference's callers
// (pose_estimation.py:61-63 -> inference.solve_pnp, inference.py:15-29 -> cv2.solvePnP, SOLVEPNP_ITERATIVE).
//
// cv2.solvePnP lives in OpenCV (third-party, not under /root/reference; reference pin opencv >=4.6,<4.12, 4.13 in this
// image).  Its published algorithm for coplanar object points (calib3d, cvFindExtrinsicCameraParams2) is restated here in fp64:
//   1. image points -> normalised camera coordinates (inverse of the distortion model by fixed-point iteration);
//   2. board-plane -> image homography by the normalised DLT (9x9 symmetric eigenproblem, cyclic Jacobi);
//   3. pose from the homography: r1, r2 = normalised columns, r3 = r1 x r2, t = h3 * 2 / (|h1| + |h2|); nearest rotation;
//   4. Levenberg-Marquardt on the pixel reprojection error with the full distortion model and analytic Jacobian, with
//      CvLevMarq's schedule: lambda = 10^-3, x10 on a worse step (<= 10^16), /10 on a better one, stop after 20 accepted
//      steps or when |delta| / |param| < FLT_EPSILON.
//   (2b: like cv::findHomography, the DLT estimate is refined on the transfer error before step 3 -- to convergence here, 10 LM
//    iterations in OpenCV; without it the two start step 4 from slightly different poses and frames on which its 20 steps do not
//    converge agreed only to ~1e-3.)
// Both implementations minimise the same function from the same start with the same schedule: measured agreement with cv2 is
// median 7e-15, max 8e-8 in rvec / tvec over the golden frames x 3 distortion models (tests/test_pnp_host.py).
//
// One warp per frame (lanes share the per-corner loops): the whole solve is ~1e5 flops, so a 256-frame batch is one
// small launch (tens of microseconds) next to a 15 ms detector + RefineNet step; points are re-read from global memory in
// every pass, so there is no per-frame capacity limit.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace dcu {
namespace pnp {

// How the per-corner loops of one frame are shared out.  Host (and a thread-per-frame kernel): one worker.  The CUDA kernel
// uses one warp per frame: lane l takes corners l, l+32, ... and the sums are combined by shuffles, then broadcast from lane 0
// so that every lane takes the same branches afterwards.  Everything that is not a loop over corners (Jacobi sweeps, the 6x6
// solve, Rodrigues) is computed redundantly by all lanes.
struct Serial {
  __host__ __device__ int first() const { return 0; }
  __host__ __device__ int step() const { return 1; }
  __host__ __device__ void sum(double*, int) const {}
};
#ifdef __CUDACC__
struct WarpLanes {
  __device__ int first() const { return (int)(threadIdx.x & 31u); }
  __device__ int step() const { return 32; }
  __device__ void sum(double* v, int n) const {
    for (int i = 0; i < n; ++i) {
      double x = v[i];
      for (int o = 16; o >= 1; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      v[i] = __shfl_sync(0xffffffffu, x, 0);
    }
  }
};
#endif

struct Cam { double fx, fy, cx, cy, k[8]; };    // k1 k2 p1 p2 k3 k4 k5 k6 (OpenCV order)

struct Pts {
  const int32_t* kp;      // [K][4] x, y, id, cell
  const float* xy;        // [K][2] refined or null
  const float* obj;       // [n_obj][2] board corner coordinates (float32, as the reference builds them)
  int n, n_obj;
  __host__ __device__ inline void get(int i, double& X, double& Y, double& u, double& v) const {
    const int id = kp[4 * i + 2];
    X = (double)obj[2 * id]; Y = (double)obj[2 * id + 1];
    if (xy) { u = (double)xy[2 * i]; v = (double)xy[2 * i + 1]; }
    else { u = (double)(float)kp[4 * i]; v = (double)(float)kp[4 * i + 1]; }
  }
};

__host__ __device__ inline void rodrigues(const double r[3], double R[9], double* dRdr /* [3][9] or null */) {
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (th < 2.220446049250313e-16) {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (dRdr) {
      for (int i = 0; i < 27; ++i) dRdr[i] = 0.0;
      dRdr[5] = dRdr[15] = dRdr[19] = -1.0;
      dRdr[7] = dRdr[11] = dRdr[21] = 1.0;
    }
    return;
  }
  const double c = cos(th), s = sin(th), c1 = 1.0 - c, it = 1.0 / th;
  const double k[3] = {r[0] * it, r[1] * it, r[2] * it};
  const double rrt[9] = {k[0] * k[0], k[0] * k[1], k[0] * k[2], k[0] * k[1], k[1] * k[1], k[1] * k[2], k[0] * k[2], k[1] * k[2], k[2] * k[2]};
  const double rx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
  for (int i = 0; i < 9; ++i) R[i] = c * ((i % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[i] + s * rx[i];
  if (!dRdr) return;
  const double drrt[27] = {k[0] + k[0], k[1], k[2], k[1], 0, 0, k[2], 0, 0,
                           0, k[0], 0, k[0], k[1] + k[1], k[2], 0, k[2], 0,
                           0, 0, k[0], 0, 0, k[1], k[0], k[1], k[2] + k[2]};
  const double drx[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0,
                          0, 0, 1, 0, 0, 0, -1, 0, 0,
                          0, -1, 0, 1, 0, 0, 0, 0, 0};
  for (int i = 0; i < 3; ++i) {
    const double ri = k[i];
    const double a0 = -s * ri, a1 = (s - 2 * c1 * it) * ri, a2 = c1 * it, a3 = (c - s * it) * ri, a4 = s * it;
    for (int j = 0; j < 9; ++j)
      dRdr[i * 9 + j] = a0 * ((j % 4 == 0) ? 1.0 : 0.0) + a1 * rrt[j] + a2 * drrt[i * 9 + j] + a3 * rx[j] + a4 * drx[i * 9 + j];
  }
}

// nearest rotation to the (nearly orthonormal) matrix M, then its rotation vector
__host__ __device__ inline void matrix_to_rvec(double M[9], double r[3]) {
  // polar decomposition by Newton iteration  M <- (M + M^-T) / 2  (converges quadratically from a near-rotation)
  for (int it = 0; it < 12; ++it) {
    const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
    if (fabs(det) < 1e-300) break;
    const double id = 1.0 / det;
    double T[9];   // M^-T = cofactor / det
    T[0] = c00 * id; T[1] = c01 * id; T[2] = c02 * id;
    T[3] = (M[2] * M[7] - M[1] * M[8]) * id; T[4] = (M[0] * M[8] - M[2] * M[6]) * id; T[5] = (M[1] * M[6] - M[0] * M[7]) * id;
    T[6] = (M[1] * M[5] - M[2] * M[4]) * id; T[7] = (M[2] * M[3] - M[0] * M[5]) * id; T[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    double d = 0;
    for (int i = 0; i < 9; ++i) { const double n = 0.5 * (M[i] + T[i]); d += fabs(n - M[i]); M[i] = n; }
    if (d < 1e-15) break;
  }
  double x = M[7] - M[5], y = M[2] - M[6], z = M[3] - M[1];
  const double s = sqrt((x * x + y * y + z * z) * 0.25);
  double c = (M[0] + M[4] + M[8] - 1.0) * 0.5;
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  const double th = acos(c);
  if (s < 1e-5) {
    if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
    double t = (M[0] + 1) * 0.5; x = sqrt(t > 0 ? t : 0);
    t = (M[4] + 1) * 0.5; y = sqrt(t > 0 ? t : 0) * (M[1] < 0 ? -1.0 : 1.0);
    t = (M[8] + 1) * 0.5; z = sqrt(t > 0 ? t : 0) * (M[2] < 0 ? -1.0 : 1.0);
    if (fabs(x) < fabs(y) && fabs(x) < fabs(z) && ((M[5] > 0) != (y * z > 0))) z = -z;
    const double n = th / sqrt(x * x + y * y + z * z);
    r[0] = x * n; r[1] = y * n; r[2] = z * n;
    return;
  }
  const double f = th / (2 * s);
  r[0] = x * f; r[1] = y * f; r[2] = z * f;
}

// cyclic Jacobi on a symmetric 9x9: eigenvector of the smallest eigenvalue -> v
__host__ __device__ inline void smallest_eigvec9(double A[81], double v[9]) {
  double V[81];
  for (int i = 0; i < 81; ++i) V[i] = (i % 10 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < 9; ++i)
      for (int j = 0; j < 9; ++j) (i == j ? diag : off) += A[i * 9 + j] * A[i * 9 + j];
    if (off <= 1e-30 * diag) break;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) {
        const double apq = A[p * 9 + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * 9 + q] - A[p * 9 + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1.0 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 9; ++k) {
          const double akp = A[k * 9 + p], akq = A[k * 9 + q];
          A[k * 9 + p] = c * akp - s * akq; A[k * 9 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 9; ++k) {
          const double apk = A[p * 9 + k], aqk = A[q * 9 + k];
          A[p * 9 + k] = c * apk - s * aqk; A[q * 9 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 9; ++k) {
          const double vkp = V[k * 9 + p], vkq = V[k * 9 + q];
          V[k * 9 + p] = c * vkp - s * vkq; V[k * 9 + q] = s * vkp + c * vkq;
        }
      }
  }
  int m = 0;
  for (int i = 1; i < 9; ++i)
    if (A[i * 9 + i] < A[m * 9 + m]) m = i;
  for (int k = 0; k < 9; ++k) v[k] = V[k * 9 + m];
}

__host__ __device__ inline void undistort(const Cam& c, double u, double v, double& x, double& y) {
  const double x0 = (u - c.cx) / c.fx, y0 = (v - c.cy) / c.fy;
  x = x0; y = y0;
  bool any = false;
  for (int i = 0; i < 8; ++i) any |= (c.k[i] != 0.0);
  if (!any) return;
  for (int it = 0; it < 20; ++it) {
    const double r2 = x * x + y * y;
    const double icd = (1 + ((c.k[7] * r2 + c.k[6]) * r2 + c.k[5]) * r2) / (1 + ((c.k[4] * r2 + c.k[1]) * r2 + c.k[0]) * r2);
    const double dx = 2 * c.k[2] * x * y + c.k[3] * (r2 + 2 * x * x), dy = c.k[2] * (r2 + 2 * y * y) + 2 * c.k[3] * x * y;
    x = (x0 - dx) * icd; y = (y0 - dy) * icd;
  }
}

// pixel projection of board point (X, Y, 0); J (optional): d(u,v)/d(r0 r1 r2 t0 t1 t2) as [2][6]
__host__ __device__ inline void project(const Cam& c, const double R[9], const double* dRdr, const double t[3], double X, double Y, double& u,
                        double& v, double* J) {
  const double Px = R[0] * X + R[1] * Y + t[0], Py = R[3] * X + R[4] * Y + t[1], Pz = R[6] * X + R[7] * Y + t[2];
  const double z = Pz != 0.0 ? 1.0 / Pz : 1.0;
  const double x = Px * z, y = Py * z;
  const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
  const double cd = 1 + c.k[0] * r2 + c.k[1] * r4 + c.k[4] * r6;
  const double icd2 = 1.0 / (1 + c.k[5] * r2 + c.k[6] * r4 + c.k[7] * r6);
  const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
  u = c.fx * (x * cd * icd2 + c.k[2] * a1 + c.k[3] * a2) + c.cx;
  v = c.fy * (y * cd * icd2 + c.k[2] * a3 + c.k[3] * a1) + c.cy;
  if (!J) return;
  for (int j = 0; j < 6; ++j) {
    double dPx, dPy, dPz;
    if (j < 3) {
      const double* d = dRdr + j * 9;
      dPx = d[0] * X + d[1] * Y; dPy = d[3] * X + d[4] * Y; dPz = d[6] * X + d[7] * Y;
    } else {
      dPx = (j == 3); dPy = (j == 4); dPz = (j == 5);
    }
    const double dx = z * (dPx - x * dPz), dy = z * (dPy - y * dPz);
    const double dr2 = 2 * x * dx + 2 * y * dy;
    const double dcd = (c.k[0] + 2 * c.k[1] * r2 + 3 * c.k[4] * r4) * dr2;
    const double dicd2 = -icd2 * icd2 * (c.k[5] + 2 * c.k[6] * r2 + 3 * c.k[7] * r4) * dr2;
    const double da1 = 2 * (x * dy + y * dx);
    J[j] = c.fx * (dx * cd * icd2 + x * dcd * icd2 + x * cd * dicd2 + c.k[2] * da1 + c.k[3] * (dr2 + 4 * x * dx));
    J[6 + j] = c.fy * (dy * cd * icd2 + y * dcd * icd2 + y * cd * dicd2 + c.k[2] * (dr2 + 4 * y * dy) + c.k[3] * da1);
  }
}

// solve the symmetric N x N system A x = b by Gaussian elimination with partial pivoting (A is JtJ with a scaled diagonal)
template <int N>
__host__ __device__ inline bool solve_n(double* A, double* b, double* x) {
  for (int i = 0; i < N; ++i) {
    int p = i;
    for (int r = i + 1; r < N; ++r)
      if (fabs(A[r * N + i]) > fabs(A[p * N + i])) p = r;
    if (fabs(A[p * N + i]) < 1e-300) return false;
    if (p != i) {
      for (int k = 0; k < N; ++k) { const double t = A[i * N + k]; A[i * N + k] = A[p * N + k]; A[p * N + k] = t; }
      const double t = b[i]; b[i] = b[p]; b[p] = t;
    }
    for (int r = i + 1; r < N; ++r) {
      const double f = A[r * N + i] / A[i * N + i];
      for (int k = i; k < N; ++k) A[r * N + k] -= f * A[i * N + k];
      b[r] -= f * b[i];
    }
  }
  for (int i = N - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < N; ++k) s -= A[i * N + k] * x[k];
    x[i] = s / A[i * N + i];
  }
  return true;
}
__host__ __device__ inline bool solve6(double A[36], double b[6], double x[6]) { return solve_n<6>(A, b, x); }

template <class Par>
__host__ __device__ inline double reproj_norm(const Par& par, const Cam& cam, const Pts& P, const double prm[6]) {
  double R[9];
  rodrigues(prm, R, nullptr);
  double e2 = 0;
  for (int i = par.first(); i < P.n; i += par.step()) {
    double X, Y, u, v, pu, pv;
    P.get(i, X, Y, u, v);
    project(cam, R, nullptr, prm + 3, X, Y, pu, pv, nullptr);
    e2 += (pu - u) * (pu - u) + (pv - v) * (pv - v);
  }
  par.sum(&e2, 1);
  return sqrt(e2);
}

// one frame: returns 1 and (rvec, tvec) on success, 0 (zeros) for < 4 corners or a degenerate configuration
template <class Par>
__host__ __device__ inline int solve_frame(const Par& par, const Cam& cam, const Pts& P, double rv[3], double tv[3]) {
  rv[0] = rv[1] = rv[2] = tv[0] = tv[1] = tv[2] = 0.0;
  if (P.n < 4) return 0;                                     // inference.py:16-17
  {
    double bad = 0;
    for (int i = par.first(); i < P.n; i += par.step())
      if (P.kp[4 * i + 2] < 0 || P.kp[4 * i + 2] >= P.n_obj) bad += 1;
    par.sum(&bad, 1);
    if (bad > 0) return 0;
  }

  // ---- 1+2: normalised DLT homography  board plane (X, Y) -> normalised image (x, y) ----
  double cM[2] = {0, 0}, cm[2] = {0, 0};
  for (int i = par.first(); i < P.n; i += par.step()) {
    double X, Y, u, v, x, y;
    P.get(i, X, Y, u, v);
    undistort(cam, u, v, x, y);
    cM[0] += X; cM[1] += Y; cm[0] += x; cm[1] += y;
  }
  par.sum(cM, 2); par.sum(cm, 2);
  const double inv_n = 1.0 / P.n;
  cM[0] *= inv_n; cM[1] *= inv_n; cm[0] *= inv_n; cm[1] *= inv_n;
  // the board lies in z = 0: OpenCV's plane alignment is the identity and its translation is -centroid
  double sM[2] = {0, 0}, sm[2] = {0, 0};
  for (int i = par.first(); i < P.n; i += par.step()) {
    double X, Y, u, v, x, y;
    P.get(i, X, Y, u, v);
    undistort(cam, u, v, x, y);
    sM[0] += fabs(X - cM[0]); sM[1] += fabs(Y - cM[1]); sm[0] += fabs(x - cm[0]); sm[1] += fabs(y - cm[1]);
  }
  par.sum(sM, 2); par.sum(sm, 2);
  if (sM[0] < 1e-300 || sM[1] < 1e-300 || sm[0] < 1e-300 || sm[1] < 1e-300) return 0;      // collinear along an axis
  sM[0] = P.n / sM[0]; sM[1] = P.n / sM[1]; sm[0] = P.n / sm[0]; sm[1] = P.n / sm[1];
  double L[81];
  for (int i = 0; i < 81; ++i) L[i] = 0;
  for (int i = par.first(); i < P.n; i += par.step()) {
    double X, Y, u, v, x, y;
    P.get(i, X, Y, u, v);
    undistort(cam, u, v, x, y);
    // source points of the DLT: board coordinates minus their centroid (OpenCV's T_transform), scaled to unit mean deviation
    X = (X - cM[0]) * sM[0]; Y = (Y - cM[1]) * sM[1];
    x = (x - cm[0]) * sm[0]; y = (y - cm[1]) * sm[1];
    const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x}, Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
    for (int a = 0; a < 9; ++a)
      for (int b = a; b < 9; ++b) L[a * 9 + b] += Lx[a] * Lx[b] + Ly[a] * Ly[b];
  }
  for (int a = 0; a < 9; ++a) par.sum(L + a * 9 + a, 9 - a);
  for (int a = 0; a < 9; ++a)
    for (int b = 0; b < a; ++b) L[a * 9 + b] = L[b * 9 + a];
  double h0[9];
  smallest_eigvec9(L, h0);
  // H = inv(Hnorm_img) * H0 * Hnorm_board, where board coordinates are relative to the centroid
  double H[9];
  {
    // H0 * Hnorm2, Hnorm2 = [sM0 0 0; 0 sM1 0; 0 0 1] (centred source has zero mean)
    double T[9];
    for (int r = 0; r < 3; ++r) { T[r * 3] = h0[r * 3] * sM[0]; T[r * 3 + 1] = h0[r * 3 + 1] * sM[1]; T[r * 3 + 2] = h0[r * 3 + 2]; }
    // invHnorm = [1/sm0 0 cm0; 0 1/sm1 cm1; 0 0 1]
    for (int k = 0; k < 3; ++k) {
      H[k] = T[k] / sm[0] + cm[0] * T[6 + k];
      H[3 + k] = T[3 + k] / sm[1] + cm[1] * T[6 + k];
      H[6 + k] = T[6 + k];
    }
  }
  if (fabs(H[8]) < 1e-300) return 0;
  for (int k = 0; k < 8; ++k) H[k] /= H[8];
  H[8] = 1.0;
  // ---- 2b: refine H on the transfer error (cv::findHomography does this for > 4 points: HomographyRefineCallback, h33 = 1) ----
  // OpenCV stops its LM after 10 iterations; here a damped Gauss-Newton runs to convergence: the same least-squares minimum, which
  // is what makes the start of step 4 -- and therefore its 20-step trajectory -- match cv2's when step 4 does not fully converge.
  if (P.n > 4) {
    double lamh = 1e-3, prev = -1.0;
    for (int it = 0; it < 30; ++it) {
      double JtJ[64], JtE[8], e2 = 0;
      for (int i = 0; i < 64; ++i) JtJ[i] = 0;
      for (int i = 0; i < 8; ++i) JtE[i] = 0;
      for (int i = par.first(); i < P.n; i += par.step()) {
        double X, Y, u, v, x, y;
        P.get(i, X, Y, u, v);
        undistort(cam, u, v, x, y);
        X -= cM[0]; Y -= cM[1];
        const double ww = 1.0 / (H[6] * X + H[7] * Y + 1.0);
        const double xi = (H[0] * X + H[1] * Y + H[2]) * ww, yi = (H[3] * X + H[4] * Y + H[5]) * ww;
        const double ex = xi - x, ey = yi - y;
        const double Jx[8] = {X * ww, Y * ww, ww, 0, 0, 0, -X * ww * xi, -Y * ww * xi};
        const double Jy[8] = {0, 0, 0, X * ww, Y * ww, ww, -X * ww * yi, -Y * ww * yi};
        e2 += ex * ex + ey * ey;
        for (int a = 0; a < 8; ++a) {
          JtE[a] += Jx[a] * ex + Jy[a] * ey;
          for (int b = a; b < 8; ++b) JtJ[a * 8 + b] += Jx[a] * Jx[b] + Jy[a] * Jy[b];
        }
      }
      for (int a = 0; a < 8; ++a) par.sum(JtJ + a * 8 + a, 8 - a);
      par.sum(JtE, 8); par.sum(&e2, 1);
      for (int a = 0; a < 8; ++a)
        for (int b = 0; b < a; ++b) JtJ[a * 8 + b] = JtJ[b * 8 + a];
      if (prev >= 0.0 && e2 > prev) { lamh *= 10.0; } else { lamh = lamh * 0.1 > 1e-12 ? lamh * 0.1 : 1e-12; }
      prev = e2;
      double A[64], b8[8], d[8];
      for (int i = 0; i < 64; ++i) A[i] = JtJ[i];
      for (int k = 0; k < 8; ++k) { A[k * 9] *= 1.0 + lamh; b8[k] = JtE[k]; }
      if (!solve_n<8>(A, b8, d)) break;
      double dn = 0, hn = 0;
      for (int k = 0; k < 8; ++k) { H[k] -= d[k]; dn += d[k] * d[k]; hn += H[k] * H[k]; }
      if (dn <= 1e-26 * hn) break;
    }
  }
  // ---- 3: pose from the homography ----
  double prm[6];
  {
    const double n1 = sqrt(H[0] * H[0] + H[3] * H[3] + H[6] * H[6]), n2 = sqrt(H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
    const double i1 = 1.0 / fmax(n1, 2.220446049250313e-16), i2 = 1.0 / fmax(n2, 2.220446049250313e-16);
    const double h1[3] = {H[0] * i1, H[3] * i1, H[6] * i1}, h2[3] = {H[1] * i2, H[4] * i2, H[7] * i2};
    const double ts = 2.0 / fmax(n1 + n2, 2.220446049250313e-16);
    double t[3] = {H[2] * ts, H[5] * ts, H[8] * ts};
    const double h3[3] = {h1[1] * h2[2] - h1[2] * h2[1], h1[2] * h2[0] - h1[0] * h2[2], h1[0] * h2[1] - h1[1] * h2[0]};
    double M[9] = {h1[0], h2[0], h3[0], h1[1], h2[1], h3[1], h1[2], h2[2], h3[2]};
    matrix_to_rvec(M, prm);
    double R[9];
    rodrigues(prm, R, nullptr);
    // t += R * (-centroid)   (the homography was estimated on centred board coordinates)
    prm[3] = t[0] - (R[0] * cM[0] + R[1] * cM[1]);
    prm[4] = t[1] - (R[3] * cM[0] + R[4] * cM[1]);
    prm[5] = t[2] - (R[6] * cM[0] + R[7] * cM[1]);
  }
  for (int k = 0; k < 6; ++k)
    if (!isfinite(prm[k])) return 0;

  // ---- 4: Levenberg-Marquardt (CvLevMarq schedule) ----
  int lam = -3, iters = 0;
  double prev_err = reproj_norm(par, cam, P, prm);
  for (;;) {
    double JtJ[36], JtE[6], R[9], dR[27];
    for (int i = 0; i < 36; ++i) JtJ[i] = 0;
    for (int i = 0; i < 6; ++i) JtE[i] = 0;
    rodrigues(prm, R, dR);
    for (int i = par.first(); i < P.n; i += par.step()) {
      double X, Y, u, v, pu, pv, J[12];
      P.get(i, X, Y, u, v);
      project(cam, R, dR, prm + 3, X, Y, pu, pv, J);
      const double eu = pu - u, ev = pv - v;
      for (int a = 0; a < 6; ++a) {
        JtE[a] += J[a] * eu + J[6 + a] * ev;
        for (int b = a; b < 6; ++b) JtJ[a * 6 + b] += J[a] * J[b] + J[6 + a] * J[6 + b];
      }
    }
    for (int a = 0; a < 6; ++a) par.sum(JtJ + a * 6 + a, 6 - a);
    par.sum(JtE, 6);
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < a; ++b) JtJ[a * 6 + b] = JtJ[b * 6 + a];
    double prev[6], cand[6], err = prev_err;
    for (int k = 0; k < 6; ++k) prev[k] = prm[k];
    bool ok = true;
    for (;;) {
      double A[36], b[6], d[6];
      for (int i = 0; i < 36; ++i) A[i] = JtJ[i];
      const double mul = 1.0 + pow(10.0, (double)lam);
      for (int k = 0; k < 6; ++k) { A[k * 7] *= mul; b[k] = JtE[k]; }
      if (!solve6(A, b, d)) { ok = false; break; }
      for (int k = 0; k < 6; ++k) cand[k] = prev[k] - d[k];
      err = reproj_norm(par, cam, P, cand);
      if (!(err <= prev_err) && ++lam <= 16) continue;      // worse (or NaN): larger damping, same Jacobian
      break;
    }
    if (!ok) break;
    for (int k = 0; k < 6; ++k) prm[k] = cand[k];
    lam = lam - 1 < -16 ? -16 : lam - 1;
    double dn = 0, pn = 0;
    for (int k = 0; k < 6; ++k) { dn += (prm[k] - prev[k]) * (prm[k] - prev[k]); pn += prev[k] * prev[k]; }
    if (++iters >= 20 || sqrt(dn) < 1.1920928955078125e-07 * sqrt(pn)) break;
    prev_err = err;
  }
  for (int k = 0; k < 6; ++k)
    if (!isfinite(prm[k])) return 0;
  rv[0] = prm[0]; rv[1] = prm[1]; rv[2] = prm[2];
  tv[0] = prm[3]; tv[1] = prm[4]; tv[2] = prm[5];
  return 1;
}

}  // namespace pnp
}  // namespace dcu
