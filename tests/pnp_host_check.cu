// Test-only: runs the host side of deepcharuco_b200/csrc/pnp_core.cuh (the SAME functions the CUDA kernel executes per
// frame) on the CPU so that tests/test_pnp_host.py can compare the restatement with cv2.solvePnP without a GPU.
// Not part of the product library.  stdin: n_frames, then per frame:
//   K  fx fy cx cy  k0..k7  cols rows square_len   then K lines  x y id     (x, y as float32 values)
// stdout per frame: ret rvec[3] tvec[3]
#include <cstdio>
#include <vector>

#include "../deepcharuco_b200/csrc/pnp_core.cuh"

int main() {
  int n;
  if (scanf("%d", &n) != 1) return 1;
  for (int f = 0; f < n; ++f) {
    int K, cols, rows;
    double sq;
    dcu::pnp::Cam cam{};
    if (scanf("%d %lf %lf %lf %lf", &K, &cam.fx, &cam.fy, &cam.cx, &cam.cy) != 5) return 1;
    for (int i = 0; i < 8; ++i)
      if (scanf("%lf", &cam.k[i]) != 1) return 1;
    if (scanf("%d %d %lf", &cols, &rows, &sq) != 3) return 1;
    const int n_obj = (cols - 1) * (rows - 1);
    std::vector<float> obj(2 * n_obj);
    for (int p = 0; p < n_obj; ++p) {       // engine.cu: pnp_object_table (inference.py:20-23)
      obj[2 * p] = (float)((double)((p % (rows - 1)) + 1) * sq);
      obj[2 * p + 1] = (float)((double)((p / (rows - 1)) + 1) * sq);
    }
    std::vector<int32_t> kp(4 * (K > 0 ? K : 1));
    std::vector<float> xy(2 * (K > 0 ? K : 1));
    for (int i = 0; i < K; ++i) {
      double x, y; int id;
      if (scanf("%lf %lf %d", &x, &y, &id) != 3) return 1;
      xy[2 * i] = (float)x; xy[2 * i + 1] = (float)y;
      kp[4 * i] = (int)x; kp[4 * i + 1] = (int)y; kp[4 * i + 2] = id; kp[4 * i + 3] = 0;
    }
    dcu::pnp::Pts P;
    P.kp = kp.data(); P.xy = xy.data(); P.obj = obj.data(); P.n = K; P.n_obj = n_obj;
    double rv[3], tv[3];
    const int ret = dcu::pnp::solve_frame(dcu::pnp::Serial(), cam, P, rv, tv);
    printf("%d %.17g %.17g %.17g %.17g %.17g %.17g\n", ret, rv[0], rv[1], rv[2], tv[0], tv[1], tv[2]);
  }
  return 0;
}
