// Test-only: compiles the kernel's per-node CRPS math (rc_crps_node.cuh) for the host so the
// algebra of the hand-derived gradient can be checked against the reference fixtures without a GPU.
#include <math.h>
#include "rc_b200.h"
#include "rc_crps_node.cuh"

extern "C" void crps_rows_host(const float* pred, const float* y, float* dpred, double* loss_mean, int* nvalid,
                               int n, int kind, int raw, float u, float xi, float t) {
  const int w = rc::loss_width(kind);
  int cnt = 0;
  for (int i = 0; i < n; ++i) cnt += !isnan(y[i]);
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    float g[5] = {0, 0, 0, 0, 0};
    if (!isnan(y[i])) {
      acc += rc::crps_node(pred + (size_t)i * w, y[i], kind, raw, u, xi, t, g);
      for (int j = 0; j < w; ++j) dpred[(size_t)i * w + j] = g[j] / (float)cnt;
    } else {
      for (int j = 0; j < w; ++j) dpred[(size_t)i * w + j] = 0.f;
    }
  }
  *loss_mean = cnt ? acc / cnt : NAN;
  *nvalid = cnt;
}
