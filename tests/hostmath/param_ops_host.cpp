// tests/hostmath/param_ops_host.cpp -- TEST INFRASTRUCTURE: the product's csrc/param_ops.cuh compiled for the host.
#include <math.h>

#include "../../gaussian-opacity-fields_b200/csrc/param_ops.cuh"

extern "C" {
void hm_activate(int P, const float* s_raw, const float* q, const float* o_raw, const float* f, float* scales, float* rot, float* op) {
  for (int i = 0; i < P; ++i) {
    const PoActivated a = po_activate(s_raw + 3 * i, q + 4 * i, o_raw[i], f[i]);
    for (int k = 0; k < 3; ++k) scales[3 * i + k] = a.scales[k];
    for (int k = 0; k < 4; ++k) rot[4 * i + k] = a.rot[k];
    op[i] = a.opacity;
  }
}
void hm_activate_backward(int P, const float* s_raw, const float* q, const float* o_raw, const float* f, const float* g_scales,
                          const float* g_rot, const float* g_op, float* d_s, float* d_q, float* d_o) {
  for (int i = 0; i < P; ++i)
    po_activate_backward(s_raw + 3 * i, q + 4 * i, o_raw[i], f[i], g_scales + 3 * i, g_rot + 4 * i, g_op[i], d_s + 3 * i, d_q + 4 * i, d_o + i);
}
void hm_adam(long n, float* p, float* m, float* v, const float* g, double lr, double b1, double b2, double eps, int step) {
  const double bias1 = 1.0 - pow(b1, (double)step), bias2 = 1.0 - pow(b2, (double)step);
  for (long i = 0; i < n; ++i)
    po_adam(p + i, m + i, v + i, g[i], (float)b2, (float)(1.0 - b1), (float)(1.0 - b2), (float)eps, (float)(lr / bias1), (float)sqrt(bias2));
}
}
