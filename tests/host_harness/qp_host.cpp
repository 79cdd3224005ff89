// qp_host.cpp -- TEST HARNESS (not shipped, not a fallback): compiles the per-demand solver of the CUDA allocator kernel
// (ml4ca_b200/csrc/qp_slsqp.cuh, the very code each GPU thread runs) for the host, so that its path can be checked
// against the reference's outputs in tests/golden/qp_config1.npz on a box without a GPU (tests/test_qp_host.py).
//   qp_host <float|double|mixed|group> <in.bin> <out.bin> [w0 .. w10 fuel]     (objective switches of :108,116-150)
// "group" = the 8-lanes-per-demand formulation of qp_group.cuh (the alternative kernel mapping) on its host backend.
// in.bin : int64 n, then tau[3][n], prev[5][n] as float64.   out.bin: x[8][n] float64 (raw, before the clean-up),
// then mode[n], iter[n], mask[n] as int32.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../ml4ca_b200/csrc/qp_group.cuh"

using namespace ml4ca::slsqp;

static Objective g_obj = default_objective();

template <typename real, typename greal>
static void run(int64_t n, const double* tau, const double* prev, double* x, int32_t* mode, int32_t* iter, int32_t* mask) {
  const Objective obj = g_obj;
  std::vector<greal> G(45);
  for (int64_t j = 0; j < n; ++j) {
    real t[3], p[5];
    for (int i = 0; i < 3; ++i) t[i] = (real)tau[i * n + j];
    for (int i = 0; i < 5; ++i) p[i] = (real)prev[i * n + j];
    Problem<real> P;
    make_problem(t, p, P);
    State<real> S;
    slsqp_init(P, obj, S);
    while (!slsqp_iterate<real, greal>(P, obj, S, G.data(), 1)) {
    }
    for (int i = 0; i < 8; ++i) x[i * n + j] = (double)S.pt.x[i];
    mode[j] = S.mode, iter[j] = S.iter;
    mask[j] = (int32_t)active_mask(P, S.pt.x, (real)1e-5);
  }
}

static void run_group(int64_t n, const double* tau, const double* prev, double* x, int32_t* mode, int32_t* iter, int32_t* mask) {
  const Objective obj = g_obj;
  HostB b;
  for (int64_t j = 0; j < n; ++j) {
    double t[3], p[5];
    for (int i = 0; i < 3; ++i) t[i] = tau[i * n + j];
    for (int i = 0; i < 5; ++i) p[i] = prev[i * n + j];
    GroupSolver<HostB> S;
    S.set_problem(b, t, p, obj);
    S.init(b);
    while (!S.iterate(b, true)) {
    }
    double xs[8];
    for (int i = 0; i < 8; ++i) xs[i] = S.x.v[i], x[i * n + j] = xs[i];
    mode[j] = S.mode, iter[j] = S.iter;
    Problem<double> P;
    make_problem(t, p, P);
    mask[j] = (int32_t)active_mask(P, xs, 1e-5);
  }
}

int main(int argc, char** argv) {
  if (argc != 4 && argc != 16) return 2;
  if (argc == 16) {
    for (int i = 0; i < 3; ++i) g_obj.ws[i] = (float)atof(argv[4 + i]), g_obj.wf[i] = (float)atof(argv[7 + i]), g_obj.wd[i] = (float)atof(argv[12 + i]);
    g_obj.wa[0] = (float)atof(argv[10]), g_obj.wa[1] = (float)atof(argv[11]);
    g_obj.fuel = atoi(argv[15]);
  }
  FILE* fi = fopen(argv[2], "rb");
  if (!fi) return 3;
  int64_t n = 0;
  if (fread(&n, sizeof n, 1, fi) != 1) return 4;
  std::vector<double> tau(3 * n), prev(5 * n), x(8 * n);
  std::vector<int32_t> mode(n), iter(n), mask(n);
  if (fread(tau.data(), sizeof(double), 3 * n, fi) != (size_t)(3 * n)) return 4;
  if (fread(prev.data(), sizeof(double), 5 * n, fi) != (size_t)(5 * n)) return 4;
  fclose(fi);
  if (!strcmp(argv[1], "group")) run_group(n, tau.data(), prev.data(), x.data(), mode.data(), iter.data(), mask.data());
  else if (!strcmp(argv[1], "double")) run<double, double>(n, tau.data(), prev.data(), x.data(), mode.data(), iter.data(), mask.data());
  else if (!strcmp(argv[1], "mixed")) run<double, float>(n, tau.data(), prev.data(), x.data(), mode.data(), iter.data(), mask.data());
  else run<float, float>(n, tau.data(), prev.data(), x.data(), mode.data(), iter.data(), mask.data());
  FILE* fo = fopen(argv[3], "wb");
  if (!fo) return 5;
  fwrite(x.data(), sizeof(double), 8 * n, fo);
  fwrite(mode.data(), sizeof(int32_t), n, fo);
  fwrite(iter.data(), sizeof(int32_t), n, fo);
  fwrite(mask.data(), sizeof(int32_t), n, fo);
  fclose(fo);
  return 0;
}
