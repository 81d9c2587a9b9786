// Builds oracle/_ref/libeq_ufc_ref.so: the reference's OWN FFC-generated element
// kernels, compiled in place from /root/reference/fenics/*.h (no copy of the
// sources enters this repository) under the interface shim in oracle/shim/.
// TEST INFRASTRUCTURE ONLY: used to pin oracle/eq_oracle.c's restated kernels
// and to generate tests/golden/*.json.
#include <ufc.h>
#include "hslD.h"                 // /root/reference/fenics/hslD.h
#include "AdvectionDiffusion.h"   // /root/reference/fenics/AdvectionDiffusion.h
#include "boundary.h"             // /root/reference/fenics/boundary.h
#include "hsl.h"                  // isotropic variants (cross-check only)
#include "hslRobin.h"

#define REF_API extern "C" __attribute__((visibility("default")))

// fenics/hslD.h:3123  w = {D11[3], D22[3], D12[3], D, dt}
REF_API void ref_hsld_cell_a(double* A, const double* d11, const double* d22, const double* d12,
                             double D, double dt, const double* xy)
{
  const double* w[7] = {d11, d22, d12, &D, &dt, nullptr, nullptr};
  hsld_cell_integral_0_otherwise().tabulate_tensor(A, w, xy, 0);
}
// fenics/hslD.h:3284 (ds(1), w[5]=r1) and :3375 (ds(2), w[6]=r2)
REF_API void ref_hsld_facet_a(double* A, double dt, double r, const double* xy, int facet, int marker)
{
  const double* w[7] = {nullptr, nullptr, nullptr, nullptr, &dt, &r, &r};
  if (marker == 1) hsld_exterior_facet_integral_0_1().tabulate_tensor(A, w, xy, (std::size_t)facet, 0);
  else             hsld_exterior_facet_integral_0_2().tabulate_tensor(A, w, xy, (std::size_t)facet, 0);
}
// fenics/hslD.h:3466  w = {u0[3], dt, f}
REF_API void ref_hsld_cell_L(double* b, const double* u0, double dt, double f, const double* xy)
{
  const double* w[7] = {u0, &dt, &f, nullptr, nullptr, nullptr, nullptr};
  hsld_cell_integral_1_otherwise().tabulate_tensor(b, w, xy, 0);
}
// fenics/hslD.h:3554 (w[3]=r1,w[4]=s1) and :3634 (w[5]=r2,w[6]=s2)
REF_API void ref_hsld_facet_L(double* b, double dt, double r, double s, const double* xy, int facet, int marker)
{
  const double* w[7] = {nullptr, &dt, nullptr, &r, &s, &r, &s};
  if (marker == 1) hsld_exterior_facet_integral_1_1().tabulate_tensor(b, w, xy, (std::size_t)facet, 0);
  else             hsld_exterior_facet_integral_1_2().tabulate_tensor(b, w, xy, (std::size_t)facet, 0);
}
// fenics/boundary.h:2652  w = {u[3]}
REF_API double ref_boundary_facet(const double* u, const double* xy, int facet)
{
  const double* w[1] = {u};
  double A[1] = {0.0};
  boundary_exterior_facet_integral_0_otherwise().tabulate_tensor(A, w, xy, (std::size_t)facet, 0);
  return A[0];
}
// fenics/AdvectionDiffusion.h:2246  w = {dt, D, v, r1, r2}
REF_API void ref_ad_cell_a(double* A, double dt, double D, double v, const double* xc)
{
  const double* w[5] = {&dt, &D, &v, nullptr, nullptr};
  advectiondiffusion_cell_integral_0_otherwise().tabulate_tensor(A, w, xc, 0);
}
REF_API void ref_ad_facet_a(double* A, double dt, double r, int facet, int marker)
{
  const double* w[5] = {&dt, nullptr, nullptr, &r, &r};
  double xc[2] = {0.0, 1.0};
  if (marker == 1) advectiondiffusion_exterior_facet_integral_0_1().tabulate_tensor(A, w, xc, (std::size_t)facet, 0);
  else             advectiondiffusion_exterior_facet_integral_0_2().tabulate_tensor(A, w, xc, (std::size_t)facet, 0);
}
// fenics/AdvectionDiffusion.h:2444  w = {u0[2], dt, D, v, r1, s1, r2, s2}
REF_API void ref_ad_cell_L(double* b, const double* u0, double dt, double D, double v, const double* xc)
{
  const double* w[8] = {u0, &dt, &D, &v, nullptr, nullptr, nullptr, nullptr};
  advectiondiffusion_cell_integral_1_otherwise().tabulate_tensor(b, w, xc, 0);
}
REF_API void ref_ad_facet_L(double* b, const double* u0, double dt, double r, double s, int facet, int marker)
{
  const double* w[8] = {u0, &dt, nullptr, nullptr, &r, &s, &r, &s};
  double xc[2] = {0.0, 1.0};
  if (marker == 1) advectiondiffusion_exterior_facet_integral_1_1().tabulate_tensor(b, w, xc, (std::size_t)facet, 0);
  else             advectiondiffusion_exterior_facet_integral_1_2().tabulate_tensor(b, w, xc, (std::size_t)facet, 0);
}
// fenics/hsl.h:3123 (isotropic a: w = {D, dt}?) -- exposed through the form's
// own coefficient order, see tests/test_oracle_ref.py
REF_API void ref_hsl_cell_a(double* A, const double* const* w, const double* xy)
{
  hsl_cell_integral_0_otherwise().tabulate_tensor(A, w, xy, 0);
}
REF_API void ref_hslrobin_cell_a(double* A, const double* const* w, const double* xy)
{
  hslrobin_cell_integral_0_otherwise().tabulate_tensor(A, w, xy, 0);
}
