// Element stiffness of the reference's 2-node bar (axial + transverse spring), evaluated with the
// rounding sequence of the numpy expression in src/fea_solver.py:30-68:
//   v = p2 - p1; L = sqrt((vx*vx + vy*vy) + vz*vz); Ls = max(L, 1e-12); n = v / Ls
//   k_ax = (E*A)/Ls; k_b = ((12*E)*I)/Ls**3
//   S_ij = (n_i*n_j)*k_ax + (delta_ij - n_j*n_i)*k_b          K_e = [[S,-S],[-S,S]]
// Every product is rounded before the following add (the explicit __d*_rn intrinsics are never
// contracted into FMAs by nvcc).  Ls**3 is the correctly rounded cube (numpy's pow is within
// 1 ulp of it; the C++ reference's L*L*L within 1 ulp as well, src/fea_petsc.cpp:117).
#pragma once
#include <cuda_runtime.h>

struct Sym3 {  // symmetric 3x3: xx, xy, xz, yy, yz, zz
  double xx, xy, xz, yy, yz, zz;
};

__device__ __forceinline__ double myc_cube_rn(double x) {
  // x^3 via double-double: (ph + pl) = x*x exactly, then (ph + pl)*x rounded once at the end
  const double ph = __dmul_rn(x, x);
  const double pl = __fma_rn(x, x, -ph);
  const double qh = __dmul_rn(ph, x);
  const double ql = __fma_rn(ph, x, -qh);
  return __dadd_rn(qh, __fma_rn(pl, x, ql));
}

struct BarConsts {
  double EA;      // E*A rounded once        (fea_solver.py:45)
  double c12EI;   // (12*E)*I, two roundings (fea_solver.py:58)
};

__device__ __forceinline__ BarConsts myc_bar_consts(double E, double A, double I) {
  BarConsts c;
  c.EA = __dmul_rn(E, A);
  c.c12EI = __dmul_rn(__dmul_rn(12.0, E), I);
  return c;
}

// Returns S (6 unique entries) and the unclamped length through *L_out.
__device__ __forceinline__ Sym3 myc_bar_block(double p1x, double p1y, double p1z, double p2x,
                                              double p2y, double p2z, const BarConsts& c,
                                              double* L_out) {
  const double vx = __dsub_rn(p2x, p1x), vy = __dsub_rn(p2y, p1y), vz = __dsub_rn(p2z, p1z);
  const double ss = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
  const double L = __dsqrt_rn(ss);
  *L_out = L;
  const double Ls = (L < 1e-12) ? 1e-12 : L;
  const double nx = __ddiv_rn(vx, Ls), ny = __ddiv_rn(vy, Ls), nz = __ddiv_rn(vz, Ls);
  const double kax = __ddiv_rn(c.EA, Ls);
  const double kb = __ddiv_rn(c.c12EI, myc_cube_rn(Ls));
  const double txx = __dmul_rn(nx, nx), txy = __dmul_rn(nx, ny), txz = __dmul_rn(nx, nz);
  const double tyy = __dmul_rn(ny, ny), tyz = __dmul_rn(ny, nz), tzz = __dmul_rn(nz, nz);
  Sym3 s;
  s.xx = __dadd_rn(__dmul_rn(txx, kax), __dmul_rn(__dsub_rn(1.0, txx), kb));
  s.xy = __dadd_rn(__dmul_rn(txy, kax), __dmul_rn(__dsub_rn(0.0, txy), kb));
  s.xz = __dadd_rn(__dmul_rn(txz, kax), __dmul_rn(__dsub_rn(0.0, txz), kb));
  s.yy = __dadd_rn(__dmul_rn(tyy, kax), __dmul_rn(__dsub_rn(1.0, tyy), kb));
  s.yz = __dadd_rn(__dmul_rn(tyz, kax), __dmul_rn(__dsub_rn(0.0, tyz), kb));
  s.zz = __dadd_rn(__dmul_rn(tzz, kax), __dmul_rn(__dsub_rn(1.0, tzz), kb));
  return s;
}

__device__ __forceinline__ double myc_sym3_get(const Sym3& s, int i, int j) {
  const int a = i < j ? i : j, b = i < j ? j : i;
  // (0,0)=xx (0,1)=xy (0,2)=xz (1,1)=yy (1,2)=yz (2,2)=zz
  return a == 0 ? (b == 0 ? s.xx : (b == 1 ? s.xy : s.xz)) : (a == 1 ? (b == 1 ? s.yy : s.yz) : s.zz);
}
