/*
 * pcg_port.c -- TEST INFRASTRUCTURE ONLY (CPU baseline; never linked into the product).
 *
 * C/OpenMP restatement of what the reference's PETSc path runs per load step once K is assembled
 * (src/fea_petsc.cpp:303-341 and src/fea_petsc_parallel.cpp:300-351 with `-pc_type jacobi`):
 *   MatZeroRowsColumnsIS(K, known, diag = 1, x, b)   -> rows/cols of known DOFs zeroed, unit
 *                                                       diagonal, b -= K[:,known] x_known,
 *                                                       b[known] = x_known          (:309-314)
 *   1e-12 added to EVERY diagonal entry                                              (:320-325)
 *   KSPCG + PCJACOBI, x0 = 0                                                         (:328-341)
 * PETSc 3.24.1 is not vendored in the reference and not installed in this image, so this is a
 * restatement of the published algorithm (Hestenes-Stiefel PCG with diagonal scaling), not a run
 * of PETSc.  Convergence is tested on the UNpreconditioned residual, ||r||2 <= rtol*||b||2, to be
 * like for like with the GPU solver (PETSc's default for KSPCG is the preconditioned norm).
 * Threads: OpenMP static row partition = the row-block partition of MPIAIJ under `mpirun -np P`.
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC oracle/pcg_port.c -o oracle/_build/libpcg_port.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int pcg_port_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* In-place Dirichlet elimination on a CSR copy + RHS.  is_known[n]: 1 for prescribed DOFs,
 * x_known[n]: prescribed values (0 elsewhere).  val is modified, b is written.  PETSc inserts the
 * diagonal entries it needs (MatZeroRowsColumns' unit diagonal, the MatSetValue(i,i,1e-12) loop);
 * a CSR cannot grow, so rows without a stored diagonal (isolated nodes have empty rows) get theirs
 * in extra_diag[i]. */
void pcg_port_zero_rows_cols(int64_t n, const int32_t* rp, const int32_t* ci, double* val,
                             const uint8_t* is_known, const double* x_known, double reg, double* b,
                             double* extra_diag) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int has_diag = 0;
    for (int32_t k = rp[i]; k < rp[i + 1]; ++k) has_diag |= (ci[k] == i);
    extra_diag[i] = has_diag ? 0.0 : (is_known[i] ? 1.0 + reg : reg);
    if (is_known[i]) {
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k) val[k] = (ci[k] == i) ? 1.0 + reg : 0.0;
      b[i] = x_known[i];
    } else {
      double s = 0.0;
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k) {
        const int32_t c = ci[k];
        if (is_known[c]) { s += val[k] * x_known[c]; val[k] = 0.0; }
        else if (c == i) val[k] += reg;
      }
      b[i] = -s;
    }
  }
}

/* Jacobi-PCG.  Returns the iteration count; *relres = ||r||/||b|| at exit.  If max_iters is hit the
 * solve simply stops (used for bounded timing samples). */
int64_t pcg_port_solve(int64_t n, const int32_t* rp, const int32_t* ci, const double* val,
                       const double* extra_diag, const double* b, double rtol, int64_t max_iters, double* x,
                       double* relres) {
  double* r = (double*)malloc(sizeof(double) * n);
  double* z = (double*)malloc(sizeof(double) * n);
  double* p = (double*)malloc(sizeof(double) * n);
  double* Ap = (double*)malloc(sizeof(double) * n);
  double* dinv = (double*)malloc(sizeof(double) * n);
  double bb = 0.0, rz = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : bb, rz)
  for (int64_t i = 0; i < n; ++i) {
    double d = extra_diag[i];
    for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
      if (ci[k] == i) d += val[k];
    dinv[i] = d != 0.0 ? 1.0 / d : 0.0;
    x[i] = 0.0;
    r[i] = b[i];
    z[i] = dinv[i] * r[i];
    p[i] = z[i];
    bb += b[i] * b[i];
    rz += r[i] * z[i];
  }
  int64_t it = 0;
  double rr = bb;
  while (it < max_iters && rr > rtol * rtol * bb) {
    double pAp = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pAp)
    for (int64_t i = 0; i < n; ++i) {
      double s = extra_diag[i] * p[i];
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k) s += val[k] * p[ci[k]];
      Ap[i] = s;
      pAp += p[i] * s;
    }
    const double alpha = rz / pAp;
    double rz_new = 0.0;
    rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new, rr)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * Ap[i];
      z[i] = dinv[i] * r[i];
      rz_new += r[i] * z[i];
      rr += r[i] * r[i];
    }
    const double beta = rz_new / rz;
    rz = rz_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    ++it;
  }
  *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  free(r); free(z); free(p); free(Ap); free(dinv);
  return it;
}
