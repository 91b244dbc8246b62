// CPD-ALS driver with the MTTKRP on the GPU (drop-in for splatt_cpd_als).
//
// Follows the iteration of the reference's cpd_als_iterate (src/cpd.c:271-387):
//   per mode: M1 = MTTKRP -> A_m = M1 * (hadamard of the other Grams)^-1
//             -> column normalise (2-norm in iteration 0, max-norm afterwards)
//             -> Gram update;   per iteration: fit from the last mode's M1.
// Per the north star the small dense algebra stays on the host; it is written
// here in plain C++ (Cholesky / triangular solves on R x R, R = rank) so the
// library has no BLAS/LAPACK dependency.  Factor matrices are kept resident on
// the device: only the matrix updated in a mode step crosses PCIe (H2D), plus
// the MTTKRP result (D2H).
#include "common.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

// reference: src/util.c:15-23 (two rand() draws per value)
double rand_val() {
  double v = 3.0 * ((double)rand() / (double)RAND_MAX);
  if (rand() % 2 == 0) v *= -1;
  return v;
}

// G = A^T A, upper triangle in row-major (what the reference's syrk call leaves,
// src/matrix.c:414-455); lower triangle is not referenced by consumers.
void gram(const double * A, uint64_t I, int R, double * G) {
  std::vector<double> acc((size_t)R * R, 0.0);
#pragma omp parallel
  {
    std::vector<double> loc((size_t)R * R, 0.0);
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < (int64_t)I; ++i) {
      const double * a = A + (size_t)i * R;
      for (int p = 0; p < R; ++p) {
        const double ap = a[p];
        double * row = loc.data() + (size_t)p * R;
        for (int q = p; q < R; ++q) row[q] += ap * a[q];
      }
    }
#pragma omp critical
    for (size_t x = 0; x < loc.size(); ++x) acc[x] += loc[x];
  }
  memcpy(G, acc.data(), sizeof(double) * R * R);
}

// Normal-equation matrix: Hadamard of the other modes' Grams, symmetrised.
// NOTE: in the reference the `1 + reg` written on the diagonal is immediately
// overwritten by the row fill (src/matrix.c:45-51), i.e. the regularisation
// parameter has no effect; that effective behaviour is reproduced.
void form_normal_matrix(const std::vector<std::vector<double>> & ata, int mode, int N, int R,
                        double * neq) {
  for (int x = 0; x < R * R; ++x) neq[x] = 1.0;
  for (int m = 0; m < N; ++m) {
    if (m == mode) continue;
    for (int i = 0; i < R; ++i)
      for (int j = i; j < R; ++j) neq[j + i * R] *= ata[m][j + i * R];
  }
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < i; ++j) neq[j + i * R] = neq[i + j * R];
}

// In-place Cholesky G = L L^T (lower, row-major).  false if not SPD.
bool cholesky(double * G, int R) {
  for (int j = 0; j < R; ++j) {
    double d = G[j + j * R];
    for (int k = 0; k < j; ++k) d -= G[k + j * R] * G[k + j * R];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    G[j + j * R] = d;
    for (int i = j + 1; i < R; ++i) {
      double s = G[j + i * R];
      for (int k = 0; k < j; ++k) s -= G[k + i * R] * G[k + j * R];
      G[j + i * R] = s / d;
    }
  }
  return true;
}

// rows of X <- rows of X * (L L^T)^-1
void cholesky_solve_rows(const double * L, int R, double * X, uint64_t I) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)I; ++i) {
    double * x = X + (size_t)i * R;
    for (int p = 0; p < R; ++p) {          // L y = x
      double s = x[p];
      for (int k = 0; k < p; ++k) s -= L[k + p * R] * x[k];
      x[p] = s / L[p + p * R];
    }
    for (int p = R - 1; p >= 0; --p) {     // L^T z = y
      double s = x[p];
      for (int k = p + 1; k < R; ++k) s -= L[p + k * R] * x[k];
      x[p] = s / L[p + p * R];
    }
  }
}

// Minimum-norm least squares for a symmetric (possibly singular) G: the role of
// the reference's GELSS fallback (src/matrix.c:566-603).  Jacobi eigen-solve,
// pseudo-inverse with the LAPACK default cut-off (rcond < 0 -> machine eps).
void pinv_solve_rows(const double * Gin, int R, double * X, uint64_t I) {
  std::vector<double> A(Gin, Gin + (size_t)R * R), V((size_t)R * R, 0.0);
  for (int i = 0; i < R; ++i) V[i + i * R] = 1.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int p = 0; p < R; ++p)
      for (int q = p + 1; q < R; ++q) off += A[q + p * R] * A[q + p * R];
    if (off < 1e-300) break;
    for (int p = 0; p < R; ++p)
      for (int q = p + 1; q < R; ++q) {
        const double apq = A[q + p * R];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[q + q * R] - A[p + p * R]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < R; ++k) {
          const double akp = A[p + k * R], akq = A[q + k * R];
          A[p + k * R] = c * akp - s * akq;
          A[q + k * R] = s * akp + c * akq;
        }
        for (int k = 0; k < R; ++k) {
          const double apk = A[k + p * R], aqk = A[k + q * R];
          A[k + p * R] = c * apk - s * aqk;
          A[k + q * R] = s * apk + c * aqk;
        }
        for (int k = 0; k < R; ++k) {
          const double vkp = V[p + k * R], vkq = V[q + k * R];
          V[p + k * R] = c * vkp - s * vkq;
          V[q + k * R] = s * vkp + c * vkq;
        }
      }
  }
  double dmax = 0;
  for (int i = 0; i < R; ++i) dmax = std::max(dmax, std::fabs(A[i + i * R]));
  const double cut = dmax * 2.220446049250313e-16;
  std::vector<double> P((size_t)R * R, 0.0);
  int erank = 0;
  for (int e = 0; e < R; ++e) {
    const double d = A[e + e * R];
    if (std::fabs(d) <= cut) continue;
    ++erank;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < R; ++j) P[j + i * R] += V[e + i * R] * V[e + j * R] / d;
  }
  printf("SPLATT:   pseudo-inverse effective rank: %d\n", erank);
#pragma omp parallel
  {
    std::vector<double> tmp(R);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < (int64_t)I; ++i) {
      double * x = X + (size_t)i * R;
      for (int j = 0; j < R; ++j) {
        double s = 0;
        for (int k = 0; k < R; ++k) s += x[k] * P[j + k * R];
        tmp[j] = s;
      }
      memcpy(x, tmp.data(), sizeof(double) * R);
    }
  }
}

// reference: p_mat_2norm / p_mat_maxnorm (src/matrix.c:86-199)
void normalize_cols(double * A, uint64_t I, int R, double * lambda, bool two_norm) {
  std::vector<double> acc(R, 0.0);
#pragma omp parallel
  {
    std::vector<double> loc(R, 0.0);
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < (int64_t)I; ++i) {
      const double * a = A + (size_t)i * R;
      if (two_norm) for (int j = 0; j < R; ++j) loc[j] += a[j] * a[j];
      else for (int j = 0; j < R; ++j) loc[j] = std::max(loc[j], a[j]);
    }
#pragma omp critical
    for (int j = 0; j < R; ++j) acc[j] = two_norm ? acc[j] + loc[j] : std::max(acc[j], loc[j]);
  }
  for (int j = 0; j < R; ++j) lambda[j] = two_norm ? std::sqrt(acc[j]) : std::max(acc[j], 1.0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)I; ++i) {
    double * a = A + (size_t)i * R;
    for (int j = 0; j < R; ++j) a[j] /= lambda[j];
  }
}

// reference: p_kruskal_norm src/cpd.c:116-152
double kruskal_norm(const std::vector<std::vector<double>> & ata, const double * lambda, int N,
                    int R) {
  std::vector<double> av((size_t)R * R, 1.0);
  for (int m = 0; m < N; ++m)
    for (int i = 0; i < R; ++i)
      for (int j = i; j < R; ++j) av[j + i * R] *= ata[m][j + i * R];
  double nm = 0;
  for (int i = 0; i < R; ++i) {
    nm += av[i + i * R] * lambda[i] * lambda[i];
    for (int j = i + 1; j < R; ++j) nm += av[j + i * R] * lambda[i] * lambda[j] * 2;
  }
  return std::fabs(nm);
}

// reference: p_tt_kruskal_inner src/cpd.c:171-218
double kruskal_inner(const double * last, const double * m1, uint64_t I, int R,
                     const double * lambda) {
  std::vector<double> acc(R, 0.0);
#pragma omp parallel
  {
    std::vector<double> loc(R, 0.0);
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < (int64_t)I; ++i)
      for (int r = 0; r < R; ++r) loc[r] += last[r + (size_t)i * R] * m1[r + (size_t)i * R];
#pragma omp critical
    for (int r = 0; r < R; ++r) acc[r] += loc[r];
  }
  double inner = 0;
  for (int r = 0; r < R; ++r) inner += acc[r] * lambda[r];
  return inner;
}

double csf_frobsq(const splatt_csf * t) {   // reference: src/csf.c:817-851
  double norm = 0;
  const int N = (int)t->nmodes;
  for (uint64_t tile = 0; tile < t->ntiles; ++tile) {
    const double * v = t->pt[tile].vals;
    if (!v) continue;
    const uint64_t n = t->pt[tile].nfibs[N - 1];
#pragma omp parallel for reduction(+ : norm) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) norm += v[i] * v[i];
  }
  return norm;
}

}  // namespace

extern "C" {

int splatt_cpd_als(splatt_csf const * const tensors, splatt_idx_t const nfactors,
                   double const * const options, splatt_kruskal * factored) {
  if (!tensors || !options || !factored || nfactors == 0) {
    fprintf(stderr, "SPLATT: splatt_cpd_als: bad arguments\n");
    return SPLATT_ERROR_BADINPUT;
  }
  const int N = (int)tensors[0].nmodes;
  const int R = (int)nfactors;
  const int ldm = R + (R & 1);
  const int verbosity = (int)options[SPLATT_OPTION_VERBOSITY];
  uint64_t dims[SPB200_MAXN], maxdim = 0;
  for (int m = 0; m < N; ++m) { dims[m] = tensors[0].dims[m]; maxdim = std::max(maxdim, dims[m]); }

  // device mirror (once) -- reference: splatt_mttkrp_alloc_ws at src/cpd.c:304
  splatt_b200_build_opts bo;
  memset(&bo, 0, sizeof(bo));
  bo.device = -1;
  bo.verbosity = verbosity;
  const char * lay = getenv("SPLATT_B200_LAYOUT");
  bo.layout = (lay && !strcmp(lay, "asgiven")) ? SPLATT_B200_LAYOUT_ASGIVEN
                                               : SPLATT_B200_LAYOUT_ALLROOT;
  splatt_b200_tensor * T = nullptr;
  int rc = splatt_b200_tensor_from_csf(tensors, (int)options[SPLATT_OPTION_CSF_ALLOC], &bo, &T);
  if (rc != SPLATT_SUCCESS) return rc;

  // factor matrices: random init in the reference's draw order (src/cpd.c:36-40)
  double * mats[SPB200_MAXN] = {nullptr};
  double * d_mats[SPB200_MAXN] = {nullptr};
  double * d_out = nullptr;
  double * m1 = nullptr;          // pinned MTTKRP result
  double * lambda = static_cast<double *>(malloc(sizeof(double) * R));
  cudaStream_t stream = nullptr;
  bool ok = lambda != nullptr;
  for (int m = 0; m < N && ok; ++m) {
    mats[m] = static_cast<double *>(malloc(sizeof(double) * dims[m] * R));
    ok = mats[m] != nullptr;
    if (ok) for (uint64_t x = 0; x < dims[m] * (uint64_t)R; ++x) mats[m][x] = rand_val();
  }
  auto h2d = [&](int m) -> cudaError_t {
    if (ldm == R) return cudaMemcpyAsync(d_mats[m], mats[m], dims[m] * (size_t)R * 8,
                                         cudaMemcpyHostToDevice, stream);
    return cudaMemcpy2DAsync(d_mats[m], (size_t)ldm * 8, mats[m], (size_t)R * 8, (size_t)R * 8,
                             dims[m], cudaMemcpyHostToDevice, stream);
  };
  ok = ok && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int m = 0; m < N && ok; ++m) {
    ok = cudaMalloc(&d_mats[m], dims[m] * (size_t)ldm * 8) == cudaSuccess &&
         cudaMemsetAsync(d_mats[m], 0, dims[m] * (size_t)ldm * 8, stream) == cudaSuccess &&
         h2d(m) == cudaSuccess;
  }
  ok = ok && cudaMalloc(&d_out, maxdim * (size_t)ldm * 8) == cudaSuccess;
  ok = ok && cudaMallocHost(&m1, maxdim * (size_t)R * 8) == cudaSuccess;
  double fit = 0, oldfit = 0;
  if (ok) {
    std::vector<std::vector<double>> ata(N, std::vector<double>((size_t)R * R));
    for (int m = 0; m < N; ++m) gram(mats[m], dims[m], R, ata[m].data());
    std::vector<double> neq((size_t)R * R);
    const double ttnormsq = csf_frobsq(tensors);
    const uint64_t niters = (uint64_t)options[SPLATT_OPTION_NITER];
    for (uint64_t it = 0; it < niters && ok; ++it) {
      auto t0 = std::chrono::steady_clock::now();
      for (int m = 0; m < N && ok; ++m) {
        // M1 = X_(m) (khatri-rao of the other factors), on the GPU
        rc = splatt_b200_mttkrp(T, m, R, ldm, d_mats, d_out, stream);
        cudaError_t e = cudaSuccess;
        if (ldm == R)
          e = cudaMemcpyAsync(m1, d_out, dims[m] * (size_t)R * 8, cudaMemcpyDeviceToHost, stream);
        else
          e = cudaMemcpy2DAsync(m1, (size_t)R * 8, d_out, (size_t)ldm * 8, (size_t)R * 8, dims[m],
                                cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (rc != SPLATT_SUCCESS || e != cudaSuccess) { ok = false; break; }
        // A_m = M1 * (hadamard Grams)^-1   (src/cpd.c:337-339, src/matrix.c:529-606)
        memcpy(mats[m], m1, dims[m] * (size_t)R * 8);
        form_normal_matrix(ata, m, N, R, neq.data());
        std::vector<double> chol(neq);
        if (cholesky(chol.data(), R)) {
          cholesky_solve_rows(chol.data(), R, mats[m], dims[m]);
        } else {
          fprintf(stderr, "SPLATT: Gram matrix is not SPD. Trying pseudo-inverse.\n");
          pinv_solve_rows(neq.data(), R, mats[m], dims[m]);
        }
        normalize_cols(mats[m], dims[m], R, lambda, it == 0);   // src/cpd.c:343-347
        gram(mats[m], dims[m], R, ata[m].data());               // src/cpd.c:350
        ok = h2d(m) == cudaSuccess;                             // keep the device copy current
      }
      if (!ok) break;
      // fit (src/cpd.c:237-265): uses the last mode's M1
      const double norm_mats = kruskal_norm(ata, lambda, N, R);
      const double inner = kruskal_inner(mats[N - 1], m1, dims[N - 1], R, lambda);
      double residual = ttnormsq + norm_mats - 2 * inner;
      if (residual > 0.) residual = std::sqrt(residual);
      fit = 1 - residual / std::sqrt(ttnormsq);
      const double secs =
          std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (verbosity > SPLATT_VERBOSITY_NONE)
        printf("  its = %3llu (%0.3fs)  fit = %0.5f  delta = %+0.4e\n",
               (unsigned long long)it + 1, secs, fit, fit - oldfit);
      if (fit == 1. || (it > 0 && std::fabs(fit - oldfit) < options[SPLATT_OPTION_TOLERANCE]))
        break;
      oldfit = fit;
    }
    // post-process (src/cpd.c:391-411): 2-normalise every factor into lambda
    if (ok) {
      std::vector<double> tmp(R);
      for (int m = 0; m < N; ++m) {
        normalize_cols(mats[m], dims[m], R, tmp.data(), true);
        for (int f = 0; f < R; ++f) lambda[f] *= tmp[f];
      }
    }
  }
  if (stream) cudaStreamSynchronize(stream);
  for (int m = 0; m < N; ++m) if (d_mats[m]) cudaFree(d_mats[m]);
  if (d_out) cudaFree(d_out);
  if (m1) cudaFreeHost(m1);
  if (stream) cudaStreamDestroy(stream);
  splatt_b200_tensor_free(T);
  if (!ok) {
    fprintf(stderr, "SPLATT: CPD-ALS failed (%s)\n", cudaGetErrorString(cudaGetLastError()));
    for (int m = 0; m < N; ++m) free(mats[m]);
    free(lambda);
    return SPLATT_ERROR_NOMEMORY;
  }
  factored->fit = fit;
  factored->rank = nfactors;
  factored->nmodes = N;
  factored->lambda = lambda;
  for (int m = 0; m < N; ++m) {
    factored->dims[m] = dims[m];
    factored->factors[m] = mats[m];
  }
  return SPLATT_SUCCESS;
}

void splatt_free_kruskal(splatt_kruskal * factored) {
  if (!factored) return;
  free(factored->lambda);
  for (splatt_idx_t m = 0; m < factored->nmodes; ++m) free(factored->factors[m]);
}

}  // extern "C"
