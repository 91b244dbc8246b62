/*
 * restate.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's algorithms on the MTTKRP / CPD-ALS hot
 * path, written for clarity, not speed.  Every function cites the reference
 * lines (ShadenSmith/splatt) whose behaviour it restates.  Pinned by
 * tests/test_oracle.py against (a) the compiled reference in oracle/_ref and
 * (b) the golden vectors in tests/golden/ that were generated from the
 * reference's own fixtures by tests/golden/make_golden.py.
 *
 * Types follow the reference's default build: idx = uint64, val = double.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 8
typedef uint64_t idx_t;
typedef double   val_t;

/* ------------------------------------------------------------------------- */
/* (1) COO streaming MTTKRP -- the reference's gold.                          */
/*     follows mttkrp_stream, src/mttkrp.c:1697-1757                          */
/* ------------------------------------------------------------------------- */
void oracle_mttkrp_coo(idx_t nmodes, idx_t const * dims, idx_t nnz, idx_t const * const * ind,
                       val_t const * vals, idx_t R, val_t const * const * mats, idx_t mode,
                       val_t * out)
{
  memset(out, 0, dims[mode] * R * sizeof(val_t));            /* :1710 */
  val_t * accum = malloc(R * sizeof(val_t));
  for(idx_t n=0; n < nnz; ++n) {
    for(idx_t f=0; f < R; ++f) accum[f] = vals[n];           /* :1730-1732 */
    for(idx_t m=0; m < nmodes; ++m) {
      if(m == mode) continue;
      val_t const * row = mats[m] + ind[m][n] * R;           /* :1738-1742 */
      for(idx_t f=0; f < R; ++f) accum[f] *= row[f];
    }
    val_t * orow = out + ind[mode][n] * R;                   /* :1746-1751 */
    for(idx_t f=0; f < R; ++f) orow[f] += accum[f];
  }
  free(accum);
}

/* ------------------------------------------------------------------------- */
/* (2) Level orders.  follows csf_find_mode_order, src/csf.c:694-726          */
/*     which: 0 SORTED_SMALLFIRST (p_order_dims_small :111-136)               */
/*            1 SORTED_BIGFIRST   (p_order_dims_large :204-238)               */
/*            2 INORDER_MINUSONE  (p_order_dims_inorder :147-166)             */
/*            3 SORTED_MINUSONE   (p_order_dims_minusone :177-195)            */
/* ------------------------------------------------------------------------- */
static void order_sorted(idx_t const * dims, idx_t n, int big_first, idx_t * perm)
{
  idx_t sorted[MAXN]; int matched[MAXN];
  for(idx_t m=0; m < n; ++m) { sorted[m] = dims[m]; matched[m] = 0; }
  for(idx_t i=1; i < n; ++i) {              /* insertion sort ascending */
    idx_t v = sorted[i]; idx_t j = i;
    while(j > 0 && sorted[j-1] > v) { sorted[j] = sorted[j-1]; --j; }
    sorted[j] = v;
  }
  if(big_first) for(idx_t m=0; m < n/2; ++m) { idx_t t = sorted[n-m-1]; sorted[n-m-1] = sorted[m]; sorted[m] = t; }
  for(idx_t mfind=0; mfind < n; ++mfind)    /* first unmatched mode of that length */
    for(idx_t mcheck=0; mcheck < n; ++mcheck)
      if(sorted[mfind] == dims[mcheck] && !matched[mcheck]) { perm[mfind] = mcheck; matched[mcheck] = 1; break; }
}
static void to_front(idx_t * perm, idx_t n, idx_t mode)
{
  for(idx_t m=0; m < n; ++m)
    if(perm[m] == mode) { memmove(perm + 1, perm, m * sizeof(idx_t)); perm[0] = mode; break; }
}
void oracle_mode_order(idx_t const * dims, idx_t nmodes, int which, idx_t mode, idx_t * perm)
{
  switch(which) {
  case 0: order_sorted(dims, nmodes, 0, perm); break;
  case 1: order_sorted(dims, nmodes, 1, perm); break;
  case 2: for(idx_t m=0; m < nmodes; ++m) perm[m] = m; to_front(perm, nmodes, mode); break;
  case 3: order_sorted(dims, nmodes, 0, perm); to_front(perm, nmodes, mode); break;
  default: break;
  }
}

/* csf_alloc policies, src/csf.c:770-814; mode -> CSF map, src/mttkrp.c:1832-1861.
 * alloc: 0 ONEMODE, 1 TWOMODE, 2 ALLMODE.  perms is ncsf x MAXN.  Returns ncsf. */
idx_t oracle_csf_policy(idx_t const * dims, idx_t nmodes, int alloc, idx_t * perms, idx_t * map)
{
  idx_t ncsf = 0;
  if(alloc == 0) { oracle_mode_order(dims, nmodes, 0, 0, perms); ncsf = 1; }
  else if(alloc == 1) {
    oracle_mode_order(dims, nmodes, 0, 0, perms);
    oracle_mode_order(dims, nmodes, 3, perms[nmodes-1], perms + MAXN);
    ncsf = 2;
  } else if(alloc == 2) {
    for(idx_t m=0; m < nmodes; ++m) oracle_mode_order(dims, nmodes, 3, m, perms + m*MAXN);
    ncsf = nmodes;
  }
  for(idx_t m=0; m < nmodes; ++m) {
    if(alloc == 0) map[m] = 0;
    else if(alloc == 1) map[m] = (perms[nmodes-1] == m) ? 1 : 0;
    else map[m] = m;
  }
  return ncsf;
}

/* ------------------------------------------------------------------------- */
/* (3) Untiled CSF construction.                                              */
/*     follows p_csf_alloc_untiled src/csf.c:468-502 (sort by the level order */
/*     :475 / src/sort.c:912-918), p_mk_outerptr :248-341, p_mk_fptr :356-458 */
/* ------------------------------------------------------------------------- */
typedef struct {
  idx_t nnz, nmodes;
  idx_t dims[MAXN], dim_perm[MAXN], dim_iperm[MAXN];
  idx_t nfibs[MAXN];
  idx_t * fptr[MAXN];
  idx_t * fids[MAXN];   /* fids[0] == NULL iff no empty root slices (:303-309) */
  val_t * vals;
} oracle_csf;

static idx_t const * const * g_sort_ind; static idx_t const * g_sort_perm; static idx_t g_sort_n;
static int cmp_nnz(void const * a, void const * b)
{
  idx_t const x = *(idx_t const *)a, y = *(idx_t const *)b;
  for(idx_t l=0; l < g_sort_n; ++l) {
    idx_t const ix = g_sort_ind[g_sort_perm[l]][x], iy = g_sort_ind[g_sort_perm[l]][y];
    if(ix != iy) return ix < iy ? -1 : 1;
  }
  return x < y ? -1 : (x > y);   /* stable */
}

oracle_csf * oracle_csf_build(idx_t nmodes, idx_t const * dims, idx_t nnz, idx_t const * const * ind,
                              val_t const * vals, idx_t const * perm)
{
  oracle_csf * c = calloc(1, sizeof(*c));
  c->nnz = nnz; c->nmodes = nmodes;
  for(idx_t m=0; m < nmodes; ++m) { c->dims[m] = dims[m]; c->dim_perm[m] = perm[m]; c->dim_iperm[perm[m]] = m; }

  idx_t * order = malloc((nnz + 1) * sizeof(idx_t));
  for(idx_t n=0; n < nnz; ++n) order[n] = n;
  g_sort_ind = ind; g_sort_perm = perm; g_sort_n = nmodes;
  qsort(order, nnz, sizeof(idx_t), cmp_nnz);

  /* leaf level: sorted nonzeros (:486-491) */
  c->nfibs[nmodes-1] = nnz;
  c->fids[nmodes-1] = malloc((nnz + 1) * sizeof(idx_t));
  c->vals = malloc((nnz + 1) * sizeof(val_t));
  for(idx_t n=0; n < nnz; ++n) { c->fids[nmodes-1][n] = ind[perm[nmodes-1]][order[n]]; c->vals[n] = vals[order[n]]; }

  /* a node starts at level l wherever any index at levels 0..l changes */
  for(idx_t l=0; l + 1 < nmodes; ++l) {
    idx_t nf = 0;
    for(idx_t n=0; n < nnz; ++n) {
      int start = (n == 0);
      for(idx_t u=0; u <= l && !start; ++u) start = ind[perm[u]][order[n]] != ind[perm[u]][order[n-1]];
      nf += start;
    }
    c->nfibs[l] = nf;
    c->fids[l] = malloc((nf + 1) * sizeof(idx_t));
    c->fptr[l] = malloc((nf + 2) * sizeof(idx_t));
  }
  /* fill ids and child pointers: fptr[l][f] = first child (a node of level l+1, or a nonzero) */
  idx_t cnt[MAXN] = {0};
  for(idx_t n=0; n < nnz; ++n) {
    idx_t first = nmodes;   /* first level whose index differs from the previous nonzero */
    if(n == 0) first = 0;
    else for(idx_t u=0; u < nmodes; ++u) if(ind[perm[u]][order[n]] != ind[perm[u]][order[n-1]]) { first = u; break; }
    for(idx_t l=first; l + 1 < nmodes; ++l) {
      c->fids[l][cnt[l]] = ind[perm[l]][order[n]];
      c->fptr[l][cnt[l]] = (l + 2 == nmodes) ? n : cnt[l+1];
      ++cnt[l];
    }
  }
  for(idx_t l=0; l + 1 < nmodes; ++l) c->fptr[l][c->nfibs[l]] = (l + 2 == nmodes) ? nnz : c->nfibs[l+1];
  if(c->nfibs[0] == dims[perm[0]]) { free(c->fids[0]); c->fids[0] = NULL; }   /* :303-309 */
  free(order);
  return c;
}

void oracle_csf_free(oracle_csf * c)
{
  if(!c) return;
  for(idx_t l=0; l < MAXN; ++l) { free(c->fptr[l]); free(c->fids[l]); }
  free(c->vals); free(c);
}
void oracle_csf_get(oracle_csf const * c, idx_t * nfibs, idx_t * perm, idx_t ** fptr, idx_t ** fids, val_t ** vals)
{
  for(idx_t l=0; l < c->nmodes; ++l) { nfibs[l] = c->nfibs[l]; perm[l] = c->dim_perm[l]; fptr[l] = c->fptr[l]; fids[l] = c->fids[l]; }
  *vals = c->vals;
}

/* ------------------------------------------------------------------------- */
/* (4) CSF MTTKRP for an output mode at any depth.                            */
/*     root  : p_csf_mttkrp_root3_* src/mttkrp.c:390-541, root_* :668-799,   */
/*             p_propagate_up :324-387                                        */
/*     intl  : intl3_* :544-607, intl_* :1096-1278                            */
/*     leaf  : leaf3_* :610-665, leaf_* :860-1029                             */
/*     Written as the recursion those iterative DFS loops implement:          */
/*       up(l,f)   = sum over children g of  U_{l+1}[id(g)] (*) up(l+1,g)     */
/*                   (for l = N-2:  sum over nonzeros  val * U_{N-1}[k])      */
/*       output d  : out[id(g)] += (prefix of levels < d) (*) up(d,g)         */
/* ------------------------------------------------------------------------- */
typedef struct { oracle_csf const * c; idx_t R; val_t const * const * U; /* by level */ val_t * out; idx_t d; } walk_t;

static void subtree_up(walk_t const * w, idx_t l, idx_t f, val_t * z)
{
  oracle_csf const * c = w->c; idx_t const R = w->R, N = c->nmodes;
  for(idx_t r=0; r < R; ++r) z[r] = 0;
  if(l == N - 2) {
    for(idx_t n=c->fptr[l][f]; n < c->fptr[l][f+1]; ++n) {          /* p_csf_process_fiber :304-321 */
      val_t const * row = w->U[N-1] + c->fids[N-1][n] * R;
      for(idx_t r=0; r < R; ++r) z[r] += c->vals[n] * row[r];
    }
    return;
  }
  val_t * zc = malloc(R * sizeof(val_t));
  for(idx_t g=c->fptr[l][f]; g < c->fptr[l][f+1]; ++g) {
    subtree_up(w, l+1, g, zc);
    val_t const * row = w->U[l+1] + c->fids[l+1][g] * R;              /* p_add_hada_clear :240-250 */
    for(idx_t r=0; r < R; ++r) z[r] += zc[r] * row[r];
  }
  free(zc);
}

static void walk_down(walk_t const * w, idx_t l, idx_t f, val_t const * prefix /* levels < l */)
{
  oracle_csf const * c = w->c; idx_t const R = w->R, N = c->nmodes;
  idx_t const id = (l == 0 && c->fids[0] == NULL) ? f : c->fids[l][f];
  if(l == w->d) {                                   /* output level */
    val_t * z = malloc(R * sizeof(val_t));
    subtree_up(w, l, f, z);
    val_t * orow = w->out + id * R;
    for(idx_t r=0; r < R; ++r) orow[r] += (prefix ? prefix[r] : 1.0) * z[r];
    free(z);
    return;
  }
  val_t * p = malloc(R * sizeof(val_t));              /* p_assign_hada :253-262 */
  val_t const * row = w->U[l] + id * R;
  for(idx_t r=0; r < R; ++r) p[r] = (prefix ? prefix[r] : 1.0) * row[r];
  if(l == N - 2) {                                  /* output is the leaf level */
    for(idx_t n=c->fptr[l][f]; n < c->fptr[l][f+1]; ++n) {          /* p_csf_process_fiber_nolock :283-301 */
      val_t * orow = w->out + c->fids[N-1][n] * R;
      for(idx_t r=0; r < R; ++r) orow[r] += c->vals[n] * p[r];
    }
  } else {
    for(idx_t g=c->fptr[l][f]; g < c->fptr[l][f+1]; ++g) walk_down(w, l+1, g, p);
  }
  free(p);
}

/* mats indexed by MODE (like the API); out is dims[mode] x R, zeroed here (:1303-1305). */
void oracle_mttkrp_csf(oracle_csf const * c, idx_t R, val_t const * const * mats, idx_t mode, val_t * out)
{
  val_t const * U[MAXN];
  for(idx_t l=0; l < c->nmodes; ++l) U[l] = mats[c->dim_perm[l]];
  memset(out, 0, c->dims[mode] * R * sizeof(val_t));
  walk_t w = { c, R, U, out, c->dim_iperm[mode] };
  for(idx_t s=0; s < c->nfibs[0]; ++s) walk_down(&w, 0, s, NULL);
}

/* ------------------------------------------------------------------------- */
/* (5) CPD-ALS.  follows cpd_als_iterate src/cpd.c:271-387 with               */
/*     mat_aTa src/matrix.c:414-455, mat_solve_normals :529-606 (Cholesky     */
/*     branch; note p_form_gram :45-51 overwrites the `1+reg` diagonal, so    */
/*     the regulariser has no effect), mat_normalize :501-525 (p_mat_2norm    */
/*     :86-144, p_mat_maxnorm :147-199), p_calc_fit src/cpd.c:237-265,        */
/*     p_kruskal_norm :116-152, p_tt_kruskal_inner :171-218,                  */
/*     cpd_post_process :391-411, rand_val src/util.c:15-23.                  */
/*     MTTKRP is the COO gold of (1).                                         */
/* ------------------------------------------------------------------------- */
static val_t rand_val(void) { val_t v = 3.0 * ((val_t) rand() / (val_t) RAND_MAX); if(rand() % 2 == 0) v *= -1; return v; }

static void gram(val_t const * A, idx_t I, idx_t R, val_t * G)
{
  memset(G, 0, R * R * sizeof(val_t));
  for(idx_t i=0; i < I; ++i) for(idx_t p=0; p < R; ++p) for(idx_t q=p; q < R; ++q) G[q + p*R] += A[p + i*R] * A[q + i*R];
}
static void normalize(val_t * A, idx_t I, idx_t R, val_t * lambda, int two)
{
  for(idx_t j=0; j < R; ++j) lambda[j] = 0;
  for(idx_t i=0; i < I; ++i) for(idx_t j=0; j < R; ++j) {
    val_t const a = A[j + i*R];
    if(two) lambda[j] += a * a; else if(a > lambda[j]) lambda[j] = a;
  }
  for(idx_t j=0; j < R; ++j) lambda[j] = two ? sqrt(lambda[j]) : (lambda[j] > 1. ? lambda[j] : 1.);
  for(idx_t i=0; i < I; ++i) for(idx_t j=0; j < R; ++j) A[j + i*R] /= lambda[j];
}

double oracle_cpd_als(idx_t nmodes, idx_t const * dims, idx_t nnz, idx_t const * const * ind,
                      val_t const * vals, idx_t R, idx_t niters, double tol, unsigned seed,
                      val_t ** factors, val_t * lambda)
{
  srand(seed);
  idx_t maxdim = 0;
  for(idx_t m=0; m < nmodes; ++m) { if(dims[m] > maxdim) maxdim = dims[m]; for(idx_t x=0; x < dims[m]*R; ++x) factors[m][x] = rand_val(); }
  val_t * m1 = malloc(maxdim * R * sizeof(val_t));
  val_t * ata[MAXN]; val_t * neq = malloc(R * R * sizeof(val_t));
  for(idx_t m=0; m < nmodes; ++m) { ata[m] = malloc(R * R * sizeof(val_t)); gram(factors[m], dims[m], R, ata[m]); }
  double ttnormsq = 0; for(idx_t n=0; n < nnz; ++n) ttnormsq += vals[n] * vals[n];
  double fit = 0, oldfit = 0;
  for(idx_t it=0; it < niters; ++it) {
    for(idx_t m=0; m < nmodes; ++m) {
      oracle_mttkrp_coo(nmodes, dims, nnz, ind, vals, R, (val_t const * const *) factors, m, m1);
      memcpy(factors[m], m1, dims[m] * R * sizeof(val_t));
      for(idx_t x=0; x < R*R; ++x) neq[x] = 1.;
      for(idx_t o=0; o < nmodes; ++o) if(o != m) for(idx_t i=0; i < R; ++i) for(idx_t j=i; j < R; ++j) neq[j + i*R] *= ata[o][j + i*R];
      for(idx_t i=0; i < R; ++i) for(idx_t j=0; j < i; ++j) neq[j + i*R] = neq[i + j*R];
      /* Cholesky (lower, row-major) + two triangular solves per row */
      for(idx_t j=0; j < R; ++j) {
        val_t d = neq[j + j*R];
        for(idx_t k=0; k < j; ++k) d -= neq[k + j*R] * neq[k + j*R];
        d = sqrt(d); neq[j + j*R] = d;
        for(idx_t i=j+1; i < R; ++i) { val_t s = neq[j + i*R]; for(idx_t k=0; k < j; ++k) s -= neq[k + i*R] * neq[k + j*R]; neq[j + i*R] = s / d; }
      }
      for(idx_t i=0; i < dims[m]; ++i) {
        val_t * x = factors[m] + i*R;
        for(idx_t p=0; p < R; ++p) { val_t s = x[p]; for(idx_t k=0; k < p; ++k) s -= neq[k + p*R] * x[k]; x[p] = s / neq[p + p*R]; }
        for(idx_t pp=R; pp > 0; --pp) { idx_t p = pp - 1; val_t s = x[p]; for(idx_t k=p+1; k < R; ++k) s -= neq[p + k*R] * x[k]; x[p] = s / neq[p + p*R]; }
      }
      normalize(factors[m], dims[m], R, lambda, it == 0);
      gram(factors[m], dims[m], R, ata[m]);
    }
    /* fit */
    double norm_mats = 0;
    for(idx_t i=0; i < R; ++i) for(idx_t j=i; j < R; ++j) {
      val_t a = 1.; for(idx_t m=0; m < nmodes; ++m) a *= ata[m][j + i*R];
      norm_mats += a * lambda[i] * lambda[j] * (i == j ? 1. : 2.);
    }
    norm_mats = fabs(norm_mats);
    double inner = 0;
    for(idx_t r=0; r < R; ++r) { double a = 0; for(idx_t i=0; i < dims[nmodes-1]; ++i) a += factors[nmodes-1][r + i*R] * m1[r + i*R]; inner += a * lambda[r]; }
    double residual = ttnormsq + norm_mats - 2 * inner;
    if(residual > 0.) residual = sqrt(residual);
    fit = 1 - residual / sqrt(ttnormsq);
    if(fit == 1. || (it > 0 && fabs(fit - oldfit) < tol)) break;
    oldfit = fit;
  }
  val_t * tmp = malloc(R * sizeof(val_t));
  for(idx_t m=0; m < nmodes; ++m) { normalize(factors[m], dims[m], R, tmp, 1); for(idx_t f=0; f < R; ++f) lambda[f] *= tmp[f]; }
  free(tmp); free(m1); free(neq);
  for(idx_t m=0; m < nmodes; ++m) free(ata[m]);
  return fit;
}
