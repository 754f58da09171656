/* ORACLE (test infrastructure, not product code): plain-C restatement of the
 * interior-point method in oracle/ipm_ref.py (IPOPT semantics, Waechter &
 * Biegler 2006, with the reference's options problem.py:57-60), operating on
 * the lowered tables (include/omg_b200.h: omg_tables).  It is the CPU baseline
 * bench.py times next to the GPU (cpu_baseline.kind = "port") and the checker
 * of the GPU parity tests at sizes where the numpy twin is too slow.
 * parity unpinned: no CasADi/IPOPT binary exists in this image.
 *
 * Build: make -C oracle   ->  oracle/_build/libipm_oracle.so
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may load it.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>
#include "../include/omg_b200.h"

#define KAPPA_EPS 10.0
#define KAPPA_MU 0.2
#define THETA_MU 1.5
#define TAU_MIN 0.99
#define S_MAX 100.0
#define KAPPA_SIGMA 1e10
#define GAMMA_THETA 1e-5
#define GAMMA_PHI 1e-8
#define ETA_PHI 1e-8
#define S_THETA 1.1
#define S_PHI 2.3
#define DELTA_LS 1.0
#define GAMMA_ALPHA 0.05
#define THETA_MAX_FACT 1e4
#define THETA_MIN_FACT 1e-4
#define DELTA_W0 1e-4
#define DELTA_W_MIN 1e-20
#define DELTA_W_MAX 1e40
#define KAPPA_W_PLUS_FIRST 100.0
#define KAPPA_W_PLUS 8.0
#define KAPPA_W_MINUS (1.0 / 3.0)
#define DELTA_C_VAL 1e-8
#define DELTA_C_EXP 0.25
#define PIV_TOL 1e-12
#define INF_BOUND 1e19
#define MAX_LS 40
#define MAXF 32
#define SOFT_RESTO_FACTOR 0.9999

static double eval_slot(const omg_termlist* L, int s, const double* V, const double* xe) {
  double acc = 0.0;
  for (int t = L->ptr[s]; t < L->ptr[s + 1]; ++t) {
    double v = L->coef[t] * V[L->cidx[t]];
    for (int k = 0; k < L->width; ++k) v *= xe[L->xi[t * L->width + k]];
    acc += v;
  }
  return acc;
}

/* intermediates (include/omg_b200.h): x_ext = [x, 1, mids] */
static void eval_mids(const omg_tables* T, const double* V, double* xe) {
  for (int l = 0; l < T->n_mid; ++l) xe[T->n + 1 + l] = eval_slot(&T->G, T->m + l, V, xe);
}

/* raw A = d row/d mid and C = d mid/d x slots of the J term list */
static void eval_jx(const omg_tables* T, const double* V, const double* xe, double* jx) {
  if (!T->n_mid) return;
  for (int s = T->nnz_j; s < T->nnz_jx; ++s) jx[s] = eval_slot(&T->J, s, V, xe);
}

/* constraint Jacobian slot: direct part + chain rule through the mids */
static double jac_slot(const omg_tables* T, int s, const double* V, const double* xe, const double* jx) {
  double v = eval_slot(&T->J, s, V, xe);
  if (T->n_mid)
    for (int e = T->jp_ptr[s]; e < T->jp_ptr[s + 1]; ++e) v += jx[T->jp_a[e]] * jx[T->jp_c[e]];
  return v;
}

static void eval_tape(const omg_tables* T, const double* p, double* V) {
  V[0] = 1.0;
  for (int i = 0; i < T->n_par; ++i) V[1 + i] = p[i];
  for (int e = 0; e < T->n_tape; ++e) {
    double acc = 0.0;
    for (int t = T->tape_ptr[e]; t < T->tape_ptr[e + 1]; ++t) {
      const int32_t* f = T->tape_fac + 4 * t;
      acc += T->tape_coef[t] * V[f[0]] * V[f[1]] * V[f[2]] * V[f[3]];
    }
    switch (T->tape_func[e]) {
      case 1: acc = 1.0 / acc; break;
      case 2: acc = (acc >= 0.0) ? 1.0 : 0.0; break;
      case 3: acc = (acc > 0.0) ? 1.0 : 0.0; break;
      case 4: acc = sin(acc); break;
      case 5: acc = cos(acc); break;
      case 6: acc = sqrt(acc); break;
      default: break;
    }
    V[1 + T->n_par + e] = acc;
  }
}

static int cmp_le(double lhs, double rhs, double base) {
  return lhs - rhs <= 10.0 * DBL_EPSILON * fabs(base);
}

/* K = L S L^T, S = diag(sign) in the permuted order of
 * lowering.build_kkt_structure; row-major full storage, lower part used. */
/* K = L S L^T in the permuted order.  mode 0 (IPOPT's inertia test): S[j] = sign of pivot
 * j as it comes; accepted when the number of negative pivots equals n_neg (Sylvester's law
 * of inertia).  mode 1: S fixed to `sign` (+1 variables, -1 equality rows), a pivot of the
 * other sign fails.  sdyn receives S.  eq_fail: too few negative pivots / failure at an
 * equality pivot -> the caller regularises the constraint block (delta_c). */
static int signed_cholesky(double* K, int N, const int32_t* sign, int n_neg, int mode, double* d0,
                           double* sdyn, int* eq_fail) {
  *eq_fail = 0;
  int neg = 0;
  for (int j = 0; j < N; ++j) d0[j] = fabs(K[j * N + j]);
  for (int j = 0; j < N; ++j) {
    const double a = K[j * N + j];
    const double sgn = mode ? (double)sign[j] : (a > 0.0 ? 1.0 : -1.0);
    const double piv = sgn * a;
    const double thr = (mode && sgn < 0) ? 0.0 : PIV_TOL * fmax(d0[j], 1e-300);
    if (!(piv > thr) || !isfinite(piv)) { *eq_fail = (sign[j] < 0); return 0; }
    sdyn[j] = sgn; neg += (sgn < 0);
    const double ljj = sqrt(piv);
    K[j * N + j] = ljj;
    for (int i = j + 1; i < N; ++i) K[i * N + j] /= (sgn * ljj);
    for (int i = j + 1; i < N; ++i) {
      const double lij = sgn * K[i * N + j];
      if (lij == 0.0) continue;
      for (int k = j + 1; k <= i; ++k) K[i * N + k] -= lij * K[k * N + j];
    }
  }
  if (neg != n_neg) { *eq_fail = (neg < n_neg); return 0; }
  return 1;
}

static void signed_solve(const double* L, int N, const double* sdyn, double* w) {
  for (int j = 0; j < N; ++j) {
    w[j] /= L[j * N + j];
    for (int i = j + 1; i < N; ++i) w[i] -= L[i * N + j] * w[j];
  }
  for (int j = 0; j < N; ++j) w[j] *= sdyn[j];
  for (int j = N - 1; j >= 0; --j) {
    w[j] /= L[j * N + j];
    for (int i = 0; i < j; ++i) w[i] -= L[j * N + i] * w[j];
  }
}


/* ---------------------------------------------------------------------------
 * Sparse alternative of the linear solver (linear_solver = 1, oracle_set_linear_solver):
 * K = L D L^T on the fill-reducing structure the product's sparse kernel uses -- constrained
 * minimum-degree ordering (an equality row is eligible once all variables it couples are
 * eliminated), symbolic factorisation once per table set, then the classical up-looking
 * numeric factorisation (elimination-tree reach per row).  Same inertia / pivot rules as
 * signed_cholesky.  It exists so that the CPU arm of bench.py runs the SAME algorithmic
 * work as the GPU arm (the dense factorisation above stays the checker's default).
 * ------------------------------------------------------------------------- */
typedef struct {
  int N, nnzA, nnzL;
  int *pos;                 /* natural node (var j / n + eq k) -> permuted index */
  int *Ap, *Ai;             /* upper triangle of permuted K by columns (row <= col), diagonal last */
  int *hslot, *jslot, *dslot;   /* Ax slot of H position q / border slot s / diagonal j (permuted) */
  int *Lp, *Li, *parent;    /* pattern of L (strictly lower, by columns), elimination tree */
  int *sign;                /* +1 / -1 by permuted index */
  int n_free;               /* leading permuted columns that are variables no equality row touches: a negative
                             * pivot there already decides the inertia test (same early rejection as the
                             * product's sparse kernel, csrc/omg_sp.cuh: SP_CHECK) */
} SpSym;

static int g_linear_solver = 0;
void oracle_set_linear_solver(int kind) { g_linear_solver = kind; }

static void spsym_free(SpSym* S) {
  if (!S) return;
  void* all[] = {S->pos, S->Ap, S->Ai, S->hslot, S->jslot, S->dslot, S->Lp, S->Li, S->parent, S->sign};
  for (unsigned k = 0; k < sizeof(all) / sizeof(all[0]); ++k) free(all[k]);
  free(S);
}

static SpSym* spsym_build(const omg_tables* T) {
  const int n = T->n, n_eq = T->kkt_n_eq, N = n + n_eq;
  SpSym* S = (SpSym*)calloc(1, sizeof(SpSym));
  S->N = N;
  char* A = (char*)calloc((size_t)N * N, 1);
  for (int q = 0; q < T->nnz_h; ++q) { const int r = T->hrow[q], c = T->hcol[q]; if (r != c) { A[(size_t)r * N + c] = 1; A[(size_t)c * N + r] = 1; } }
  int* pending = (int*)calloc(n_eq + 1, sizeof(int));
  for (int k = 0; k < n_eq; ++k) {
    const int i = T->kkt_eq_rows[k];
    for (int s = T->jrow_ptr[i]; s < T->jrow_ptr[i + 1]; ++s) {
      const int c = T->jcol[s];
      if (!A[(size_t)(n + k) * N + c]) { A[(size_t)(n + k) * N + c] = 1; A[(size_t)c * N + n + k] = 1; pending[k]++; }
    }
  }
  char* orig = (char*)malloc((size_t)N * N); memcpy(orig, A, (size_t)N * N);
  int* deg = (int*)calloc(N, sizeof(int)); char* gone = (char*)calloc(N, 1);
  for (int a = 0; a < N; ++a) for (int b = 0; b < N; ++b) deg[a] += A[(size_t)a * N + b];
  S->pos = (int*)malloc(sizeof(int) * N);
  int* nb = (int*)malloc(sizeof(int) * N);
  for (int step = 0; step < N; ++step) {
    int v = -1;
    for (int a = 0; a < N; ++a) {
      if (gone[a] || (a >= n && pending[a - n] > 0)) continue;
      if (v < 0 || deg[a] < deg[v]) v = a;
    }
    int cnt = 0;
    for (int b = 0; b < N; ++b) if (A[(size_t)v * N + b] && !gone[b]) nb[cnt++] = b;
    for (int x = 0; x < cnt; ++x) for (int y = x + 1; y < cnt; ++y) {
      const int a = nb[x], b = nb[y];
      if (!A[(size_t)a * N + b]) { A[(size_t)a * N + b] = 1; A[(size_t)b * N + a] = 1; deg[a]++; deg[b]++; }
    }
    for (int x = 0; x < cnt; ++x) deg[nb[x]]--;
    gone[v] = 1;
    if (v < n) for (int k = 0; k < n_eq; ++k) if (orig[(size_t)(n + k) * N + v]) pending[k]--;
    S->pos[v] = step;
  }
  /* upper triangle of the permuted pattern by columns (the diagonal always present, last) */
  int* inv = (int*)malloc(sizeof(int) * N);
  for (int a = 0; a < N; ++a) inv[S->pos[a]] = a;
  S->Ap = (int*)calloc(N + 1, sizeof(int));
  for (int j = 0; j < N; ++j) {
    int c = 1;
    for (int i = 0; i < j; ++i) if (orig[(size_t)inv[i] * N + inv[j]]) ++c;
    S->Ap[j + 1] = S->Ap[j] + c;
  }
  S->nnzA = S->Ap[N];
  S->Ai = (int*)malloc(sizeof(int) * S->nnzA);
  S->dslot = (int*)malloc(sizeof(int) * N);
  for (int j = 0; j < N; ++j) {
    int p = S->Ap[j];
    for (int i = 0; i < j; ++i) if (orig[(size_t)inv[i] * N + inv[j]]) S->Ai[p++] = i;
    S->Ai[p] = j; S->dslot[j] = p;
  }
  #define SLOT_OF(pa, pb, out) { const int lo_ = (pa) < (pb) ? (pa) : (pb), hi_ = (pa) < (pb) ? (pb) : (pa); out = -1; \
    for (int p_ = S->Ap[hi_]; p_ < S->Ap[hi_ + 1]; ++p_) if (S->Ai[p_] == lo_) { out = p_; break; } }
  S->hslot = (int*)malloc(sizeof(int) * (T->nnz_h + 1));
  for (int q = 0; q < T->nnz_h; ++q) SLOT_OF(S->pos[T->hrow[q]], S->pos[T->hcol[q]], S->hslot[q]);
  S->jslot = (int*)malloc(sizeof(int) * (T->nnz_j + 1));
  for (int s = 0; s < T->nnz_j; ++s) S->jslot[s] = -1;
  for (int k = 0; k < n_eq; ++k) {
    const int i = T->kkt_eq_rows[k];
    for (int s = T->jrow_ptr[i]; s < T->jrow_ptr[i + 1]; ++s) SLOT_OF(S->pos[n + k], S->pos[T->jcol[s]], S->jslot[s]);
  }
  #undef SLOT_OF
  S->sign = (int*)malloc(sizeof(int) * N);
  for (int j = 0; j < N; ++j) S->sign[j] = 1;
  for (int k = 0; k < n_eq; ++k) S->sign[S->pos[n + k]] = -1;
  {
    char* touched = (char*)calloc(N, 1);                 /* by permuted index */
    for (int k = 0; k < n_eq; ++k) {
      const int i = T->kkt_eq_rows[k];
      touched[S->pos[n + k]] = 1;
      for (int s = T->jrow_ptr[i]; s < T->jrow_ptr[i + 1]; ++s) touched[S->pos[T->jcol[s]]] = 1;
    }
    S->n_free = 0;
    while (S->n_free < N && !touched[S->n_free]) S->n_free++;
    free(touched);
  }
  /* elimination tree + column counts of L (up-looking symbolic pass) */
  S->parent = (int*)malloc(sizeof(int) * N); S->Lp = (int*)calloc(N + 1, sizeof(int));
  int* flag = (int*)malloc(sizeof(int) * N); int* lnz = (int*)calloc(N, sizeof(int));
  for (int k = 0; k < N; ++k) {
    S->parent[k] = -1; flag[k] = k;
    for (int p = S->Ap[k]; p < S->Ap[k + 1]; ++p)
      for (int i = S->Ai[p]; i < k && flag[i] != k; i = S->parent[i]) {
        if (S->parent[i] == -1) S->parent[i] = k;
        lnz[i]++; flag[i] = k;
      }
  }
  for (int k = 0; k < N; ++k) S->Lp[k + 1] = S->Lp[k] + lnz[k];
  S->nnzL = S->Lp[N];
  S->Li = (int*)malloc(sizeof(int) * (S->nnzL + 1));
  free(A); free(orig); free(pending); free(deg); free(gone); free(nb); free(inv); free(flag); free(lnz);
  return S;
}

/* numeric up-looking L D L^T; returns 1 on success (inertia = n_neg negatives), like signed_cholesky */
static int sparse_ldl(const SpSym* S, const double* Ax, double* Lx, int* Li, double* D, double* Y, int* pattern,
                      int* flag, int* lnz, int n_neg, int mode, int* eq_fail) {
  const int N = S->N;
  int neg = 0;
  *eq_fail = 0;
  for (int k = 0; k < N; ++k) {
    Y[k] = 0.0; flag[k] = k; lnz[k] = 0;
    int top = N;
    for (int p = S->Ap[k]; p < S->Ap[k + 1]; ++p) {
      int i = S->Ai[p];
      Y[i] += Ax[p];
      int len = 0;
      for (; flag[i] != k; i = S->parent[i]) { pattern[len++] = i; flag[i] = k; }
      while (len > 0) pattern[--top] = pattern[--len];
    }
    double dk = Y[k]; Y[k] = 0.0;
    const double d0 = fabs(Ax[S->dslot[k]]);
    for (; top < N; ++top) {
      const int i = pattern[top];
      const double yi = Y[i]; Y[i] = 0.0;
      const int p2 = S->Lp[i] + lnz[i];
      for (int p = S->Lp[i]; p < p2; ++p) Y[Li[p]] -= Lx[p] * yi;
      const double lki = yi / D[i];
      dk -= lki * yi;
      Li[p2] = k; Lx[p2] = lki; lnz[i]++;
    }
    const int isneg = dk < 0.0;
    const double piv = fabs(dk);
    int bad;
    if (mode) bad = (isneg != (S->sign[k] < 0)) || !(piv > ((S->sign[k] < 0) ? 0.0 : PIV_TOL * fmax(d0, 1e-300))) || !isfinite(piv);
    else bad = !(piv > PIV_TOL * fmax(d0, 1e-300)) || !isfinite(piv);
    if (bad) { *eq_fail = (S->sign[k] < 0); return 0; }
    neg += isneg;
    if (mode == 0 && isneg && k < S->n_free) { *eq_fail = 0; return 0; }   /* v^T H v < 0 with J_eq v = 0 */
    D[k] = dk;
  }
  if (mode == 0 && neg != n_neg) { *eq_fail = (neg < n_neg); return 0; }
  return 1;
}

static void sparse_solve(const SpSym* S, const double* Lx, const int* Li, const double* D, const int* lnz, double* w) {
  const int N = S->N;
  for (int j = 0; j < N; ++j) { const double wj = w[j]; for (int p = S->Lp[j]; p < S->Lp[j] + lnz[j]; ++p) w[Li[p]] -= Lx[p] * wj; }
  for (int j = 0; j < N; ++j) w[j] /= D[j];
  for (int j = N - 1; j >= 0; --j) { double acc = w[j]; for (int p = S->Lp[j]; p < S->Lp[j] + lnz[j]; ++p) acc -= Lx[p] * w[Li[p]]; w[j] = acc; }
}

typedef struct {
  double *V, *xe, *xt, *g, *s, *y, *zL, *zU, *dsc, *sL, *sU, *beq, *sig, *wv, *ds, *dy,
         *dzL, *dzU, *gt, *st, *jval, *gf, *K, *rhs, *rx, *d0, *sol;
  int *rt, *eqidx, *eqrow;
  double *jx, *mu, *sol2, *sdyn, *wx;
  /* sparse linear solver (sym != 0) */
  const SpSym* sym; double *Ax, *Lx, *Dd, *Yw; int *Lis, *pat, *flg, *lnz;
} Work;

static void* xalloc(size_t n) { return calloc(n ? n : 1, 1); }

static void work_alloc(Work* w, const omg_tables* T, int Nmax) {
  const int n = T->n, m = T->m;
  w->V = xalloc(sizeof(double) * T->n_v);
  w->xe = xalloc(sizeof(double) * (n + 1 + T->n_mid)); w->xt = xalloc(sizeof(double) * (n + 1 + T->n_mid));
  w->jx = xalloc(sizeof(double) * (T->n_mid ? T->nnz_jx : 1)); w->mu = xalloc(sizeof(double) * (T->n_mid + 1));
  w->wx = xalloc(sizeof(double) * (T->nnz_wx + 1));
  w->sol2 = xalloc(sizeof(double) * (n + 1));
  w->sdyn = xalloc(sizeof(double) * Nmax);
  double** mv[] = {&w->g, &w->s, &w->y, &w->zL, &w->zU, &w->dsc, &w->sL, &w->sU, &w->beq,
                   &w->sig, &w->wv, &w->ds, &w->dy, &w->dzL, &w->dzU, &w->gt, &w->st};
  for (unsigned k = 0; k < sizeof(mv) / sizeof(mv[0]); ++k) *mv[k] = xalloc(sizeof(double) * m);
  w->jval = xalloc(sizeof(double) * T->nnz_j);
  w->gf = xalloc(sizeof(double) * n); w->rx = xalloc(sizeof(double) * n);
  w->K = xalloc(sizeof(double) * Nmax * Nmax); w->rhs = xalloc(sizeof(double) * Nmax);
  w->d0 = xalloc(sizeof(double) * Nmax); w->sol = xalloc(sizeof(double) * Nmax);
  w->rt = xalloc(sizeof(int) * m); w->eqidx = xalloc(sizeof(int) * m);
  w->eqrow = xalloc(sizeof(int) * (m + 1));
  w->sym = 0; w->Ax = w->Lx = w->Dd = w->Yw = 0; w->Lis = w->pat = w->flg = w->lnz = 0;
}

static void work_attach_sparse(Work* w, const SpSym* S) {
  w->sym = S;
  w->Ax = xalloc(sizeof(double) * (S->nnzA + 1)); w->Lx = xalloc(sizeof(double) * (S->nnzL + 1));
  w->Dd = xalloc(sizeof(double) * S->N); w->Yw = xalloc(sizeof(double) * S->N);
  w->Lis = xalloc(sizeof(int) * (S->nnzL + 1)); w->pat = xalloc(sizeof(int) * S->N);
  w->flg = xalloc(sizeof(int) * S->N); w->lnz = xalloc(sizeof(int) * S->N);
}

static void work_free(Work* w) {
  void* all[] = {w->V, w->xe, w->xt, w->g, w->s, w->y, w->zL, w->zU, w->dsc, w->sL, w->sU, w->beq,
                 w->sig, w->wv, w->ds, w->dy, w->dzL, w->dzU, w->gt, w->st, w->jval, w->gf, w->rx,
                 w->K, w->rhs, w->d0, w->sol, w->rt, w->eqidx, w->eqrow, w->jx, w->mu, w->sol2, w->sdyn, w->wx,
                 w->Ax, w->Lx, w->Dd, w->Yw, w->Lis, w->pat, w->flg, w->lnz};
  for (unsigned k = 0; k < sizeof(all) / sizeof(all[0]); ++k) free(all[k]);
}

static void solve_one(const omg_tables* T, const omg_options* O, Work* w, const double* x0,
                      const double* par, const double* lbg, const double* ubg, const double* lam0,
                      double* xout, double* lamout, double* fout, int* status_out, int* iters_out) {
  const int n = T->n, m = T->m;
  double* V = w->V; double* xe = w->xe; double* xt = w->xt;
  double *g = w->g, *s = w->s, *y = w->y, *zL = w->zL, *zU = w->zU, *dsc = w->dsc, *sL = w->sL,
         *sU = w->sU, *beq = w->beq, *sig = w->sig, *wv = w->wv, *ds = w->ds, *dy = w->dy,
         *dzL = w->dzL, *dzU = w->dzU, *gt = w->gt, *st = w->st, *jval = w->jval, *gf = w->gf;
  int *rt = w->rt, *eqidx = w->eqidx, *eqrow = w->eqrow;
  eval_tape(T, par, V);
  for (int i = 0; i < n; ++i) xe[i] = x0[i];
  xe[n] = 1.0; xt[n] = 1.0;
  double* jx = w->jx; double* mu_mid = w->mu;
  eval_mids(T, V, xe); eval_jx(T, V, xe, jx);
  /* scaling */
  const double smg = O->scaling_max_gradient;
  double fmaxv = 0.0;
  for (int j = 0; j < n; ++j) fmaxv = fmax(fmaxv, fabs(eval_slot(&T->DF, j, V, xe)));
  const double fsc = (fmaxv > smg) ? fmax(smg / fmaxv, 1e-8) : 1.0;
  int n_eq = 0, n_bounds = 0;
  for (int i = 0; i < m; ++i) {
    double gm = 0.0;
    for (int sl = T->jrow_ptr[i]; sl < T->jrow_ptr[i + 1]; ++sl)
      gm = fmax(gm, fabs(jac_slot(T, sl, V, xe, jx)));
    const double d = (gm > smg) ? fmax(smg / gm, 1e-8) : 1.0;
    dsc[i] = d;
    const double lb = lbg[i], ub = ubg[i];
    const int eq = (lb == ub);
    const int hL = (lb > -INF_BOUND) && !eq, hU = (ub < INF_BOUND) && !eq;
    rt[i] = (hL ? 1 : 0) | (hU ? 2 : 0) | (eq ? 4 : 0);
    if (eq) { eqrow[n_eq] = i; eqidx[i] = n_eq++; } else eqidx[i] = -1;
    n_bounds += hL + hU;
    double l = lb * d, u = ub * d;
    beq[i] = l;
    if (hL) l -= O->bound_relax_factor * fmax(1.0, fabs(l));
    if (hU) u += O->bound_relax_factor * fmax(1.0, fabs(u));
    sL[i] = l; sU[i] = u;
    const double gi = d * eval_slot(&T->G, i, V, xe);
    g[i] = gi;
    double si = gi;
    double pl = O->bound_push * fmax(1.0, fabs(l)), pu = O->bound_push * fmax(1.0, fabs(u));
    if (hL && hU) { pl = fmin(pl, O->bound_frac * (u - l)); pu = fmin(pu, O->bound_frac * (u - l)); }
    if (hL) si = fmax(si, l + pl);
    if (hU) si = fmin(si, u - pu);
    s[i] = si;
    const double yi = lam0 ? lam0[i] * fsc / d : 0.0;
    y[i] = yi;
    zL[i] = hL ? fmax(O->mult_bound_push, -yi) : 0.0;
    zU[i] = hU ? fmax(O->mult_bound_push, yi) : 0.0;
  }
  int mismatch = (n_eq != T->kkt_n_eq);
  for (int k = 0; k < n_eq && !mismatch; ++k) if (eqrow[k] != T->kkt_eq_rows[k]) mismatch = 1;
  if (mismatch) {   /* equality pattern differs from the lowered structure */
    for (int i = 0; i < n; ++i) xout[i] = x0[i];
    for (int i = 0; i < m; ++i) lamout[i] = 0.0;
    *fout = 0.0; *status_out = OMG_ERROR_IN_STEP_COMPUTATION; *iters_out = 0;
    return;
  }
  const int N = n + n_eq;
  double mu = O->mu_init, tau = fmax(TAU_MIN, 1.0 - mu);
  double theta_max = -1.0, theta_min = -1.0, delta_w_last = 0.0;
  double filt[2 * MAXF]; int nfilt = 0;
  double f = fsc * eval_slot(&T->F, 0, V, xe);
  int status = OMG_MAX_ITER_EXCEEDED, iter = 0, n_restarts = 0;

  for (iter = 0;; ++iter) {
    eval_jx(T, V, xe, jx);
    double cinf = 0, maxprod = 0, minprod = 1e300, viol = 0, rsinf = 0, rsinf_un = 0, ysum = 0,
           zsum = 0, theta = 0, logsum = 0, rxinf = 0;
    for (int i = 0; i < m; ++i) {
      const int r = rt[i]; const double d = dsc[i];
      for (int sl = T->jrow_ptr[i]; sl < T->jrow_ptr[i + 1]; ++sl) jval[sl] = d * jac_slot(T, sl, V, xe, jx);
      const double gi = g[i], si = s[i], yi = y[i];
      const double ci = (r & 4) ? gi - beq[i] : gi - si;
      cinf = fmax(cinf, fabs(ci)); theta += fabs(ci);
      if (r & 1) { const double dl = si - sL[i], pz = dl * zL[i];
        maxprod = fmax(maxprod, pz); minprod = fmin(minprod, pz); logsum += log(dl); zsum += zL[i]; }
      if (r & 2) { const double du = sU[i] - si, pz = du * zU[i];
        maxprod = fmax(maxprod, pz); minprod = fmin(minprod, pz); logsum += log(du); zsum += zU[i]; }
      const double gun = gi / d;
      if (r & 6) viol = fmax(viol, gun - ubg[i]);
      if (r & 5) viol = fmax(viol, lbg[i] - gun);
      if (!(r & 4)) { const double rs = fabs(-yi - zL[i] + zU[i]); rsinf = fmax(rsinf, rs); rsinf_un = fmax(rsinf_un, rs * d); }
      ysum += fabs(yi);
    }
    for (int j = 0; j < n; ++j) gf[j] = fsc * eval_slot(&T->DF, j, V, xe);
    memcpy(w->rx, gf, sizeof(double) * n);
    for (int sl = 0; sl < T->nnz_j; ++sl) w->rx[T->jcol[sl]] += jval[sl] * y[T->jrow[sl]];
    for (int j = 0; j < n; ++j) rxinf = fmax(rxinf, fabs(w->rx[j]));
    const double dinf = fmax(rxinf, rsinf), dinf_un = fmax(rxinf, rsinf_un) / fsc;
    const double s_d = fmax(S_MAX, (ysum + zsum) / fmax(1.0, (double)(m + n_bounds))) / S_MAX;
    const double s_c = fmax(S_MAX, zsum / fmax(1.0, (double)n_bounds)) / S_MAX;
    const double cmpl0 = n_bounds ? fmax(fabs(maxprod), fabs(minprod)) : 0.0;
    const double E0 = fmax(fmax(dinf / s_d, cinf), cmpl0 / s_c);
    if (!isfinite(E0)) { status = OMG_INVALID_NUMBER_DETECTED; break; }
    if (E0 <= O->tol && dinf_un <= O->dual_inf_tol && viol <= O->constr_viol_tol &&
        cmpl0 / fsc <= O->compl_inf_tol) { status = OMG_SOLVE_SUCCEEDED; break; }
    if (iter >= O->max_iter) { status = OMG_MAX_ITER_EXCEEDED; break; }
    const double mu_min = fmin(O->tol, O->compl_inf_tol * fsc) / (KAPPA_EPS + 1.0);
    for (;;) {
      const double cm = n_bounds ? fmax(fabs(maxprod - mu), fabs(minprod - mu)) : 0.0;
      const double Emu = fmax(fmax(dinf / s_d, cinf), cm / s_c);
      if (Emu <= KAPPA_EPS * mu && mu > mu_min) {
        mu = fmax(mu_min, fmin(KAPPA_MU * mu, pow(mu, THETA_MU)));
        tau = fmax(TAU_MIN, 1.0 - mu); nfilt = 0;
      } else break;
    }
    if (theta_max < 0.0) { theta_max = THETA_MAX_FACT * fmax(1.0, theta); theta_min = THETA_MIN_FACT * fmax(1.0, theta); }
    const double phi = f - mu * logsum;
    for (int i = 0; i < m; ++i) {
      const int r = rt[i]; double sg = 0, ph = 0, rd = 0;
      if (!(r & 4)) { const double si = s[i]; rd = g[i] - si;
        if (r & 1) { const double dl = si - sL[i]; sg += zL[i] / dl; ph -= mu / dl; }
        if (r & 2) { const double du = sU[i] - si; sg += zU[i] / du; ph += mu / du; } }
      sig[i] = sg; wv[i] = (r & 4) ? y[i] : (sg * rd + ph);
    }
    /* ---- assemble + factorise -------------------------------------------- */
    double delta_w = 0.0, delta_c = 0.0; int first_try = 1, ok = 0;
    double* K = w->K; double* rhs = w->rhs;
    for (int l = 0; l < T->n_mid; ++l) {   /* multipliers of the mids: mu = A^T lambda */
      double acc = 0.0;
      for (int e = T->mu_ptr[l]; e < T->mu_ptr[l + 1]; ++e) acc += y[T->mu_row[e]] * dsc[T->mu_row[e]] * jx[T->mu_slot[e]];
      mu_mid[l] = acc;
    }
    const SpSym* Sy = w->sym;           /* sparse linear solver: values go to the CSC slots */
#define KH(q) (Sy ? &w->Ax[Sy->hslot[q]] : &K[(size_t)(T->kkt_pos_var[T->hrow[q]] > T->kkt_pos_var[T->hcol[q]] ? T->kkt_pos_var[T->hrow[q]] : T->kkt_pos_var[T->hcol[q]]) * N + \
                                              (T->kkt_pos_var[T->hrow[q]] > T->kkt_pos_var[T->hcol[q]] ? T->kkt_pos_var[T->hcol[q]] : T->kkt_pos_var[T->hrow[q]])])
    for (;;) {
      if (Sy) memset(w->Ax, 0, sizeof(double) * Sy->nnzA); else memset(K, 0, sizeof(double) * N * N);
      for (int q = 0; q < T->nnz_h; ++q) {
        double acc = 0.0;
        for (int e = T->hp_ptr[q]; e < T->hp_ptr[q + 1]; ++e) acc += sig[T->hp_row[e]] * jval[T->hp_s1[e]] * jval[T->hp_s2[e]];
        if (T->hrow[q] == T->hcol[q]) acc += delta_w;
        *KH(q) = acc;
      }
      for (int q = 0; q < T->nnz_w + T->nnz_wx; ++q) {
        double acc = 0.0; const omg_termlist* L = &T->W;
        for (int t = L->ptr[q]; t < L->ptr[q + 1]; ++t) {
          double v = L->coef[t] * V[L->cidx[t]];
          for (int k = 0; k < L->width; ++k) v *= xe[L->xi[t * L->width + k]];
          const int lr = L->lrow[t];
          v *= (lr < m) ? (y[lr] * dsc[lr]) : (lr == m ? fsc : mu_mid[lr - m - 1]);
          acc += v;
        }
        if (q >= T->nnz_w) { w->wx[q - T->nnz_w] = acc; continue; }   /* extra slot X[l,k] / M[l1,l2] */
        const int h = T->w2h[q];
        *KH(h) += acc;
      }
      for (int e = 0; e < (T->nnz_wx ? T->n_xq : 0); ++e) {   /* X^T C + C^T X + C^T M C (include/omg_b200.h) */
        double acc = 0.0;
        for (int r = T->xq_ptr[e]; r < T->xq_ptr[e + 1]; ++r)
          acc += w->wx[T->xq_w[r]] * jx[T->xq_a[r]] * (T->xq_b[r] >= 0 ? jx[T->xq_b[r]] : 1.0);
        const int h = T->xq_h[e];
        *KH(h) += acc;
      }
      for (int k = 0; k < n_eq; ++k) {
        const int i = eqrow[k];
        const int pk = Sy ? Sy->pos[n + k] : T->kkt_pos_eq[k];
        for (int sl = T->jrow_ptr[i]; sl < T->jrow_ptr[i + 1]; ++sl) {
          if (Sy) { w->Ax[Sy->jslot[sl]] = jval[sl]; continue; }
          const int b = T->kkt_pos_var[T->jcol[sl]];
          K[(pk > b ? pk : b) * N + (pk > b ? b : pk)] = jval[sl];
        }
        if (Sy) w->Ax[Sy->dslot[pk]] = -delta_c; else K[pk * N + pk] = -delta_c;
        rhs[pk] = -(g[i] - beq[i]);
      }
#define PV(j) (Sy ? Sy->pos[j] : T->kkt_pos_var[j])
      for (int j = 0; j < n; ++j) rhs[PV(j)] = -gf[j];
      for (int sl = 0; sl < T->nnz_j; ++sl) rhs[PV(T->jcol[sl])] -= jval[sl] * wv[T->jrow[sl]];
      int eq_fail = 0;
      if (Sy) ok = sparse_ldl(Sy, w->Ax, w->Lx, w->Lis, w->Dd, w->Yw, w->pat, w->flg, w->lnz, n_eq, O->inertia_mode, &eq_fail);
      else ok = signed_cholesky(K, N, T->kkt_sign, n_eq, O->inertia_mode, w->d0, w->sdyn, &eq_fail);
      if (ok) break;
      if (eq_fail) delta_c = DELTA_C_VAL * pow(mu, DELTA_C_EXP);
      if (first_try) { delta_w = (delta_w_last == 0.0) ? DELTA_W0 : fmax(DELTA_W_MIN, KAPPA_W_MINUS * delta_w_last); first_try = 0; }
      else delta_w *= (delta_w_last == 0.0) ? KAPPA_W_PLUS_FIRST : KAPPA_W_PLUS;
      if (delta_w > DELTA_W_MAX) break;
    }
    if (!ok) { status = OMG_ERROR_IN_STEP_COMPUTATION; break; }
    if (delta_w > 0.0) delta_w_last = delta_w;
    if (Sy) sparse_solve(Sy, w->Lx, w->Lis, w->Dd, w->lnz, rhs); else signed_solve(K, N, w->sdyn, rhs);
    double* dx = w->sol;   /* natural order: variables, then equality multipliers */
    for (int j = 0; j < n; ++j) dx[j] = rhs[PV(j)];
    for (int k = 0; k < n_eq; ++k) dx[n + k] = rhs[Sy ? Sy->pos[n + k] : T->kkt_pos_eq[k]];
#undef PV
#undef KH
    double a_p = 1.0, a_d = 1.0, gphi = 0.0;
    for (int i = 0; i < m; ++i) {
      const int r = rt[i]; double jd = 0.0;
      for (int sl = T->jrow_ptr[i]; sl < T->jrow_ptr[i + 1]; ++sl) jd += jval[sl] * dx[T->jcol[sl]];
      if (r & 4) { ds[i] = 0; dy[i] = dx[n + eqidx[i]]; dzL[i] = 0; dzU[i] = 0; }
      else {
        const double si = s[i], dsi = jd + (g[i] - si); double ph = 0, a = 0, b = 0;
        if (r & 1) { const double dl = si - sL[i]; ph -= mu / dl; a = mu / dl - zL[i] - (zL[i] / dl) * dsi;
          if (dsi < 0.0) a_p = fmin(a_p, -tau * dl / dsi);
          if (a < 0.0) a_d = fmin(a_d, -tau * zL[i] / a); }
        if (r & 2) { const double du = sU[i] - si; ph += mu / du; b = mu / du - zU[i] + (zU[i] / du) * dsi;
          if (dsi > 0.0) a_p = fmin(a_p, tau * du / dsi);
          if (b < 0.0) a_d = fmin(a_d, -tau * zU[i] / b); }
        ds[i] = dsi; dzL[i] = a; dzU[i] = b; dy[i] = sig[i] * dsi + ph - y[i];
        gphi += ph * dsi;
      }
    }
    for (int j = 0; j < n; ++j) gphi += gf[j] * dx[j];
    double a_min;
    if (gphi < 0.0) { a_min = fmin(GAMMA_THETA, GAMMA_PHI * theta / (-gphi));
      if (theta <= theta_min) a_min = fmin(a_min, DELTA_LS * pow(theta, S_THETA) / pow(-gphi, S_PHI)); }
    else a_min = GAMMA_THETA;
    a_min *= GAMMA_ALPHA;
    double alpha = a_p, ft = 0.0; int accepted = 0, ftype = 0, n_ls = 0;
    while (alpha >= a_min && n_ls < MAX_LS) {
      ++n_ls;
      for (int j = 0; j < n; ++j) xt[j] = xe[j] + alpha * dx[j];
      eval_mids(T, V, xt);
      double tht = 0.0, lg = 0.0;
      for (int i = 0; i < m; ++i) {
        const int r = rt[i]; const double gi = dsc[i] * eval_slot(&T->G, i, V, xt);
        gt[i] = gi;
        if (r & 4) tht += fabs(gi - beq[i]);
        else { const double si = s[i] + alpha * ds[i]; st[i] = si; tht += fabs(gi - si);
          if (r & 1) lg += log(si - sL[i]);
          if (r & 2) lg += log(sU[i] - si); }
      }
      ft = fsc * eval_slot(&T->F, 0, V, xt);
      const double pht = ft - mu * lg;
      int okk = isfinite(pht) && isfinite(tht) && tht <= theta_max;
      if (okk) for (int q = 0; q < nfilt; ++q) if (!(tht < filt[2 * q] || pht < filt[2 * q + 1])) { okk = 0; break; }
      ftype = 0;
      if (okk) {
        const int switching = (theta <= theta_min && gphi < 0.0 && alpha * pow(-gphi, S_PHI) > DELTA_LS * pow(theta, S_THETA));
        if (switching) { okk = cmp_le(pht - phi, ETA_PHI * alpha * gphi, phi); ftype = okk; }
        else okk = cmp_le(tht, (1.0 - GAMMA_THETA) * theta, theta) || cmp_le(pht - phi, -GAMMA_PHI * theta, phi);
      }
      if (okk) { accepted = 1; break; }
      alpha *= 0.5;
    }
    if (!accepted && O->soft_resto) {
      /* soft restoration (ipm_ref.py): accept a step along the same direction that
       * reduces the primal-dual error of the barrier problem */
      double pd0 = 0.0;
      for (int j = 0; j < n; ++j) pd0 += fabs(w->rx[j]);
      for (int i = 0; i < m; ++i) {
        const int r = rt[i];
        if (r & 4) { pd0 += fabs(g[i] - beq[i]); continue; }
        pd0 += fabs(-y[i] - zL[i] + zU[i]) + fabs(g[i] - s[i]);
        if (r & 1) pd0 += fabs((s[i] - sL[i]) * zL[i] - mu);
        if (r & 2) pd0 += fabs((sU[i] - s[i]) * zU[i] - mu);
      }
      alpha = a_p;
      for (int n_try = 0; n_try < 12; ++n_try) {
        for (int j = 0; j < n; ++j) xt[j] = xe[j] + alpha * dx[j];
        eval_mids(T, V, xt); eval_jx(T, V, xt, jx);
        const double az = fmin(alpha, a_d);
        double pdt = 0.0;
        for (int j = 0; j < n; ++j) w->sol2[j] = fsc * eval_slot(&T->DF, j, V, xt);
        for (int i = 0; i < m; ++i) {
          const int r = rt[i];
          const double gi = dsc[i] * eval_slot(&T->G, i, V, xt);
          gt[i] = gi;
          const double yt = y[i] + alpha * dy[i];
          for (int sl = T->jrow_ptr[i]; sl < T->jrow_ptr[i + 1]; ++sl)
            w->sol2[T->jcol[sl]] += dsc[i] * jac_slot(T, sl, V, xt, jx) * yt;
          if (r & 4) { pdt += fabs(gi - beq[i]); continue; }
          const double si = s[i] + alpha * ds[i];
          st[i] = si;
          const double zl = zL[i] + az * dzL[i], zu = zU[i] + az * dzU[i];
          pdt += fabs(-yt - zl + zu) + fabs(gi - si);
          if (r & 1) pdt += fabs((si - sL[i]) * zl - mu);
          if (r & 2) pdt += fabs((sU[i] - si) * zu - mu);
        }
        for (int j = 0; j < n; ++j) pdt += fabs(w->sol2[j]);
        if (isfinite(pdt) && pdt <= SOFT_RESTO_FACTOR * pd0) {
          accepted = 1; ftype = 1; nfilt = 0;
          ft = fsc * eval_slot(&T->F, 0, V, xt);
          break;
        }
        alpha *= 0.5;
      }
    }
    if (!accepted) {
      if (n_restarts < O->max_restarts) {   /* feasibility restart (see ipm_ref.py) */
        ++n_restarts;
        mu = O->restart_mu; tau = fmax(TAU_MIN, 1.0 - mu);
        for (int i = 0; i < m; ++i) {
          const int r = rt[i];
          double si = g[i];
          if (r & 1) si = fmax(si, sL[i] + O->restart_push * fmax(1.0, fabs(sL[i])));
          if (r & 2) si = fmin(si, sU[i] - O->restart_push * fmax(1.0, fabs(sU[i])));
          s[i] = si; y[i] = 0.0;
          zL[i] = (r & 1) ? mu / (si - sL[i]) : 0.0;
          zU[i] = (r & 2) ? mu / (sU[i] - si) : 0.0;
        }
        nfilt = 0; theta_max = -1.0; delta_w_last = 0.0;
        continue;
      }
      status = OMG_RESTORATION_FAILED; break;
    }
    if (!ftype) {
      const double th = (1.0 - GAMMA_THETA) * theta, ph = phi - GAMMA_PHI * theta; int nf = 0;
      for (int q = 0; q < nfilt; ++q) if (!(filt[2 * q] >= th && filt[2 * q + 1] >= ph)) { filt[2 * nf] = filt[2 * q]; filt[2 * nf + 1] = filt[2 * q + 1]; ++nf; }
      if (nf >= MAXF) { memmove(filt, filt + 2, sizeof(double) * 2 * (nf - 1)); --nf; }
      filt[2 * nf] = th; filt[2 * nf + 1] = ph; nfilt = nf + 1;
    }
    f = ft;
    for (int j = 0; j < n + 1 + T->n_mid; ++j) xe[j] = xt[j];
    for (int i = 0; i < m; ++i) {
      const int r = rt[i]; g[i] = gt[i]; y[i] += alpha * dy[i];
      if (!(r & 4)) { const double si = st[i]; s[i] = si;
        if (r & 1) { const double dl = si - sL[i]; double z = zL[i] + a_d * dzL[i]; z = fmin(fmax(z, mu / (KAPPA_SIGMA * dl)), KAPPA_SIGMA * mu / dl); zL[i] = z; }
        if (r & 2) { const double du = sU[i] - si; double z = zU[i] + a_d * dzU[i]; z = fmin(fmax(z, mu / (KAPPA_SIGMA * du)), KAPPA_SIGMA * mu / du); zU[i] = z; } }
    }
  }
  for (int i = 0; i < n; ++i) xout[i] = xe[i];
  for (int i = 0; i < m; ++i) lamout[i] = y[i] * dsc[i] / fsc;
  *fout = f / fsc; *status_out = status; *iters_out = iter;
}

typedef struct {
  const omg_tables* T; const omg_options* O; int B, shared, Nw;
  const double *x0, *p, *lbg, *ubg, *lam0; double *x, *lam, *f; int *status, *iters;
  int* next;
  const SpSym* sym;
} Job;

static void* worker(void* arg) {
  Job* J = (Job*)arg;
  const omg_tables* T = J->T;
  Work w;
  work_alloc(&w, T, J->Nw);
  if (J->sym) work_attach_sparse(&w, J->sym);
  for (;;) {
    const int b = __atomic_fetch_add(J->next, 1, __ATOMIC_RELAXED);
    if (b >= J->B) break;
    const double* lb = J->lbg + (J->shared ? 0 : (size_t)b * T->m);
    const double* ub = J->ubg + (J->shared ? 0 : (size_t)b * T->m);
    int ne = 0;
    for (int i = 0; i < T->m; ++i) if (lb[i] == ub[i]) ++ne;
    if (T->n + ne > J->Nw) { J->status[b] = OMG_ERROR_IN_STEP_COMPUTATION; J->iters[b] = 0; continue; }
    solve_one(T, J->O, &w, J->x0 + (size_t)b * T->n, J->p + (size_t)b * T->n_par, lb, ub,
              J->lam0 ? J->lam0 + (size_t)b * T->m : 0, J->x + (size_t)b * T->n,
              J->lam + (size_t)b * T->m, J->f + b, J->status + b, J->iters + b);
  }
  work_free(&w);
  return 0;
}

/* Solve B instances on `threads` host threads (one instance per thread at a time). */
int oracle_solve_batch(const omg_tables* T, const omg_options* O, int B, const double* x0,
                       const double* p, const double* lbg, const double* ubg, int shared,
                       const double* lam0, double* x, double* lam, double* f, int* status,
                       int* iters, int threads) {
  int neq = 0;
  for (int i = 0; i < T->m; ++i) if (T->lbg[i] == T->ubg[i]) ++neq;
  int next = 0;
  Job J = {T, O, B, shared, shared ? T->n + neq + 8 : T->n + T->m, x0, p, lbg, ubg, lam0,
           x, lam, f, status, iters, &next, 0};
  SpSym* sym = 0;
  if (g_linear_solver == 1 && shared) {     /* the structure follows the tables' equality rows */
    int same = 1;
    for (int i = 0; i < T->m; ++i) if ((lbg[i] == ubg[i]) != (T->lbg[i] == T->ubg[i])) same = 0;
    if (same) sym = spsym_build(T);
  }
  J.sym = sym;
  if (shared) {   /* size the border from the bounds actually passed */
    int ne = 0;
    for (int i = 0; i < T->m; ++i) if (lbg[i] == ubg[i]) ++ne;
    J.Nw = T->n + ne + 8;
  }
  if (threads < 1) threads = 1;
  if (threads > B) threads = B;
  if (threads == 1) { worker(&J); spsym_free(sym); return 0; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  for (int t = 0; t < threads; ++t) pthread_create(&th[t], 0, worker, &J);
  for (int t = 0; t < threads; ++t) pthread_join(th[t], 0);
  free(th);
  spsym_free(sym);
  return 0;
}

/* ---------------------------------------------------------------------------
 * Feasibility phase (oracle/ipm_ref.py feasibility_lm; product: omg_feas_batch):
 * Levenberg-Marquardt on v(x) = g - clip(g, lbg, ubg).  Same order of operations as
 * the CUDA kernel: normal equations summed over the Jacobian slots in ascending
 * order, right-looking dense Cholesky with the right-hand side carried as row n.
 * ------------------------------------------------------------------------- */
static double feas_residual(const omg_tables* T, const double* V, double* xq, const double* lb,
                            const double* ub, double* v, double* vmax) {
  eval_mids(T, V, xq);
  double ss = 0.0, mx = 0.0;
  for (int i = 0; i < T->m; ++i) {
    const double g = eval_slot(&T->G, i, V, xq);
    const double vi = (g < lb[i]) ? g - lb[i] : ((g > ub[i]) ? g - ub[i] : 0.0);
    v[i] = vi; ss += vi * vi; if (fabs(vi) > mx) mx = fabs(vi);
  }
  *vmax = mx;
  return 0.5 * ss;
}

int oracle_feas_batch(const omg_tables* T, int B, const double* x0, const double* p,
                      const double* lbg, const double* ubg, int shared, int max_steps,
                      double* x, double* viol, int* steps_out) {
  const int n = T->n, m = T->m, n_xe = n + 1 + T->n_mid;
  double* V = xalloc(sizeof(double) * T->n_v);
  double* jx = xalloc(sizeof(double) * (T->nnz_jx > T->nnz_j ? T->nnz_jx : T->nnz_j));
  double* jv = xalloc(sizeof(double) * T->nnz_j);
  double* v = xalloc(sizeof(double) * m); double* vt = xalloc(sizeof(double) * m);
  double* A = xalloc(sizeof(double) * n * n); double* L = xalloc(sizeof(double) * (n + 1) * n);
  double* rhs = xalloc(sizeof(double) * n); double* dx = xalloc(sizeof(double) * n);
  double* xe = xalloc(sizeof(double) * n_xe); double* xt = xalloc(sizeof(double) * n_xe);
  int* rs = xalloc(sizeof(int) * (m + 1));       /* slot range of each row (slots are row-sorted) */
  for (int s = 0; s < T->nnz_j; ++s) rs[T->jrow[s] + 1]++;
  for (int i = 0; i < m; ++i) rs[i + 1] += rs[i];
  for (int b = 0; b < B; ++b) {
    const double* lb = lbg + (shared ? 0 : (size_t)b * m);
    const double* ub = ubg + (shared ? 0 : (size_t)b * m);
    eval_tape(T, p + (size_t)b * T->n_par, V);
    for (int i = 0; i < n_xe; ++i) xe[i] = xt[i] = (i < n) ? x0[(size_t)b * n + i] : ((i == n) ? 1.0 : 0.0);
    double vmax, lam = 1e-3;
    double phi = feas_residual(T, V, xe, lb, ub, v, &vmax);
    int steps = 0;
    while (steps < max_steps && vmax > 1e-8) {
      eval_jx(T, V, xe, jx);
      for (int s = 0; s < T->nnz_j; ++s) jv[s] = jac_slot(T, s, V, xe, jx);
      memset(A, 0, sizeof(double) * n * n);
      for (int c = 0; c < n; ++c) rhs[c] = 0.0;
      for (int s1 = 0; s1 < T->nnz_j; ++s1) {
        const int c = T->jcol[s1], r = T->jrow[s1];
        if (v[r] == 0.0) continue;
        rhs[c] -= jv[s1] * v[r];
        for (int s2 = rs[r]; s2 < rs[r + 1]; ++s2) A[(size_t)c * n + T->jcol[s2]] += jv[s1] * jv[s2];
      }
      int accepted = 0;
      for (int attempt = 0; attempt < 12 && !accepted; ++attempt) {
        for (int i = 0; i <= n; ++i)
          for (int k = 0; k < n; ++k)
            L[(size_t)i * n + k] = (i == n) ? rhs[k] : ((k <= i) ? A[(size_t)i * n + k] + ((k == i) ? lam : 0.0) : 0.0);
        int ok = 1;
        for (int j = 0; j < n && ok; ++j) {
          const double d = L[(size_t)j * n + j];
          if (!(d > 0.0 && d < 1e300)) { ok = 0; break; }
          const double pv = sqrt(d), inv = 1.0 / pv;
          L[(size_t)j * n + j] = pv;
          for (int i = j + 1; i <= n; ++i) L[(size_t)i * n + j] *= inv;
          for (int i = j + 1; i <= n; ++i) {
            double* Li = L + (size_t)i * n;
            const double lij = Li[j];
            const int kend = (i < n) ? i : n - 1;
            for (int k = j + 1; k <= kend; ++k) Li[k] -= lij * L[(size_t)k * n + j];
          }
        }
        double pt = 0.0, vmt = 0.0;
        if (ok) {
          double* w = L + (size_t)n * n;
          for (int j = n - 1; j >= 0; --j) {
            dx[j] = w[j] / L[(size_t)j * n + j];
            for (int k = 0; k < j; ++k) w[k] -= L[(size_t)j * n + k] * dx[j];
          }
          for (int i = 0; i < n; ++i) xt[i] = xe[i] + dx[i];
          pt = feas_residual(T, V, xt, lb, ub, vt, &vmt);
        }
        if (ok && pt < phi) {
          memcpy(xe, xt, sizeof(double) * n_xe); memcpy(v, vt, sizeof(double) * m);
          phi = pt; vmax = vmt; lam = fmax(lam / 10.0, 1e-12); accepted = 1;
        } else {
          lam *= 10.0;
        }
      }
      if (!accepted) break;
      ++steps;
    }
    memcpy(x + (size_t)b * n, xe, sizeof(double) * n);
    viol[b] = vmax; steps_out[b] = steps;
  }
  void* all[] = {V, jx, jv, v, vt, A, L, rhs, dx, xe, xt, rs};
  for (unsigned k = 0; k < sizeof(all) / sizeof(all[0]); ++k) free(all[k]);
  return 0;
}

void oracle_default_options(omg_options* o) {
  o->tol = 1e-3; o->constr_viol_tol = 1e-4; o->dual_inf_tol = 1.0; o->compl_inf_tol = 1e-4;
  o->mu_init = 0.1; o->bound_push = 1e-3; o->bound_frac = 1e-3; o->mult_bound_push = 1e-3;
  o->bound_relax_factor = 1e-8; o->scaling_max_gradient = 100.0; o->max_iter = 3000; o->trace = 0;
  o->max_restarts = 5; o->soft_resto = 1; o->restart_mu = 1.0; o->restart_push = 0.1;
  o->inertia_mode = 0; o->reserved = 0;
}
