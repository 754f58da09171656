// omg_b200.cu -- batched primal-dual interior-point solve of OMG-tools' spline
// NLP on B200 (sm_100a).  One thread block per problem instance; the condensed
// KKT matrix lives packed in shared memory for the whole solve.
//
// Replaces the CasADi+IPOPT call of the reference (omgtools/problems/
// problem.py:113, optilayer.py:49-60).  Algorithm = oracle/ipm_ref.py (IPOPT
// semantics, Waechter & Biegler 2006); tables = basics/lowering.py.
//
// Kernel phases per interior-point iteration (all inside one launch):
//   rows   : g, J slots, residuals, error terms        (thread per constraint row)
//   cols   : grad f, dual residual J^T y               (thread per variable)
//   barrier: convergence test / monotone mu update     (uniform scalar code)
//   assem  : H = W + J^T Sigma J gathered into packed K, equality border, rhs row
//   factor : blocked Cholesky (16-column panels, warp-shuffle diagonal block,
//            4x4 register-tiled trailing update), rhs eliminated as extra row
//   solve  : blocked back substitution
//   step   : ds, dy, dz, fraction-to-boundary, filter line search, update
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>

#include "../../include/omg_b200.h"

#define NT 256            // threads per block
#define NWARP (NT / 32)
#define NB 16             // Cholesky panel width
#define MAXF 32           // filter capacity
#define NRED 12           // values per fused block reduction
#define FULL 0xffffffffu
#define TRACE_COLS 8
#define TRACE_ROWS 512

// ---------------------------------------------------------------------------
// device-side table views
// ---------------------------------------------------------------------------
struct PTerm {                     // packed term (width <= 2): one 16-byte load
  double coef; unsigned short cidx, x0, x1, lrow;
};

struct TL {
  int n_out, n_terms, width;
  const int* ptr; const double* coef; const int* cidx; const int* xi; const int* lrow;
  const PTerm* pk;                 // non-null when width <= 2
};

struct DevTab {
  int n, m, n_par, n_v, n_tape, n_levels, nnz_j, nnz_w, nnz_h, n_hp;
  int n_eq, N;                     // structural equality rows; N = n + n_eq (order of K)
  int env_size, n_panels, max_panel_rows;
  const int *eq_rows, *pos_var, *pos_eq, *ksign, *env_first, *env_ptr, *hdst, *jdst, *kdiag,
            *panel_ptr, *panel_rows, *panel_cmin;
  const unsigned* hp_pack;         // pair list packed: s1 | s2 << 16 (nnz_j < 65536)
  const int *tape_func, *tape_ptr, *tape_fac, *level_ptr; const double* tape_coef;
  TL G, F, DF, J, W;
  const int *jrow, *jcol, *jrow_ptr, *jcol_ptr, *jcol_slot;
  const int *wrow, *wcol, *w2h, *hrow, *hcol, *hp_ptr, *hp_s1, *hp_s2, *hp_row;
};

struct Smem {                      // offsets in doubles
  int K, Pt, PtS, xe, xt, dx, u, gf, diag0, invd, V, red, filt, rbase, total;
  int sgn, eptr, efirst, pptr, prow, pcmin;   // structure arrays cached in shared memory
  int LDP;
};

struct Batch {
  int B; int bounds_shared;
  const double *x0, *p, *lbg, *ubg, *lam0;
  double *x, *lam, *f; int *status, *iters;
  double* dscr; int* iscr; int dscr_stride, iscr_stride;
  int* counter; double* trace;
};

struct Ctl {                       // uniform per-block control scalars
  double mu, tau, theta_max, theta_min, delta_w, delta_w_last, delta_c;
  double alpha, a_p, a_d, theta, phi, gphi, f, ft, fsc;
  int n_eq, n_bounds, iter, status, fail, eq_fail, first_try, nfilt, accepted, inst;
};

// IPOPT constants not exposed as options (oracle/ipm_ref.py DEFAULTS)
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
#define DBL_EPS 2.220446049250313e-16

enum { OP_MAX = 0, OP_MIN = 1, OP_SUM = 2 };

// phase timers (debug, opt.trace=1, instance 0): cycles per phase summed over the solve
#define NPHASE 16
#define TICK(k) do { if (tracing && threadIdx.x == 0) { const long long t_ = clock64(); \
  phase_cyc[k] += (double)(t_ - phase_t0); phase_t0 = t_; } } while (0)

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }

__device__ __forceinline__ PTerm load_pterm(const PTerm* p) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  PTerm t;
  t.coef = __hiloint2double((int)r.y, (int)r.x);
  t.cidx = (unsigned short)(r.z & 0xffffu); t.x0 = (unsigned short)(r.z >> 16);
  t.x1 = (unsigned short)(r.w & 0xffffu); t.lrow = (unsigned short)(r.w >> 16);
  return t;
}

__device__ __forceinline__ double eval_slot(const TL& L, int s, const double* __restrict__ V,
                                            const double* __restrict__ xe) {
  double acc = 0.0;
  const int lo = L.ptr[s], hi = L.ptr[s + 1], w = L.width;
  if (L.pk) {
    for (int t = lo; t < hi; ++t) {
      const PTerm q = load_pterm(L.pk + t);
      acc += q.coef * V[q.cidx] * xe[q.x0] * xe[q.x1];
    }
    return acc;
  }
  for (int t = lo; t < hi; ++t) {
    double v = L.coef[t] * V[L.cidx[t]];
    for (int k = 0; k < w; ++k) v *= xe[L.xi[t * w + k]];
    acc += v;
  }
  return acc;
}

// Fused block reduction of NR values; every thread returns with the results.
template <int NR>
__device__ __forceinline__ void block_reduce(double (&v)[NR], const int (&op)[NR], double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double o = __shfl_down_sync(FULL, v[r], off);
      if (op[r] == OP_MAX) v[r] = fmax(v[r], o);
      else if (op[r] == OP_MIN) v[r] = fmin(v[r], o);
      else v[r] += o;
    }
    if (lane == 0) red[warp * NR + r] = v[r];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double a = red[r];
#pragma unroll
    for (int w = 1; w < NWARP; ++w) {
      double o = red[w * NR + r];
      if (op[r] == OP_MAX) a = fmax(a, o);
      else if (op[r] == OP_MIN) a = fmin(a, o);
      else a += o;
    }
    v[r] = a;
  }
  __syncthreads();
}

__device__ __forceinline__ bool cmp_le(double lhs, double rhs, double base) {
  return lhs - rhs <= 10.0 * DBL_EPS * fabs(base);
}

// ---------------------------------------------------------------------------
// Blocked right-looking factorisation K = L S L^T on envelope storage.
//   row i (permuted order) is stored from column first[i] (multiple of NB) to i
//   at K[env_ptr[i] + j - first[i]]; row N is the right-hand side (never a
//   pivot), so after the sweep it holds S L^{-1} r.
// Per 16-column panel: (1) warp 0 factors the diagonal block in registers with
// shuffles, (2) one thread per reached row does the panel solve, (3) the
// trailing update runs over the rows the panel reaches, 32x16 super-tiles per
// warp, 4x4 register tiles per lane with 128-bit shared loads.
// Pivot j must satisfy sign[j]*pivot > PIV_TOL*|K_jj| (variables) or > 0
// (equality rows); otherwise ctl->fail (eq_fail for an equality pivot).
// ---------------------------------------------------------------------------
struct KS {   // shared-memory copies of the KKT structure arrays
  const double* sgn; const int* eptr; const int* efirst; const int* pptr; const int* prow;
  const int* pcmin;
};

__device__ void factor_env(const DevTab& T, const KS& ks, double* __restrict__ K, double* __restrict__ Pt,
                           double* __restrict__ PtS, int LDP, const double* __restrict__ diag0,
                           double* __restrict__ invd, int* __restrict__ rbase, Ctl* ctl, double* pc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = T.N;
  int* rrow = rbase + LDP;
  long long t0 = clock64();
#define FT(k) do { if (pc && tid == 0) { const long long t_ = clock64(); pc[k] += (double)(t_ - t0); t0 = t_; } } while (0)
  for (int pb = 0; pb < T.n_panels; ++pb) {
    const int kb = pb * NB;
    const int nb = min(NB, N - kb);
    // ---- 1. diagonal block (warp 0) ------------------------------------------
    if (warp == 0) {
      double a[NB];
      const int row = kb + lane;
      const int base = (lane < nb) ? (ks.eptr[row] + kb - ks.efirst[row]) : 0;
#pragma unroll
      for (int c = 0; c < NB; ++c) a[c] = (lane < nb && c <= lane) ? K[base + c] : 0.0;
      const double my_s = (lane < nb) ? ks.sgn[row] : 1.0;
      const double my_thr = (lane < nb && my_s > 0.0) ? PIV_TOL * fmax(diag0[row], 1e-300) : 0.0;
      bool ok = true;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < nb && ok) {
          const double sj = __shfl_sync(FULL, my_s, j);
          const double thr = __shfl_sync(FULL, my_thr, j);
          const double d = sj * __shfl_sync(FULL, a[j], j);
          if (!(d > thr) || !isfinite(d)) {
            ok = false;
            if (lane == 0) { ctl->fail = 1; ctl->eq_fail = (sj < 0.0) ? 1 : 0; }
          } else {
            const double inv = rsqrt(d);   // 62 cycles on B200 (tools/ubench/lat.cu), faster than a float seed + Newton
            if (lane == j) { a[j] = d * inv; invd[kb + j] = inv; }
            else if (lane > j) a[j] *= inv * sj;
#pragma unroll
            for (int k = j + 1; k < NB; ++k) {
              const double lkj = __shfl_sync(FULL, a[j], k);
              if (lane >= k) a[k] -= sj * a[j] * lkj;
            }
          }
        }
      }
      if (ok) {
#pragma unroll
        for (int c = 0; c < NB; ++c) if (lane < nb && c <= lane) K[base + c] = a[c];
      }
    }
    __syncthreads();
    FT(7);
    if (ctl->fail) return;
    // ---- 2. panel solve over the rows this panel reaches ----------------------
    const int p0 = ks.pptr[pb];
    const int nrows = ks.pptr[pb + 1] - p0;
    for (int rr = tid; rr < nrows; rr += NT) {
      const int r = ks.prow[p0 + rr];
      const int rb = ks.eptr[r] - ks.efirst[r];
      double sg[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) sg[c] = (c < nb) ? ks.sgn[kb + c] : 1.0;
      rbase[rr] = rb; rrow[rr] = r;
      double* Kr = K + rb + kb;
      double a[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) a[c] = (c < nb) ? Kr[c] : 0.0;
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        if (c < nb) {
          double v = a[c];
          const int rc = kb + c;
          const double* Lc = K + ks.eptr[rc] + kb - ks.efirst[rc];
#pragma unroll
          for (int j = 0; j < NB; ++j)
            if (j < c) v -= sg[j] * a[j] * Lc[j];
          a[c] = v * invd[rc] * sg[c];
        }
      }
#pragma unroll
      for (int c = 0; c < NB; ++c)
        if (c < nb) { Kr[c] = a[c]; Pt[c * LDP + rr] = a[c]; PtS[c * LDP + rr] = sg[c] * a[c]; }
    }
    // zero padding so that vector loads past nrows are harmless
    for (int q = tid; q < NB * 4; q += NT) {
      const int c = q >> 2, rr = nrows + (q & 3);
      if (rr < LDP) { Pt[c * LDP + rr] = 0.0; PtS[c * LDP + rr] = 0.0; }
    }
    __syncthreads();
    FT(8);
    // ---- 3. trailing update ------------------------------------------------------
    // list positions a (rows) x b (columns), b <= a; super-tiles 32 (a) x 16 (b)
    const int nsa = (nrows + 31) >> 5, nsb = (nrows + 15) >> 4;
    const int ty = lane >> 2, tx = lane & 3;
    for (int st = warp; st < nsa * nsb; st += NWARP) {
      const int sa = st / nsb, sb = st - sa * nsb;
      if (sb * 16 > sa * 32 + 31) continue;          // entirely above the diagonal
      const int a0 = sa * 32 + ty * 4, b0 = sb * 16 + tx * 4;
      if (a0 >= nrows || b0 >= nrows || b0 > a0 + 3) continue;
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
      for (int c = 0; c < nb; ++c) {
        const double2* Pa = reinterpret_cast<const double2*>(Pt + c * LDP + a0);
        const double2* Pb = reinterpret_cast<const double2*>(PtS + c * LDP + b0);
        const double2 r01 = Pa[0], r23 = Pa[1], c01 = Pb[0], c23 = Pb[1];
        const double ri[4] = {r01.x, r01.y, r23.x, r23.y};
        const double ck[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] += ri[a] * ck[b];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int la = a0 + a;
        if (la < nrows) {
          const int rba = rbase[la];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int lb = b0 + b;
            if (lb <= la && lb < nrows) {
              const int cb = rrow[lb];
              if (cb < N) K[rba + cb] -= acc[a][b];
            }
          }
        }
      }
    }
    __syncthreads();
    FT(9);
  }
#undef FT
}

// Back substitution L^T u = w on envelope storage (w = row N of L on entry).
__device__ void back_solve_env(const DevTab& T, const KS& ks, const double* __restrict__ K,
                               const double* __restrict__ invd, double* __restrict__ w) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = T.N;
  for (int pb = T.n_panels - 1; pb >= 0; --pb) {
    const int kb = pb * NB;
    const int nb = min(NB, N - kb);
    if (warp == 0) {
      double wv = (lane < nb) ? w[kb + lane] : 0.0;
      for (int j = nb - 1; j >= 0; --j) {
        const int rj = kb + j;
        const double uj = __shfl_sync(FULL, wv, j) * invd[rj];
        if (lane == j) wv = uj;
        else if (lane < j) wv -= K[ks.eptr[rj] + kb - ks.efirst[rj] + lane] * uj;
      }
      if (lane < nb) w[kb + lane] = wv;
    }
    __syncthreads();
    // columns left of the block: c in [first(block), kb)
    const int cmin = ks.pcmin[pb];
    for (int c = cmin + tid; c < kb; c += NT) {
      double acc = w[c];
      for (int j = 0; j < nb; ++j) {
        const int rj = kb + j, fj = ks.efirst[rj];
        if (c >= fj) acc -= K[ks.eptr[rj] + c - fj] * w[rj];
      }
      w[c] = acc;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// the solver kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 2)
omg_ipm_kernel(const DevTab T, const omg_options O, const Batch A, const Smem S) {
  extern __shared__ double sm[];
  __shared__ Ctl ctl;
  __shared__ double phase_cyc[NPHASE];
  double* K = sm + S.K;
  double* Pt = sm + S.Pt;
  double* PtS = sm + S.PtS;
  double* u = sm + S.u;
  int* rbase = reinterpret_cast<int*>(sm + S.rbase);
  double* xe = sm + S.xe;
  double* xt = sm + S.xt;
  double* dx = sm + S.dx;
  double* gf = sm + S.gf;
  double* diag0 = sm + S.diag0;
  double* invd = sm + S.invd;
  double* V = sm + S.V;
  double* red = sm + S.red;
  double* filt = sm + S.filt;
  const int tid = threadIdx.x;
  const int n = T.n, m = T.m;
  {  // KKT structure arrays -> shared memory (once per block)
    double* sgn = sm + S.sgn;
    int* eptr = reinterpret_cast<int*>(sm + S.eptr);
    int* efirst = reinterpret_cast<int*>(sm + S.efirst);
    int* pptr = reinterpret_cast<int*>(sm + S.pptr);
    int* prow = reinterpret_cast<int*>(sm + S.prow);
    int* pcmin = reinterpret_cast<int*>(sm + S.pcmin);
    for (int i = tid; i < T.N; i += NT) sgn[i] = (double)T.ksign[i];
    for (int i = tid; i < T.N + 2; i += NT) eptr[i] = T.env_ptr[i];
    for (int i = tid; i < T.N + 1; i += NT) efirst[i] = T.env_first[i];
    for (int i = tid; i < T.n_panels + 1; i += NT) pptr[i] = T.panel_ptr[i];
    for (int i = tid; i < T.panel_ptr[T.n_panels]; i += NT) prow[i] = T.panel_rows[i];
    for (int i = tid; i < T.n_panels; i += NT) pcmin[i] = T.panel_cmin[i];
  }
  KS ks;
  ks.sgn = sm + S.sgn; ks.eptr = reinterpret_cast<const int*>(sm + S.eptr);
  ks.efirst = reinterpret_cast<const int*>(sm + S.efirst);
  ks.pptr = reinterpret_cast<const int*>(sm + S.pptr);
  ks.prow = reinterpret_cast<const int*>(sm + S.prow);
  ks.pcmin = reinterpret_cast<const int*>(sm + S.pcmin);
  __syncthreads();

  // per-block global scratch (L2 resident)
  double* D = A.dscr + (size_t)blockIdx.x * A.dscr_stride;
  double* g = D;            double* s = g + m;      double* y = s + m;
  double* zL = y + m;       double* zU = zL + m;    double* dsc = zU + m;
  double* sL = dsc + m;     double* sU = sL + m;    double* sig = sU + m;
  double* wv = sig + m;     double* ds = wv + m;    double* dy = ds + m;
  double* dzL = dy + m;     double* dzU = dzL + m;  double* gt = dzU + m;
  double* st = gt + m;      double* jval = st + m;  double* beq = jval + T.nnz_j;
  double* jsv = beq + m;
  int* I = A.iscr + (size_t)blockIdx.x * A.iscr_stride;
  int* rt = I;              int* eqidx = rt + m;    int* eqrow = eqidx + m;

  for (;;) {
    // ---- fetch next instance ------------------------------------------------
    if (tid == 0) ctl.inst = atomicAdd(A.counter, 1);
    __syncthreads();
    const int inst = ctl.inst;
    if (inst >= A.B) return;
    const double* x0 = A.x0 + (size_t)inst * n;
    const double* par = A.p + (size_t)inst * T.n_par;
    const double* lbg = A.lbg + (A.bounds_shared ? 0 : (size_t)inst * m);
    const double* ubg = A.ubg + (A.bounds_shared ? 0 : (size_t)inst * m);
    const bool tracing = (O.trace != 0) && inst == 0 && A.trace != nullptr;
    long long phase_t0 = clock64();
    if (tracing && tid == 0) for (int k = 0; k < NPHASE; ++k) phase_cyc[k] = 0.0;

    // ---- S1: parameter tape ---------------------------------------------------
    for (int i = tid; i < 1 + T.n_par; i += NT) V[i] = (i == 0) ? 1.0 : par[i - 1];
    __syncthreads();
    for (int l = 0; l < T.n_levels; ++l) {
      for (int e = T.level_ptr[l] + tid; e < T.level_ptr[l + 1]; e += NT) {
        double acc = 0.0;
        for (int t = T.tape_ptr[e]; t < T.tape_ptr[e + 1]; ++t) {
          const int* f = T.tape_fac + 4 * t;
          acc += T.tape_coef[t] * V[f[0]] * V[f[1]] * V[f[2]] * V[f[3]];
        }
        switch (T.tape_func[e]) {
          case 1: acc = 1.0 / acc; break;
          case 2: acc = (acc >= 0.0) ? 1.0 : 0.0; break;
          case 3: acc = (acc > 0.0) ? 1.0 : 0.0; break;
          case 4: acc = sin(acc); break;
          case 5: acc = cos(acc); break;
          case 6: acc = sqrt(acc); break;
          default: break;
        }
        V[1 + T.n_par + e] = acc;
      }
      __syncthreads();
    }
    // ---- S2: x ------------------------------------------------------------------
    for (int i = tid; i <= n; i += NT) { xe[i] = (i < n) ? x0[i] : 1.0; xt[i] = 1.0; }
    __syncthreads();

    // ---- S3: row classification, scaling, starting point -----------------------
    double fmaxv = 0.0;
    for (int j = tid; j < n; j += NT) fmaxv = fmax(fmaxv, fabs(eval_slot(T.DF, j, V, xe)));
    {
      double r1[1] = {fmaxv}; const int o1[1] = {OP_MAX};
      block_reduce<1>(r1, o1, red);
      fmaxv = r1[0];
    }
    const double smg = O.scaling_max_gradient;
    const double fsc = (fmaxv > smg) ? fmax(smg / fmaxv, 1e-8) : 1.0;
    for (int i = tid; i < m; i += NT) {
      double gm = 0.0;
      for (int sl = T.jrow_ptr[i]; sl < T.jrow_ptr[i + 1]; ++sl)
        gm = fmax(gm, fabs(eval_slot(T.J, sl, V, xe)));
      const double d = (gm > smg) ? fmax(smg / gm, 1e-8) : 1.0;
      dsc[i] = d;
      const double lb = lbg[i], ub = ubg[i];
      const bool eq = (lb == ub);
      const bool hL = (lb > -INF_BOUND) && !eq, hU = (ub < INF_BOUND) && !eq;
      rt[i] = (hL ? 1 : 0) | (hU ? 2 : 0) | (eq ? 4 : 0);
      double l = lb * d, u = ub * d;
      beq[i] = l;
      if (hL) l -= O.bound_relax_factor * fmax(1.0, fabs(l));
      if (hU) u += O.bound_relax_factor * fmax(1.0, fabs(u));
      sL[i] = l; sU[i] = u;
      const double gi = d * eval_slot(T.G, i, V, xe);
      g[i] = gi;
      double si = gi;
      const double k1 = O.bound_push, k2 = O.bound_frac;
      double pl = k1 * fmax(1.0, fabs(l)), pu = k1 * fmax(1.0, fabs(u));
      if (hL && hU) { pl = fmin(pl, k2 * (u - l)); pu = fmin(pu, k2 * (u - l)); }
      if (hL) si = fmax(si, l + pl);
      if (hU) si = fmin(si, u - pu);
      s[i] = si;
      double yi = 0.0;
      if (A.lam0) yi = A.lam0[(size_t)inst * m + i] * fsc / d;
      y[i] = yi;
      zL[i] = hL ? fmax(O.mult_bound_push, -yi) : 0.0;
      zU[i] = hU ? fmax(O.mult_bound_push, yi) : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
      int ne = 0, nbnd = 0, bad = 0;
      for (int i = 0; i < m; ++i) {
        const int r = rt[i];
        if (r & 4) {
          if (ne < T.n_eq && T.eq_rows[ne] == i) { eqrow[ne] = i; eqidx[i] = ne; } else bad = 1;
          ++ne;
        } else eqidx[i] = -1;
        nbnd += (r & 1) + ((r >> 1) & 1);
      }
      if (ne != T.n_eq) bad = 1;
      ctl.n_eq = bad ? -1 : ne; ctl.n_bounds = nbnd;
      ctl.mu = O.mu_init; ctl.tau = fmax(TAU_MIN, 1.0 - O.mu_init);
      ctl.theta_max = -1.0; ctl.theta_min = -1.0;
      ctl.delta_w_last = 0.0; ctl.nfilt = 0; ctl.status = -1; ctl.iter = 0;
      ctl.fsc = fsc; ctl.alpha = 0.0; ctl.delta_w = 0.0;
      ctl.f = fsc * eval_slot(T.F, 0, V, xe);
    }
    __syncthreads();
    if (ctl.n_eq < 0) {   // equality pattern differs from the lowered structure
      if (tid == 0) { A.status[inst] = OMG_ERROR_IN_STEP_COMPUTATION; A.iters[inst] = 0; A.f[inst] = 0.0; }
      for (int i = tid; i < n; i += NT) A.x[(size_t)inst * n + i] = x0[i];
      for (int i = tid; i < m; i += NT) A.lam[(size_t)inst * m + i] = 0.0;
      __syncthreads();
      continue;
    }
    const int n_eq = ctl.n_eq;
    const int N = T.N;
    const int n_bounds = ctl.n_bounds;

    // =========================== IP iterations ===============================
    for (int iter = 0;; ++iter) {
      TICK(0);   // setup / previous accept
      // ---- I1: rows: g (kept from trial), Jacobian values, residual terms -------
      double rv[NRED];
      // 0 cinf(max) 1 maxprod(max) 2 minprod(min) 3 viol(max) 4 rsinf(max)
      // 5 rsinf_un(max) 6 ysum 7 zsum 8 theta 9 logsum
      const int rop[NRED] = {OP_MAX, OP_MAX, OP_MIN, OP_MAX, OP_MAX, OP_MAX,
                             OP_SUM, OP_SUM, OP_SUM, OP_SUM, OP_MAX, OP_SUM};
      for (int r = 0; r < NRED; ++r) rv[r] = 0.0;
      rv[2] = 1e300;
      for (int i = tid; i < m; i += NT) {
        const int r = rt[i];
        const double d = dsc[i];
        for (int sl = T.jrow_ptr[i]; sl < T.jrow_ptr[i + 1]; ++sl)
          jval[sl] = d * eval_slot(T.J, sl, V, xe);
        const double gi = g[i], si = s[i], yi = y[i];
        double ci;
        if (r & 4) ci = gi - beq[i];
        else ci = gi - si;
        rv[0] = fmax(rv[0], fabs(ci));
        rv[8] += fabs(ci);
        if (r & 1) { const double dl = si - sL[i]; const double pz = dl * zL[i];
          rv[1] = fmax(rv[1], pz); rv[2] = fmin(rv[2], pz); rv[9] += log(dl); rv[7] += zL[i]; }
        if (r & 2) { const double du = sU[i] - si; const double pz = du * zU[i];
          rv[1] = fmax(rv[1], pz); rv[2] = fmin(rv[2], pz); rv[9] += log(du); rv[7] += zU[i]; }
        const double gun = gi / d;
        if (r & 6) rv[3] = fmax(rv[3], gun - ubg[i]);
        if (r & 5) rv[3] = fmax(rv[3], lbg[i] - gun);
        if (!(r & 4)) { const double rs = fabs(-yi - zL[i] + zU[i]);
          rv[4] = fmax(rv[4], rs); rv[5] = fmax(rv[5], rs * d); }
        rv[6] += fabs(yi);
      }
      __syncthreads();   // jval visible to the column pass
      TICK(1);   // row pass
      // ---- I2: columns: grad f, dual residual ------------------------------------
      for (int j = tid; j < n; j += NT) {
        const double gj = ctl.fsc * eval_slot(T.DF, j, V, xe);
        gf[j] = gj;
        double rx = gj;
        for (int q = T.jcol_ptr[j]; q < T.jcol_ptr[j + 1]; ++q) {
          const int sl = T.jcol_slot[q];
          rx += jval[sl] * y[T.jrow[sl]];
        }
        rv[10] = fmax(rv[10], fabs(rx));
      }
      block_reduce<NRED>(rv, rop, red);
      const double cinf = rv[0], maxprod = rv[1], minprod = rv[2], viol = rv[3];
      const double dinf = fmax(rv[10], rv[4]);
      const double dinf_un = fmax(rv[10], rv[5]) / ctl.fsc;
      const double ysum = rv[6], zsum = rv[7], theta = rv[8], logsum = rv[9];
      const double s_d = fmax(S_MAX, (ysum + zsum) / fmax(1.0, (double)(m + n_bounds))) / S_MAX;
      const double s_c = fmax(S_MAX, zsum / fmax(1.0, (double)n_bounds)) / S_MAX;
      double mu = ctl.mu;
      const double cmpl0 = n_bounds ? fmax(fabs(maxprod), fabs(minprod)) : 0.0;
      const double E0 = fmax(fmax(dinf / s_d, cinf), cmpl0 / s_c);
      TICK(2);   // column pass + reduction
      // ---- I3: termination + barrier update (uniform) ---------------------------
      int status = -1;
      if (!isfinite(E0)) status = OMG_INVALID_NUMBER_DETECTED;
      else if (E0 <= O.tol && dinf_un <= O.dual_inf_tol && viol <= O.constr_viol_tol &&
               cmpl0 / ctl.fsc <= O.compl_inf_tol) status = OMG_SOLVE_SUCCEEDED;
      else if (iter >= O.max_iter) status = OMG_MAX_ITER_EXCEEDED;
      if (tracing && tid == 0 && iter < TRACE_ROWS) {
        double* tr = A.trace + iter * TRACE_COLS;
        tr[0] = iter; tr[1] = ctl.f / ctl.fsc; tr[2] = cinf; tr[3] = dinf; tr[4] = mu; tr[5] = E0;
        tr[6] = ctl.alpha; tr[7] = ctl.delta_w;
      }
      if (status >= 0) { if (tid == 0) { ctl.status = status; ctl.iter = iter; } break; }
      {
        const double mu_min = fmin(O.tol, O.compl_inf_tol * ctl.fsc) / (KAPPA_EPS + 1.0);
        bool changed = false;
        for (;;) {
          const double cm = n_bounds ? fmax(fabs(maxprod - mu), fabs(minprod - mu)) : 0.0;
          const double Emu = fmax(fmax(dinf / s_d, cinf), cm / s_c);
          if (Emu <= KAPPA_EPS * mu && mu > mu_min) {
            mu = fmax(mu_min, fmin(KAPPA_MU * mu, pow(mu, THETA_MU)));
            changed = true;
          } else break;
        }
        __syncthreads();
        if (tid == 0) {
          ctl.mu = mu; ctl.tau = fmax(TAU_MIN, 1.0 - mu);
          if (changed) ctl.nfilt = 0;
          if (ctl.theta_max < 0.0) {
            ctl.theta_max = THETA_MAX_FACT * fmax(1.0, theta);
            ctl.theta_min = THETA_MIN_FACT * fmax(1.0, theta);
          }
          ctl.theta = theta;
          ctl.phi = ctl.f - mu * logsum;
          ctl.delta_w = 0.0; ctl.delta_c = 0.0; ctl.first_try = 1;
        }
      }
      __syncthreads();
      const double tau = ctl.tau;
      TICK(3);   // barrier logic
      // ---- I4: Sigma and w = Sigma r_d + phi_s ------------------------------------
      for (int i = tid; i < m; i += NT) {
        const int r = rt[i];
        double sg = 0.0, ph = 0.0, rd = 0.0;
        if (!(r & 4)) {
          const double si = s[i];
          rd = g[i] - si;
          if (r & 1) { const double dl = si - sL[i]; sg += zL[i] / dl; ph -= mu / dl; }
          if (r & 2) { const double du = sU[i] - si; sg += zU[i] / du; ph += mu / du; }
        }
        sig[i] = sg;
        wv[i] = (r & 4) ? y[i] : (sg * rd + ph);
        for (int sl = T.jrow_ptr[i]; sl < T.jrow_ptr[i + 1]; ++sl) jsv[sl] = sg * jval[sl];
      }
      // ---- I6: Hessian slots into gt-scratch? -> kept in st[] (nnz_w <= m assumed no) ---
      __syncthreads();

      TICK(4);   // sigma pass
      // ---- I7/I8: assemble + factorise, with inertia correction -----------------
      for (;;) {
        for (int q = tid; q < T.env_size; q += NT) K[q] = 0.0;
        __syncthreads();
        // H positions: gather J^T Sigma J (+ delta_w on the diagonal)
        for (int q = tid; q < T.nnz_h; q += NT) {
          double acc = 0.0;
          for (int e = T.hp_ptr[q]; e < T.hp_ptr[q + 1]; ++e) {
            const unsigned pk = __ldg(T.hp_pack + e);
            acc += jsv[pk & 0xffffu] * jval[pk >> 16];
          }
          if (T.hrow[q] == T.hcol[q]) acc += ctl.delta_w;
          K[T.hdst[q]] = acc;
        }
        __syncthreads();
        TICK(5);   // zero + H gather
        // Lagrangian Hessian W (lambda = y*dsc, objective factor fsc)
        for (int q = tid; q < T.nnz_w; q += NT) {
          double acc = 0.0;
          const TL& L = T.W;
          for (int t = L.ptr[q]; t < L.ptr[q + 1]; ++t) {
            double v; int lr;
            if (L.pk) {
              const PTerm pt = load_pterm(L.pk + t);
              v = pt.coef * V[pt.cidx] * xe[pt.x0] * xe[pt.x1]; lr = pt.lrow;
            } else {
              v = L.coef[t] * V[L.cidx[t]];
              for (int k = 0; k < L.width; ++k) v *= xe[L.xi[t * L.width + k]];
              lr = L.lrow[t];
            }
            v *= (lr < m) ? (y[lr] * dsc[lr]) : ctl.fsc;
            acc += v;
          }
          K[T.hdst[T.w2h[q]]] += acc;
        }
        __syncthreads();
        for (int j = tid; j < n; j += NT) {
          const int pj = T.pos_var[j];
          diag0[pj] = fabs(K[T.kdiag[pj]]);
        }
        // equality border + right-hand-side row
        const int rhs0 = ks.eptr[N];
        for (int k = tid; k < n_eq; k += NT) {
          const int i = eqrow[k], pk = T.pos_eq[k];
          for (int sl = T.jrow_ptr[i]; sl < T.jrow_ptr[i + 1]; ++sl) K[T.jdst[sl]] = jval[sl];
          K[T.kdiag[pk]] = -ctl.delta_c;
          diag0[pk] = ctl.delta_c;
          K[rhs0 + pk] = -(g[i] - beq[i]);
        }
        for (int j = tid; j < n; j += NT) {
          double acc = gf[j];
          for (int q = T.jcol_ptr[j]; q < T.jcol_ptr[j + 1]; ++q) {
            const int sl = T.jcol_slot[q];
            acc += jval[sl] * wv[T.jrow[sl]];
          }
          K[rhs0 + T.pos_var[j]] = -acc;
        }
        if (tid == 0) { ctl.fail = 0; ctl.eq_fail = 0; }
        __syncthreads();
        TICK(6);   // W + border + rhs
        factor_env(T, ks, K, Pt, PtS, S.LDP, diag0, invd, rbase, &ctl, tracing ? phase_cyc : nullptr);
        __syncthreads();
        phase_t0 = clock64();
        if (!ctl.fail) break;
        // inertia correction (IPOPT algorithm IC on the condensed matrix)
        if (tid == 0) {
          if (ctl.eq_fail) ctl.delta_c = DELTA_C_VAL * pow(mu, DELTA_C_EXP);
          if (ctl.first_try) {
            ctl.delta_w = (ctl.delta_w_last == 0.0) ? DELTA_W0
                          : fmax(DELTA_W_MIN, KAPPA_W_MINUS * ctl.delta_w_last);
            ctl.first_try = 0;
          } else {
            ctl.delta_w *= (ctl.delta_w_last == 0.0) ? KAPPA_W_PLUS_FIRST : KAPPA_W_PLUS;
          }
        }
        __syncthreads();
        if (ctl.delta_w > DELTA_W_MAX) break;
      }
      if (ctl.fail) {
        if (tid == 0) { ctl.status = OMG_ERROR_IN_STEP_COMPUTATION; ctl.iter = iter; }
        __syncthreads();
        break;
      }
      if (tid == 0 && ctl.delta_w > 0.0) ctl.delta_w_last = ctl.delta_w;
      // ---- I9: solve -----------------------------------------------------------
      for (int j = tid; j < N; j += NT) u[j] = K[ks.eptr[N] + j];
      __syncthreads();
      back_solve_env(T, ks, K, invd, u);
      TICK(10);  // back substitution
      for (int j = tid; j < n; j += NT) dx[j] = u[T.pos_var[j]];
      for (int k = tid; k < n_eq; k += NT) dx[n + k] = u[T.pos_eq[k]];
      __syncthreads();
      // ---- I10: ds, dy, dz, fraction to the boundary --------------------------
      double sv[4];   // 0 a_p(min) 1 a_d(min) 2 gphi(sum) 3 unused
      const int sop[4] = {OP_MIN, OP_MIN, OP_SUM, OP_SUM};
      sv[0] = 1.0; sv[1] = 1.0; sv[2] = 0.0; sv[3] = 0.0;
      for (int i = tid; i < m; i += NT) {
        const int r = rt[i];
        double jd = 0.0;
        for (int sl = T.jrow_ptr[i]; sl < T.jrow_ptr[i + 1]; ++sl) jd += jval[sl] * dx[T.jcol[sl]];
        if (r & 4) {
          ds[i] = 0.0; dy[i] = dx[n + eqidx[i]]; dzL[i] = 0.0; dzU[i] = 0.0;
        } else {
          const double si = s[i];
          const double dsi = jd + (g[i] - si);
          double ph = 0.0, dl = 1.0, du = 1.0, a = 0.0, b = 0.0;
          if (r & 1) { dl = si - sL[i]; ph -= mu / dl;
            a = mu / dl - zL[i] - (zL[i] / dl) * dsi;
            if (dsi < 0.0) sv[0] = fmin(sv[0], -tau * dl / dsi);
            if (a < 0.0) sv[1] = fmin(sv[1], -tau * zL[i] / a); }
          if (r & 2) { du = sU[i] - si; ph += mu / du;
            b = mu / du - zU[i] + (zU[i] / du) * dsi;
            if (dsi > 0.0) sv[0] = fmin(sv[0], tau * du / dsi);
            if (b < 0.0) sv[1] = fmin(sv[1], -tau * zU[i] / b); }
          ds[i] = dsi; dzL[i] = a; dzU[i] = b;
          dy[i] = sig[i] * dsi + ph - y[i];
          sv[2] += ph * dsi;
        }
      }
      for (int j = tid; j < n; j += NT) sv[2] += gf[j] * dx[j];
      block_reduce<4>(sv, sop, red);
      const double a_p = sv[0], a_d = sv[1], gphi = sv[2];
      TICK(11);  // step pass
      // ---- I11: filter line search --------------------------------------------
      const double theta0 = ctl.theta, phi0 = ctl.phi;
      double a_min;
      if (gphi < 0.0) {
        a_min = fmin(GAMMA_THETA, GAMMA_PHI * theta0 / (-gphi));
        if (theta0 <= ctl.theta_min)
          a_min = fmin(a_min, DELTA_LS * pow(theta0, S_THETA) / pow(-gphi, S_PHI));
      } else a_min = GAMMA_THETA;
      a_min *= GAMMA_ALPHA;
      double alpha = a_p;
      bool accepted = false, ftype = false;
      double ft = 0.0;
      int n_ls = 0;
      while (alpha >= a_min && n_ls < MAX_LS) {
        ++n_ls;
        for (int j = tid; j < n; j += NT) xt[j] = xe[j] + alpha * dx[j];
        __syncthreads();
        double tv[3];  // 0 theta 1 logsum 2 unused
        const int top[3] = {OP_SUM, OP_SUM, OP_SUM};
        tv[0] = 0.0; tv[1] = 0.0; tv[2] = 0.0;
        for (int i = tid; i < m; i += NT) {
          const int r = rt[i];
          const double gi = dsc[i] * eval_slot(T.G, i, V, xt);
          gt[i] = gi;
          if (r & 4) tv[0] += fabs(gi - beq[i]);
          else {
            const double si = s[i] + alpha * ds[i];
            st[i] = si;
            tv[0] += fabs(gi - si);
            if (r & 1) tv[1] += log(si - sL[i]);
            if (r & 2) tv[1] += log(sU[i] - si);
          }
        }
        block_reduce<3>(tv, top, red);
        ft = ctl.fsc * eval_slot(T.F, 0, V, xt);
        const double tht = tv[0], pht = ft - mu * tv[1];
        bool ok = isfinite(pht) && isfinite(tht) && tht <= ctl.theta_max;
        if (ok) {
          const int nf = ctl.nfilt;
          for (int q = 0; q < nf; ++q)
            if (!(tht < filt[2 * q] || pht < filt[2 * q + 1])) { ok = false; break; }
        }
        ftype = false;
        if (ok) {
          const bool switching = (theta0 <= ctl.theta_min && gphi < 0.0 &&
                                  alpha * pow(-gphi, S_PHI) > DELTA_LS * pow(theta0, S_THETA));
          if (switching) { ok = cmp_le(pht - phi0, ETA_PHI * alpha * gphi, phi0); ftype = ok; }
          else ok = cmp_le(tht, (1.0 - GAMMA_THETA) * theta0, theta0) ||
                    cmp_le(pht - phi0, -GAMMA_PHI * theta0, phi0);
        }
        if (ok) { accepted = true; break; }
        alpha *= 0.5;
      }
      if (!accepted) {
        if (tid == 0) { ctl.status = OMG_RESTORATION_FAILED; ctl.iter = iter; }
        __syncthreads();
        break;
      }
      __syncthreads();
      if (tid == 0) {
        if (!ftype) {        // augment the filter, dropping dominated entries
          const double th = (1.0 - GAMMA_THETA) * theta0, ph = phi0 - GAMMA_PHI * theta0;
          int nf = 0;
          for (int q = 0; q < ctl.nfilt; ++q)
            if (!(filt[2 * q] >= th && filt[2 * q + 1] >= ph)) {
              filt[2 * nf] = filt[2 * q]; filt[2 * nf + 1] = filt[2 * q + 1]; ++nf; }
          if (nf >= MAXF) {
            for (int q = 1; q < nf; ++q) { filt[2 * (q - 1)] = filt[2 * q]; filt[2 * (q - 1) + 1] = filt[2 * q + 1]; }
            --nf;
          }
          filt[2 * nf] = th; filt[2 * nf + 1] = ph; ++nf;
          ctl.nfilt = nf;
        }
        ctl.f = ft; ctl.alpha = alpha;
      }
      TICK(12);  // line search
      // ---- I12: accept -------------------------------------------------------------
      for (int j = tid; j < n; j += NT) xe[j] = xt[j];
      for (int i = tid; i < m; i += NT) {
        const int r = rt[i];
        g[i] = gt[i];
        y[i] += alpha * dy[i];
        if (!(r & 4)) {
          const double si = st[i];
          s[i] = si;
          if (r & 1) { const double dl = si - sL[i]; double z = zL[i] + a_d * dzL[i];
            z = fmin(fmax(z, mu / (KAPPA_SIGMA * dl)), KAPPA_SIGMA * mu / dl); zL[i] = z; }
          if (r & 2) { const double du = sU[i] - si; double z = zU[i] + a_d * dzU[i];
            z = fmin(fmax(z, mu / (KAPPA_SIGMA * du)), KAPPA_SIGMA * mu / du); zU[i] = z; }
        }
      }
      __syncthreads();
    }  // iterations

    // ---- write results -----------------------------------------------------------
    __syncthreads();
    for (int i = tid; i < n; i += NT) A.x[(size_t)inst * n + i] = xe[i];
    for (int i = tid; i < m; i += NT) A.lam[(size_t)inst * m + i] = y[i] * dsc[i] / ctl.fsc;
    if (tracing && tid == 0) {
      TICK(13);
      double* tr = A.trace + (TRACE_ROWS - 2) * TRACE_COLS;
      for (int k = 0; k < NPHASE; ++k) tr[k] = phase_cyc[k];
    }
    if (tid == 0) {
      A.f[inst] = ctl.f / ctl.fsc;
      A.status[inst] = ctl.status;
      A.iters[inst] = ctl.iter;
    }
    __syncthreads();
  }
}

// warm-start shift: x[b, off + c*len + i] <- sum_k T[i,k] x[b, off + c*len + k]
__global__ void omg_shift_kernel(double* x, int B, int n, int n_blocks, const int* offs,
                                 const int* lens, const int* ncols, const int* toffs,
                                 const double* Tm) {
  extern __shared__ double xs[];
  const int b = blockIdx.x;
  if (b >= B) return;
  double* xb = x + (size_t)b * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = xb[i];
  __syncthreads();
  for (int blk = 0; blk < n_blocks; ++blk) {
    const int L = lens[blk], nc = ncols[blk], off = offs[blk];
    const double* Tb = Tm + toffs[blk];
    for (int e = threadIdx.x; e < L * nc; e += blockDim.x) {
      const int c = e / L, i = e % L;
      double acc = 0.0;
      for (int k = 0; k < L; ++k) acc += Tb[i * L + k] * xs[off + c * L + k];
      xb[off + c * L + i] = acc;
    }
  }
}

// ===========================================================================
// host side: C ABI
// ===========================================================================
static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  set_err(std::string(#call) + ": " + cudaGetErrorString(e_)); return -1; } } while (0)

struct omg_problem {
  int device = 0;
  DevTab T;
  Smem S;
  omg_options opt;
  std::vector<void*> allocs;
  int n_sm = 0, ctas_per_sm = 1;
  size_t smem_bytes = 0;
  double* dscr = nullptr; int* iscr = nullptr; int scr_ctas = 0;
  int dscr_stride = 0, iscr_stride = 0;
  int* counter = nullptr; double* trace = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false; int launches = 0;
  // staging buffers for the _host entry point
  double *hx0 = nullptr, *hp = nullptr, *hlb = nullptr, *hub = nullptr, *hlam0 = nullptr,
         *hx = nullptr, *hlam = nullptr, *hf = nullptr;
  int *hst = nullptr, *hit = nullptr; int hostB = 0, host_shared = -1;
};

template <typename Tp>
static const Tp* upload(omg_problem* h, const Tp* src, size_t count, bool* ok) {
  if (count == 0) count = 1;
  void* d = nullptr;
  if (cudaMalloc(&d, count * sizeof(Tp)) != cudaSuccess) { *ok = false; return nullptr; }
  h->allocs.push_back(d);
  if (src && cudaMemcpy(d, src, count * sizeof(Tp), cudaMemcpyHostToDevice) != cudaSuccess) *ok = false;
  return (const Tp*)d;
}

static TL upload_tl(omg_problem* h, const omg_termlist& L, int n_one, bool* ok) {
  TL t;
  t.n_out = L.n_out; t.n_terms = L.n_terms; t.width = L.width;
  t.ptr = upload(h, L.ptr, (size_t)L.n_out + 1, ok);
  t.coef = upload(h, L.coef, (size_t)L.n_terms, ok);
  t.cidx = upload(h, L.cidx, (size_t)L.n_terms, ok);
  t.xi = upload(h, L.xi, (size_t)L.n_terms * L.width, ok);
  t.lrow = L.lrow ? upload(h, L.lrow, (size_t)L.n_terms, ok) : nullptr;
  t.pk = nullptr;
  if (L.width <= 2) {
    std::vector<PTerm> pk((size_t)L.n_terms + 1);
    for (int k = 0; k < L.n_terms; ++k) {
      PTerm& q = pk[k];
      q.coef = L.coef[k]; q.cidx = (unsigned short)L.cidx[k];
      q.x0 = (unsigned short)(L.width >= 1 ? L.xi[(size_t)k * L.width] : n_one);
      q.x1 = (unsigned short)(L.width >= 2 ? L.xi[(size_t)k * L.width + 1] : n_one);
      q.lrow = (unsigned short)(L.lrow ? L.lrow[k] : 0);
    }
    t.pk = upload(h, pk.data(), pk.size(), ok);
  }
  return t;
}

extern "C" {

int omg_abi_version(void) { return OMG_ABI_VERSION; }
const char* omg_last_error(void) { return g_err.c_str(); }

void omg_default_options(omg_options* o) {
  o->tol = 1e-3; o->constr_viol_tol = 1e-4; o->dual_inf_tol = 1.0; o->compl_inf_tol = 1e-4;
  o->mu_init = 0.1; o->bound_push = 1e-3; o->bound_frac = 1e-3; o->mult_bound_push = 1e-3;
  o->bound_relax_factor = 1e-8; o->scaling_max_gradient = 100.0;
  o->max_iter = 3000; o->trace = 0;
}

omg_problem* omg_problem_create(const omg_tables* tb, const omg_options* opt, int device) {
  if (!tb || tb->abi_version != OMG_ABI_VERSION) { set_err("omg_tables ABI version mismatch"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_err("no CUDA device available: libomgb200 has no CPU fallback"); return nullptr; }
  if (device < 0 || device >= ndev) { set_err("invalid device index"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); return nullptr; }
  omg_problem* h = new omg_problem();
  h->device = device;
  if (opt) h->opt = *opt; else omg_default_options(&h->opt);
  bool ok = true;
  DevTab& T = h->T;
  T.n = tb->n; T.m = tb->m; T.n_par = tb->n_par; T.n_v = tb->n_v;
  T.n_tape = tb->n_tape; T.n_levels = tb->n_levels;
  T.nnz_j = tb->nnz_j; T.nnz_w = tb->nnz_w; T.nnz_h = tb->nnz_h; T.n_hp = tb->n_hp;
  T.n_eq = tb->kkt_n_eq; T.N = tb->kkt_n;
  T.env_size = tb->env_size; T.max_panel_rows = tb->max_panel_rows;
  T.n_panels = (tb->kkt_n + NB - 1) / NB;
  if (tb->kkt_n != tb->n + tb->kkt_n_eq) { set_err("inconsistent KKT structure"); ok = false; }
  for (int i = 0; ok && i <= tb->kkt_n; ++i)
    if (tb->env_first[i] % NB != 0) { set_err("env_first must be a multiple of the panel width"); ok = false; }
  T.eq_rows = upload(h, tb->kkt_eq_rows, tb->kkt_n_eq, &ok);
  T.pos_var = upload(h, tb->kkt_pos_var, tb->n, &ok);
  T.pos_eq = upload(h, tb->kkt_pos_eq, tb->kkt_n_eq, &ok);
  T.ksign = upload(h, tb->kkt_sign, tb->kkt_n, &ok);
  T.env_first = upload(h, tb->env_first, (size_t)tb->kkt_n + 1, &ok);
  T.env_ptr = upload(h, tb->env_ptr, (size_t)tb->kkt_n + 2, &ok);
  T.hdst = upload(h, tb->kkt_hdst, tb->nnz_h, &ok);
  T.jdst = upload(h, tb->kkt_jdst, tb->nnz_j, &ok);
  T.kdiag = upload(h, tb->kkt_diag, tb->kkt_n, &ok);
  T.panel_ptr = upload(h, tb->kkt_panel_ptr, (size_t)T.n_panels + 1, &ok);
  T.panel_rows = upload(h, tb->kkt_panel_rows, tb->n_panel_rows, &ok);
  {
    std::vector<int> cmin(T.n_panels > 0 ? T.n_panels : 1);
    for (int pb = 0; pb < T.n_panels; ++pb) {
      int c = pb * NB;
      for (int r = pb * NB; r < tb->kkt_n && r < (pb + 1) * NB; ++r) c = tb->env_first[r] < c ? tb->env_first[r] : c;
      cmin[pb] = c;
    }
    T.panel_cmin = upload(h, cmin.data(), cmin.size(), &ok);
  }
  T.tape_func = upload(h, tb->tape_func, tb->n_tape, &ok);
  T.tape_ptr = upload(h, tb->tape_ptr, (size_t)tb->n_tape + 1, &ok);
  T.tape_coef = upload(h, tb->tape_coef, tb->n_tape_terms, &ok);
  T.tape_fac = upload(h, tb->tape_fac, (size_t)tb->n_tape_terms * 4, &ok);
  T.level_ptr = upload(h, tb->level_ptr, (size_t)tb->n_levels + 1, &ok);
  const bool small_idx = tb->n < 65535 && tb->n_v < 65536 && tb->m < 65535 && tb->nnz_j < 65536;
  if (!small_idx) { set_err("problem too large for 16-bit packed indices"); ok = false; }
  T.G = upload_tl(h, tb->G, tb->n, &ok); T.F = upload_tl(h, tb->F, tb->n, &ok);
  T.DF = upload_tl(h, tb->DF, tb->n, &ok);
  T.J = upload_tl(h, tb->J, tb->n, &ok); T.W = upload_tl(h, tb->W, tb->n, &ok);
  T.jrow = upload(h, tb->jrow, tb->nnz_j, &ok); T.jcol = upload(h, tb->jcol, tb->nnz_j, &ok);
  T.jrow_ptr = upload(h, tb->jrow_ptr, (size_t)tb->m + 1, &ok);
  {  // CSC view of the Jacobian pattern
    std::vector<int> cptr(tb->n + 1, 0), cslot(tb->nnz_j > 0 ? tb->nnz_j : 1);
    for (int s = 0; s < tb->nnz_j; ++s) cptr[tb->jcol[s] + 1]++;
    for (int j = 0; j < tb->n; ++j) cptr[j + 1] += cptr[j];
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (int s = 0; s < tb->nnz_j; ++s) cslot[fill[tb->jcol[s]]++] = s;
    T.jcol_ptr = upload(h, cptr.data(), cptr.size(), &ok);
    T.jcol_slot = upload(h, cslot.data(), (size_t)tb->nnz_j, &ok);
  }
  T.wrow = upload(h, tb->wrow, tb->nnz_w, &ok); T.wcol = upload(h, tb->wcol, tb->nnz_w, &ok);
  T.w2h = upload(h, tb->w2h, tb->nnz_w, &ok);
  T.hrow = upload(h, tb->hrow, tb->nnz_h, &ok); T.hcol = upload(h, tb->hcol, tb->nnz_h, &ok);
  T.hp_ptr = upload(h, tb->hp_ptr, (size_t)tb->nnz_h + 1, &ok);
  T.hp_s1 = upload(h, tb->hp_s1, tb->n_hp, &ok); T.hp_s2 = upload(h, tb->hp_s2, tb->n_hp, &ok);
  T.hp_row = upload(h, tb->hp_row, tb->n_hp, &ok);
  {
    std::vector<unsigned> pack((size_t)tb->n_hp + 1);
    for (int e = 0; e < tb->n_hp; ++e)
      pack[e] = (unsigned)tb->hp_s1[e] | ((unsigned)tb->hp_s2[e] << 16);
    T.hp_pack = upload(h, pack.data(), pack.size(), &ok);
  }
  {  // every diagonal (j,j) must be part of the H pattern (delta_w lands there)
    std::vector<char> has(tb->n, 0);
    for (int q = 0; q < tb->nnz_h; ++q) if (tb->hrow[q] == tb->hcol[q]) has[tb->hrow[q]] = 1;
    for (int j = 0; j < tb->n; ++j) if (!has[j]) {
      set_err("H pattern lacks a diagonal entry (variable without constraint)"); ok = false; break; }
  }
  // shared-memory layout
  Smem& S = h->S;
  const int N = T.N;
  int off = 0;
  auto take = [&](int cnt) { int o = off; off += (cnt + 1) & ~1; return o; };
  S.K = take(T.env_size);
  S.LDP = (T.max_panel_rows + 4 + 3) & ~3;
  S.Pt = take(NB * S.LDP); S.PtS = take(NB * S.LDP);
  S.rbase = take(S.LDP);               // 2*LDP ints: row base offsets + row ids
  S.xe = take(T.n + 1); S.xt = take(T.n + 1); S.dx = take(N + 1); S.u = take(N + 1);
  S.gf = take(T.n);
  S.diag0 = take(N + 1); S.invd = take(N + 1); S.V = take(T.n_v);
  S.red = take(NWARP * NRED); S.filt = take(2 * MAXF);
  S.sgn = take(N); S.eptr = take((N + 2 + 1) / 2); S.efirst = take((N + 1 + 1) / 2);
  S.pptr = take((T.n_panels + 1 + 1) / 2); S.prow = take((tb->n_panel_rows + 1) / 2);
  S.pcmin = take((T.n_panels + 1) / 2);
  S.total = off;
  h->smem_bytes = (size_t)off * sizeof(double);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) ok = false;
  h->n_sm = prop.multiProcessorCount;
  if (ok && h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) {
    char buf[256];
    snprintf(buf, sizeof buf, "KKT block does not fit shared memory: need %zu B, have %zu B (n=%d, n_eq=%d)",
             h->smem_bytes, (size_t)prop.sharedMemPerBlockOptin, T.n, T.n_eq);
    set_err(buf); ok = false;
  }
  // the attribute is per function (shared by all handles): opt in to the device maximum
  if (ok) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, omg_ipm_kernel) != cudaSuccess) { set_err("cudaFuncGetAttributes failed"); ok = false; }
    else if (cudaFuncSetAttribute(omg_ipm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(prop.sharedMemPerBlockOptin - fa.sharedSizeBytes)) != cudaSuccess) {
      set_err(std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(cudaGetLastError())); ok = false; }
  }
  if (ok) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, omg_ipm_kernel, NT, h->smem_bytes);
    h->ctas_per_sm = occ > 0 ? occ : 1;
    h->dscr_stride = 18 * T.m + 2 * T.nnz_j + 8;
    h->iscr_stride = 2 * T.m + T.n_eq + 8;
    if (cudaMalloc(&h->counter, sizeof(int)) != cudaSuccess) ok = false;
    if (cudaMalloc(&h->trace, sizeof(double) * TRACE_ROWS * TRACE_COLS) != cudaSuccess) ok = false;
    else cudaMemset(h->trace, 0, sizeof(double) * TRACE_ROWS * TRACE_COLS);
    cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1);
  }
  if (!ok) { if (g_err.empty()) set_err("device allocation/upload failed"); omg_problem_destroy(h); return nullptr; }
  return h;
}

void omg_problem_destroy(omg_problem* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (void* p : h->allocs) cudaFree(p);
  if (h->dscr) cudaFree(h->dscr);
  if (h->iscr) cudaFree(h->iscr);
  if (h->counter) cudaFree(h->counter);
  if (h->trace) cudaFree(h->trace);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  void* st[] = {h->hx0, h->hp, h->hlb, h->hub, h->hlam0, h->hx, h->hlam, h->hf, h->hst, h->hit};
  for (void* p : st) if (p) cudaFree(p);
  delete h;
}

int omg_set_options(omg_problem* h, const omg_options* opt) {
  if (!h || !opt) { set_err("null argument"); return -1; }
  h->opt = *opt; return 0;
}

int omg_get_info(omg_problem* h, int32_t* n, int32_t* m, int32_t* n_par, int32_t* smem_bytes,
                 int32_t* ctas_per_sm, int32_t* n_sm) {
  if (!h) { set_err("null handle"); return -1; }
  if (n) *n = h->T.n; if (m) *m = h->T.m; if (n_par) *n_par = h->T.n_par;
  if (smem_bytes) *smem_bytes = (int32_t)h->smem_bytes;
  if (ctas_per_sm) *ctas_per_sm = h->ctas_per_sm; if (n_sm) *n_sm = h->n_sm;
  return 0;
}

int omg_solve_batch(omg_problem* h, int32_t B, const double* x0, const double* p,
                    const double* lbg, const double* ubg, int32_t bounds_shared,
                    const double* lam_g0, double* x, double* lam_g, double* f,
                    int32_t* status, int32_t* iters, void* stream_) {
  if (!h) { set_err("null handle"); return -1; }
  if (B <= 0) return 0;
  if (!x0 || !p || !lbg || !ubg || !x || !lam_g || !f || !status || !iters) { set_err("null buffer"); return -1; }
  cudaStream_t stream = (cudaStream_t)stream_;
  CK(cudaSetDevice(h->device));
  int grid = h->n_sm * h->ctas_per_sm;
  if (grid > B) grid = B;
  if (grid > h->scr_ctas) {
    if (h->dscr) cudaFree(h->dscr);
    if (h->iscr) cudaFree(h->iscr);
    h->dscr = nullptr; h->iscr = nullptr;
    const int want = h->n_sm * h->ctas_per_sm;
    CK(cudaMalloc(&h->dscr, (size_t)want * h->dscr_stride * sizeof(double)));
    CK(cudaMalloc(&h->iscr, (size_t)want * h->iscr_stride * sizeof(int)));
    h->scr_ctas = want;
  }
  Batch A;
  A.B = B; A.bounds_shared = bounds_shared;
  A.x0 = x0; A.p = p; A.lbg = lbg; A.ubg = ubg; A.lam0 = lam_g0;
  A.x = x; A.lam = lam_g; A.f = f; A.status = status; A.iters = iters;
  A.dscr = h->dscr; A.iscr = h->iscr; A.dscr_stride = h->dscr_stride; A.iscr_stride = h->iscr_stride;
  A.counter = h->counter; A.trace = h->trace;
  CK(cudaMemsetAsync(h->counter, 0, sizeof(int), stream));
  CK(cudaEventRecord(h->ev0, stream));
  omg_ipm_kernel<<<grid, NT, h->smem_bytes, stream>>>(h->T, h->opt, A, h->S);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, stream));
  h->timed = true; h->launches = 1;
  return 0;
}

int omg_last_timing(omg_problem* h, float* kernel_ms, int32_t* launches) {
  if (!h || !h->timed) { set_err("no solve recorded"); return -1; }
  CK(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (kernel_ms) *kernel_ms = ms;
  if (launches) *launches = h->launches;
  return 0;
}

int omg_get_trace(omg_problem* h, double* out, int32_t max_rows) {
  if (!h || !out) { set_err("null argument"); return -1; }
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  int rows = max_rows < TRACE_ROWS ? max_rows : TRACE_ROWS;
  CK(cudaMemcpy(out, h->trace, sizeof(double) * rows * TRACE_COLS, cudaMemcpyDeviceToHost));
  return rows;
}

static int ensure_staging(omg_problem* h, int B, int shared) {
  if (B <= h->hostB && shared == h->host_shared) return 0;
  void** st[] = {(void**)&h->hx0, (void**)&h->hp, (void**)&h->hlb, (void**)&h->hub, (void**)&h->hlam0,
                 (void**)&h->hx, (void**)&h->hlam, (void**)&h->hf, (void**)&h->hst, (void**)&h->hit};
  for (void** p : st) if (*p) { cudaFree(*p); *p = nullptr; }
  const size_t n = h->T.n, m = h->T.m, np_ = h->T.n_par, b = B;
  CK(cudaMalloc(&h->hx0, b * n * 8)); CK(cudaMalloc(&h->hp, b * (np_ ? np_ : 1) * 8));
  CK(cudaMalloc(&h->hlb, (shared ? 1 : b) * m * 8)); CK(cudaMalloc(&h->hub, (shared ? 1 : b) * m * 8));
  CK(cudaMalloc(&h->hlam0, b * m * 8));
  CK(cudaMalloc(&h->hx, b * n * 8)); CK(cudaMalloc(&h->hlam, b * m * 8)); CK(cudaMalloc(&h->hf, b * 8));
  CK(cudaMalloc(&h->hst, b * 4)); CK(cudaMalloc(&h->hit, b * 4));
  h->hostB = B; h->host_shared = shared;
  return 0;
}

int omg_solve_batch_host(omg_problem* h, int32_t B, const double* x0, const double* p,
                         const double* lbg, const double* ubg, int32_t bounds_shared,
                         const double* lam_g0, double* x, double* lam_g, double* f,
                         int32_t* status, int32_t* iters) {
  if (!h) { set_err("null handle"); return -1; }
  if (B <= 0) return 0;
  CK(cudaSetDevice(h->device));
  if (ensure_staging(h, B, bounds_shared ? 1 : 0)) return -1;
  const size_t n = h->T.n, m = h->T.m, np_ = h->T.n_par, b = B;
  CK(cudaMemcpyAsync(h->hx0, x0, b * n * 8, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(h->hp, p, b * np_ * 8, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(h->hlb, lbg, (bounds_shared ? 1 : b) * m * 8, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(h->hub, ubg, (bounds_shared ? 1 : b) * m * 8, cudaMemcpyHostToDevice, 0));
  if (lam_g0) CK(cudaMemcpyAsync(h->hlam0, lam_g0, b * m * 8, cudaMemcpyHostToDevice, 0));
  if (omg_solve_batch(h, B, h->hx0, h->hp, h->hlb, h->hub, bounds_shared, lam_g0 ? h->hlam0 : nullptr,
                      h->hx, h->hlam, h->hf, h->hst, h->hit, nullptr)) return -1;
  CK(cudaMemcpyAsync(x, h->hx, b * n * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(lam_g, h->hlam, b * m * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(f, h->hf, b * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(status, h->hst, b * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(iters, h->hit, b * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return 0;
}

int omg_shift_batch(omg_problem* h, int32_t B, double* x, int32_t n_blocks, const int32_t* offs,
                    const int32_t* lens, const int32_t* ncols, const double* Tm, void* stream_) {
  if (!h || !x || !offs || !lens || !ncols || !Tm) { set_err("null argument"); return -1; }
  if (B <= 0 || n_blocks <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  CK(cudaSetDevice(h->device));
  std::vector<int> toffs(n_blocks);
  int tot = 0;
  for (int b = 0; b < n_blocks; ++b) { toffs[b] = tot; tot += lens[b] * lens[b]; }
  int *d_i = nullptr; double* d_T = nullptr;
  CK(cudaMalloc(&d_i, sizeof(int) * 4 * n_blocks));
  CK(cudaMalloc(&d_T, sizeof(double) * tot));
  CK(cudaMemcpyAsync(d_i, offs, sizeof(int) * n_blocks, cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(d_i + n_blocks, lens, sizeof(int) * n_blocks, cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(d_i + 2 * n_blocks, ncols, sizeof(int) * n_blocks, cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(d_i + 3 * n_blocks, toffs.data(), sizeof(int) * n_blocks, cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(d_T, Tm, sizeof(double) * tot, cudaMemcpyHostToDevice, stream));
  omg_shift_kernel<<<B, 128, sizeof(double) * h->T.n, stream>>>(x, B, h->T.n, n_blocks, d_i, d_i + n_blocks,
                                                               d_i + 2 * n_blocks, d_i + 3 * n_blocks, d_T);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(stream));
  cudaFree(d_i); cudaFree(d_T);
  return 0;
}

}  // extern "C"
