// omg_b200.cu -- batched primal-dual interior-point solve of OMG-tools' spline
// NLP on B200 (sm_100a).  One thread block per problem instance (persistent
// blocks pull instances from a counter; 2 x 256 threads or 1 x 512 per SM); the
// per-instance state (KKT envelope, Jacobian values, iterate vectors) lives in
// shared memory for the duration of the solve, constant tables stream from L2.
// Kernel variants: omg_ipm_kernel / _2cta (everything in shared memory) and
// omg_ipm_kernel_xl / _xl_2cta (problems with shared intermediates or structures
// larger than one SM's shared memory: tape, chain-rule slots and, if needed, the
// KKT envelope in an L2-resident per-block scratch).
//
// Replaces the CasADi+IPOPT call of the reference (omgtools/problems/
// problem.py:113, optilayer.py:49-60).  Algorithm = oracle/ipm_ref.py (IPOPT
// semantics, Waechter & Biegler 2006); tables = basics/lowering.py.
//
// Design rules (from tools/ubench/lat.cu on B200: DFMA 8 cyc, LDS 29, SHFL.64 27,
// rsqrt 62, __syncthreads 28, dependent LDG ~525 cycles):
//   * no dependent chains through global memory: every pass reads one record
//     per work item (row / column / H position / W slot) and then streams
//     contiguous 32-byte term records with independent loads;
//   * the sequential part of the factorisation (8x8 diagonal blocks, 8x8
//     triangular solves) runs in the registers of a single thread -- no
//     shuffles, little code;
//   * shared-memory placement is adaptive: arrays that do not fit fall back to
//     an L2-resident per-block scratch (generic pointers, same code path).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/omg_b200.h"

// The same source also builds with g++ as a functional CPU emulation of the kernels
// (tools/cpu_emu: its cuda_runtime.h stand-in defines OMG_CPU_EMU and these two macros;
// test infrastructure for a GPU-less container, never loaded by the product).
#ifndef OMG_CPU_EMU
#define OMG_DYN_SHARED(name) extern __shared__ __align__(16) double name[]
#define OMG_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#define NT ((int)blockDim.x)   // threads per block: 512 (1 block/SM) or 256 (2 blocks/SM)
#define NWARP (NT / 32)
#define MAX_NT 512
#define MAX_NWARP (MAX_NT / 32)
#define NB 8              // panel width (== lowering.KKT_NB)
#define MAXF 32           // filter capacity
#define NRED 12           // values per fused block reduction
#define FULL 0xffffffffu
#define TRACE_COLS 8
#define TRACE_ROWS 512
#define MAXW 5            // max x-factors per term record

// ---------------------------------------------------------------------------
// table records (device)
// ---------------------------------------------------------------------------
struct __align__(16) PTerm {       // 32 bytes: coef * V[cidx] * prod x_ext[x[k]]
  double coef;
  unsigned short cidx, aux;        // aux: lambda row (W terms) or slot offset in row (J terms)
  unsigned short x[MAXW];
  unsigned short pad[5];
};
struct __align__(16) RowRec { int g0, g1, jt0, jt1, s0, ns, pad0, pad1; };  // per constraint row
struct __align__(16) HqRec { int dst, p0, p1, diag; };                      // per H position
struct __align__(16) WRec { int dst, t0, t1, pad; };                        // per Hessian slot

enum { A_JVAL = 0, A_SIG, A_JSV, A_G, A_S, A_Y, A_ZU, A_DSC, A_SU, A_DS, A_DY, A_DZU, A_GT,
       A_ST, A_WV, A_ZL, A_SL, A_DZL, A_BEQ, N_ARR };

struct DevTab {
  int n, m, n_par, n_v, n_tape, n_levels, nnz_j, nnz_w, nnz_h, n_eq, N;
  int env_size, n_panels, max_panel_rows, n_panel_rows, n_f;
  const int *tape_func, *tape_ptr, *tape_fac, *level_ptr; const double* tape_coef;
  const PTerm *Gt, *Jt, *Wt, *Ft, *DFt;
  const RowRec* rowrec; const WRec* wrec; const HqRec* hq; const unsigned* hpack;
  const int *dfptr;                     // [n+1] grad-f term ranges per column
  const int *colptr; const unsigned* colrec;   // CSC of J: slot | row << 16
  const unsigned short* jcol16;         // [nnz_j] column of each slot
  const int *eq_rows, *pos_var, *pos_eq, *ksign, *env_first, *env_ptr, *jdst, *kdiag,
            *panel_ptr, *panel_rows, *panel_cmin;
  // intermediates (XL kernel only): G term ranges of the mids, term ranges of the extra
  // J slots (A = d row/d mid, C = d mid/d x), chain-rule pair lists, mu = A^T lambda lists
  int n_mid, nnz_jx, n_hq_heavy;
  const int2* midg; const int* jtptr; const int* jrow;
  const int *jp_ptr, *jp_a, *jp_c, *mu_ptr, *mu_row, *mu_slot;
  const int* jxvar; int n_jxvar;       // extra J slots with an x factor (the others are constant in a solve)
  // J^T Sigma J gather by row chunks staged in shared memory (XL kernel, Jacobian values in scratch):
  // chunk c = the slots [hc_slot[c], hc_slot[c+1]) of whole rows; virtual thread vt owns H positions;
  // its k-th record of chunk c sits at hc_rec[hc_base[c] + k * hc_vt + vt] = {dst, s1 | s2<<16} (chunk-relative)
  int hc_nchunk, hc_vt;
  const int *hc_slot, *hc_base, *hc_cnt; const uint2* hc_rec;
  // extra Hessian slots (mid coefficient depends on x or on another mid): wrec[nnz_w ..
  // nnz_w+nnz_wx) hold the term ranges, xq the H positions that gather
  // sum Wx[e.x] * Jx[e.y] * (e.z >= 0 ? Jx[e.z] : 1)
  int nnz_wx, n_xq;
  const HqRec* xq; const int4* xqp;
};

struct Smem {                      // offsets in doubles
  int K, Pt, PtS, Ld, xe, xt, dx, u, gf, diag0, invd, V, red, filt, rbase, rt8;
  int sgn, eptr, efirst, pptr, prow, pcmin;
  int arr[N_ARR];                  // >= 0: shared offset; < 0: -(scratch offset + 1)
  int LDP, total;
  int hst, hch;                    // staging buffer of the chunked H gather: 2 x hch doubles at hst
  int Kg, Vg, jxg, mug, Kcg, wxg;  // XL kernel: scratch offsets (K only if S.K < 0)
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
  double alpha, theta, phi, f, fsc;
  int n_eq, n_bounds, iter, status, fail, eq_fail, first_try, nfilt, inst, n_restart;
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
#define SOFT_RESTO_FACTOR 0.9999
#define DBL_EPS 2.220446049250313e-16

enum { OP_MAX = 0, OP_MIN = 1, OP_SUM = 2 };

// phase timers (debug, opt.trace=1, instance 0): cycles per phase summed over the solve
#define NPHASE 16
#define TICK(k) do { if (tracing && threadIdx.x == 0) { const long long t_ = clock64(); \
  phase_cyc[k] += (double)(t_ - phase_t0); phase_t0 = t_; } } while (0)

__device__ __forceinline__ double term_value(const PTerm* __restrict__ p, const double* __restrict__ V,
                                             const double* __restrict__ xe, int* aux) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  double v = __hiloint2double((int)a.y, (int)a.x) * V[a.z & 0xffffu];
  *aux = (int)(a.z >> 16);
  v *= xe[a.w & 0xffffu]; v *= xe[a.w >> 16];
  v *= xe[b.x & 0xffffu]; v *= xe[b.x >> 16];
  v *= xe[b.y & 0xffffu];
  return v;
}

__device__ __forceinline__ double eval_range(const PTerm* __restrict__ t, int lo, int hi,
                                             const double* __restrict__ V, const double* __restrict__ xe) {
  double acc = 0.0;
  int aux;
#pragma unroll 4
  for (int k = lo; k < hi; ++k) acc += term_value(t + k, V, xe, &aux);
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

OMG_DYN_SHARED(sm);

// shared-memory copies of the KKT structure arrays, addressed from the block's
// dynamic shared array so the compiler emits LDS (not generic loads)
#define KS_SGN    (sm + S.sgn)        // pivot signs S of K = L S L^T (rewritten by every factorisation)
#define KS_EPTR   (reinterpret_cast<const int*>(sm + S.eptr))
#define KS_EFIRST (reinterpret_cast<const int*>(sm + S.efirst))
#define KS_PPTR   (reinterpret_cast<const int*>(sm + S.pptr))
#define KS_PROW   (reinterpret_cast<const int*>(sm + S.prow))
#define KS_PCMIN  (reinterpret_cast<const int*>(sm + S.pcmin))

// ---------------------------------------------------------------------------
// Blocked right-looking factorisation K = L S L^T on envelope storage.
//   row i (permuted order) is stored from column first[i] (multiple of NB) to i
//   at K[env_ptr[i] + j - first[i]]; row N is the right-hand side (never a
//   pivot), so after the sweep it holds S L^{-1} r.
// Per 8-column panel: (1) ONE thread factors the 8x8 diagonal block in
// registers, (2) one thread per reached row does the panel solve, (3) the
// rank-8 trailing update runs over 2x2 register tiles of the reached rows.
// Pivot j must satisfy sign[j]*pivot > PIV_TOL*|K_jj| (variables) or > 0
// (equality rows); otherwise ctl->fail (eq_fail for an equality pivot).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void factor_env(const DevTab& T, const Smem& S, double* K, Ctl* ctl, double* pc,
                                           const int mode) {
  const int tid = threadIdx.x;
  const int N = T.N;
  const int LDP = S.LDP;
  double* Pt = sm + S.Pt; double* PtS = sm + S.PtS; double* Ld = sm + S.Ld;
  const double* diag0 = sm + S.diag0; double* invd = sm + S.invd;
  int* rbase = reinterpret_cast<int*>(sm + S.rbase);
  int* rrow = rbase + LDP;
  double* sgn = KS_SGN; const int* eptr = KS_EPTR; const int* efirst = KS_EFIRST;
  int n_neg = 0;               // negative pivots so far (thread 0)
  const int* pptr = KS_PPTR; const int* prow = KS_PROW;
  long long t0 = clock64();
#define FT(k) do { if (pc && tid == 0) { const long long t_ = clock64(); pc[k] += (double)(t_ - t0); t0 = t_; } } while (0)
  for (int pb = 0; pb < T.n_panels; ++pb) {
    const int kb = pb * NB;
    const int nb = min(NB, N - kb);
    // ---- 1. diagonal block: staged into Ld (64 threads), factored by ONE thread in
    //         registers (branch-free pivots), scattered back by 64 threads ----------
    if (tid < NB * NB) {
      const int r = tid >> 3, c = tid & 7;
      double v = 0.0;
      if (r < nb && c <= r) v = K[eptr[kb + r] + kb - efirst[kb + r] + c];
      Ld[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double a[NB][NB];
#pragma unroll
      for (int r = 0; r < NB; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = Ld[r * NB + c];
      int bad_at = -1;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        // rows >= nb are zero-padded: give them a unit pivot so the arithmetic stays finite
        const bool live = (j < nb);
        // mode 0 (IPOPT's inertia test): S takes the sign of the pivot as it comes, only the
        // number of negative pivots is checked at the end; mode 1: S fixed (+ variables,
        // - equality rows), a pivot of the other sign fails
        // (|pivot| keeps the sign selection off the dependent chain pivot -> rsqrt -> column)
        const double sj = !live ? 1.0 : (mode ? sgn[kb + j] : (a[j][j] > 0.0 ? 1.0 : -1.0));
        const double d = !live ? 1.0 : (mode ? sj * a[j][j] : fabs(a[j][j]));
        const double thr = (live && (mode == 0 || sj > 0.0)) ? PIV_TOL * fmax(diag0[kb + j], 1e-300) : 0.0;
        const bool fail_j = !(d > thr) || !(d < 1e300);
        if (fail_j && bad_at < 0) bad_at = kb + j;
        if (live) { sgn[kb + j] = sj; n_neg += (sj < 0.0) ? 1 : 0; }
        const double inv = rsqrt(d);
        a[j][j] = d * inv;
        if (live) invd[kb + j] = inv;
        const double f = inv * sj;
#pragma unroll
        for (int r = j + 1; r < NB; ++r) a[r][j] *= f;
#pragma unroll
        for (int r = j + 1; r < NB; ++r) {
          const double lr = sj * a[r][j];
#pragma unroll
          for (int c = j + 1; c <= r; ++c) a[r][c] -= lr * a[c][j];
        }
      }
      if (bad_at >= 0) { ctl->fail = 1; ctl->eq_fail = (T.ksign[bad_at] < 0) ? 1 : 0; }
      else if (mode == 0 && n_neg > T.n_eq) { ctl->fail = 1; ctl->eq_fail = 0; }   // too many already: stop early
#pragma unroll
      for (int r = 0; r < NB; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) Ld[r * NB + c] = a[r][c];
    }
    __syncthreads();
    if (!ctl->fail && tid < NB * NB) {
      const int r = tid >> 3, c = tid & 7;
      if (r < nb && c <= r) K[eptr[kb + r] + kb - efirst[kb + r] + c] = Ld[tid];
    }
    __syncthreads();
    FT(7);
    if (ctl->fail) return;
    // ---- 2. panel solve over the rows this panel reaches ----------------------
    const int p0 = pptr[pb];
    const int nrows = pptr[pb + 1] - p0;
    if (tid < nrows) {
      const int rr = tid;
      const int r = prow[p0 + rr];
      const int rb = eptr[r] - efirst[r];
      rbase[rr] = rb; rrow[rr] = r;
      double* Kr = K + rb + kb;
      double a[NB], sg[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) { a[c] = (c < nb) ? Kr[c] : 0.0; sg[c] = (c < nb) ? sgn[kb + c] : 1.0; }
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        if (c < nb) {
          double v = a[c];
#pragma unroll
          for (int j = 0; j < c; ++j) v -= sg[j] * a[j] * Ld[c * NB + j];
          a[c] = v * invd[kb + c] * sg[c];
        }
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) if (c < nb) Kr[c] = a[c];
      // panel buffers as double2 [column pair][row]: conflict-free stores here, and
      // at most 2-way conflicts for the 2x2-tile loads of the trailing update
      double2* P2 = reinterpret_cast<double2*>(Pt);
      double2* PS2 = reinterpret_cast<double2*>(PtS);
#pragma unroll
      for (int q = 0; q < NB / 2; ++q) {
        const double v0 = (2 * q < nb) ? a[2 * q] : 0.0, v1 = (2 * q + 1 < nb) ? a[2 * q + 1] : 0.0;
        P2[q * LDP + rr] = make_double2(v0, v1);
        PS2[q * LDP + rr] = make_double2(sg[2 * q] * v0, sg[2 * q + 1] * v1);
      }
    } else if (tid < nrows + 2 && tid < LDP) {   // zero padding rows for the 2x2 tiles
      double2* P2 = reinterpret_cast<double2*>(Pt);
      double2* PS2 = reinterpret_cast<double2*>(PtS);
#pragma unroll
      for (int q = 0; q < NB / 2; ++q) { P2[q * LDP + tid] = make_double2(0.0, 0.0); PS2[q * LDP + tid] = make_double2(0.0, 0.0); }
    }
    __syncthreads();
    FT(8);
    // ---- 3. trailing update: 2x2 tiles over list positions (a >= b) -----------
    const int nt2 = (nrows + 1) >> 1;
    const int W = nt2 + 1, H2 = (nt2 + 1) >> 1;
    for (int t = tid; t < H2 * W; t += NT) {
      const int ra = t / W, cb = t - ra * W;
      int ta, tb;
      if (cb <= ra) { ta = ra; tb = cb; }
      else { ta = nt2 - 1 - ra; tb = cb - ra - 1; if (ta == ra) continue; }
      const int a0 = ta * 2, b0 = tb * 2;
      const double2* A0 = reinterpret_cast<const double2*>(Pt) + a0;
      const double2* B0 = reinterpret_cast<const double2*>(PtS) + b0;
      double acc00 = 0.0, acc01 = 0.0, acc10 = 0.0, acc11 = 0.0;
#pragma unroll
      for (int q = 0; q < NB / 2; ++q) {
        const double2 x0 = A0[q * LDP], x1 = A0[q * LDP + 1], y0 = B0[q * LDP], y1 = B0[q * LDP + 1];
        acc00 += x0.x * y0.x; acc00 += x0.y * y0.y;
        acc01 += x0.x * y1.x; acc01 += x0.y * y1.y;
        acc10 += x1.x * y0.x; acc10 += x1.y * y0.y;
        acc11 += x1.x * y1.x; acc11 += x1.y * y1.y;
      }
      // scatter into K (only b <= a, column < N)
      const int la0 = a0, la1 = a0 + 1, lb0 = b0, lb1 = b0 + 1;
      const int c0 = rrow[lb0];
      const int c1 = (lb1 < nrows) ? rrow[lb1] : N;
      if (la0 < nrows) {
        const int rb = rbase[la0];
        if (lb0 <= la0 && c0 < N) K[rb + c0] -= acc00;
        if (lb1 <= la0 && c1 < N) K[rb + c1] -= acc01;
      }
      if (la1 < nrows) {
        const int rb = rbase[la1];
        if (lb0 <= la1 && c0 < N) K[rb + c0] -= acc10;
        if (lb1 <= la1 && c1 < N) K[rb + c1] -= acc11;
      }
    }
    __syncthreads();
    FT(9);
  }
#undef FT
  if (mode == 0 && tid == 0 && !ctl->fail && n_neg != T.n_eq) {   // Sylvester: wrong inertia
    ctl->fail = 1; ctl->eq_fail = (n_neg < T.n_eq) ? 1 : 0;
  }
}

// Back substitution L^T u = w on envelope storage (w = row N of L on entry).
__device__ __forceinline__ void back_solve_env(const DevTab& T, const Smem& S, const double* K) {
  const int tid = threadIdx.x;
  const int N = T.N;
  const double* invd = sm + S.invd; double* w = sm + S.u;
  const int* eptr = KS_EPTR; const int* efirst = KS_EFIRST; const int* pcmin = KS_PCMIN;
  for (int pb = T.n_panels - 1; pb >= 0; --pb) {
    const int kb = pb * NB;
    const int nb = min(NB, N - kb);
    // per-row bases of the block (uniform; 16 independent shared loads)
    int rbs[NB], rfs[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int rj = min(kb + j, N - 1);
      rfs[j] = efirst[rj]; rbs[j] = eptr[rj] - rfs[j];
    }
    if (tid == 0) {        // 8x8 upper-triangular solve in registers
      double wv[NB], dinv[NB], L[NB][NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        wv[j] = (j < nb) ? w[kb + j] : 0.0;
        dinv[j] = (j < nb) ? invd[kb + j] : 1.0;
#pragma unroll
        for (int c = 0; c < j; ++c) L[j][c] = (j < nb) ? K[rbs[j] + kb + c] : 0.0;
      }
#pragma unroll
      for (int j = NB - 1; j >= 0; --j) {
        const double uj = wv[j] * dinv[j];
        wv[j] = uj;
#pragma unroll
        for (int c = 0; c < j; ++c) wv[c] -= L[j][c] * uj;
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) if (j < nb) w[kb + j] = wv[j];
    }
    __syncthreads();
    const int cmin = pcmin[pb];
    for (int c = cmin + tid; c < kb; c += NT) {
      double l[NB], uu[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const bool in = (j < nb) && (c >= rfs[j]);
        l[j] = in ? K[rbs[j] + c] : 0.0;
        uu[j] = (j < nb) ? w[kb + j] : 0.0;
      }
      double acc0 = w[c], acc1 = 0.0;
#pragma unroll
      for (int j = 0; j < NB; j += 2) { acc0 -= l[j] * uu[j]; acc1 -= l[j + 1] * uu[j + 1]; }
      w[c] = acc0 + acc1;
    }
    __syncthreads();
  }
}

// Constraint Jacobian of the XL kernel, one thread per slot (rows differ widely in size).
// First the extra slots jx[s], s in [nnz_j, nnz_jx): A = d row/d mid (parameter-only) and
// C = d mid/d x; then every constraint slot = direct terms + chain rule
// sum_e A[jp_a[e]] * C[jp_c[e]], scaled by the row scaling dsc (nullptr: unscaled).
// Ends with a block barrier.
__device__ __forceinline__ void jac_xl(const DevTab& T, const double* __restrict__ V,
                                       const double* __restrict__ xe, double* jx, double* jval,
                                       const double* dsc, const bool first) {
  const int tid = threadIdx.x;
  if (first) {   // every extra slot; the parameter-only ones (A of config 4) keep this value
    for (int s = T.nnz_j + tid; s < T.nnz_jx; s += NT)
      jx[s] = eval_range(T.Jt, T.jtptr[s], T.jtptr[s + 1], V, xe);
  } else {
    for (int q = tid; q < T.n_jxvar; q += NT) {
      const int s = T.jxvar[q];
      jx[s] = eval_range(T.Jt, T.jtptr[s], T.jtptr[s + 1], V, xe);
    }
  }
  __syncthreads();
  for (int s = tid; s < T.nnz_j; s += NT) {
    double acc = eval_range(T.Jt, T.jtptr[s], T.jtptr[s + 1], V, xe);
    if (T.n_mid)
      for (int e = T.jp_ptr[s]; e < T.jp_ptr[s + 1]; ++e) acc += jx[T.jp_a[e]] * jx[T.jp_c[e]];
    jval[s] = dsc ? dsc[T.jrow[s]] * acc : acc;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// the solver kernel
// ---------------------------------------------------------------------------
// XL = true: problems with intermediates (n_mid > 0) and/or a KKT envelope / parameter tape
// that exceeds shared memory: V (and K if needed) live in the block's L2-resident scratch.
template <bool XL>
__device__ __forceinline__ void ipm_body(const DevTab& T, const omg_options& O, const Batch& A, const Smem& S) {
  __shared__ Ctl ctl;
  __shared__ double phase_cyc[NPHASE];
  double* const Dx = A.dscr + (size_t)blockIdx.x * A.dscr_stride;
  double* K = (XL && S.K < 0) ? Dx + S.Kg : sm + S.K;
  double* u = sm + S.u;
  double* xe = sm + S.xe;
  double* xt = sm + S.xt;
  double* dx = sm + S.dx;
  double* gf = sm + S.gf;
  double* diag0 = sm + S.diag0;
  double* V = (XL && S.V < 0) ? Dx + S.Vg : sm + S.V;
  double* jx = XL ? Dx + S.jxg - T.nnz_j : nullptr;   // indexed by slot id >= nnz_j
  double* mu_mid = XL ? Dx + S.mug : nullptr;
  double* wx = XL ? Dx + S.wxg : nullptr;        // values of the cross-Hessian slots
  double* Kc = Dx + S.Kcg;
  const int n_xe = T.n + 1 + (XL ? T.n_mid : 0);
  double* red = sm + S.red;
  double* filt = sm + S.filt;
  unsigned char* rt = reinterpret_cast<unsigned char*>(sm + S.rt8);
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
    for (int i = tid; i < T.n_panel_rows; i += NT) prow[i] = T.panel_rows[i];
    for (int i = tid; i < T.n_panels; i += NT) pcmin[i] = T.panel_cmin[i];
  }
  __syncthreads();

  // per-instance arrays: shared memory if they fit, else L2-resident scratch
  double* D = A.dscr + (size_t)blockIdx.x * A.dscr_stride;
#define ARR(k) ((S.arr[k] >= 0) ? (sm + S.arr[k]) : (D + (-(S.arr[k]) - 1)))
  double* jval = ARR(A_JVAL); double* sig = ARR(A_SIG); double* jsv = ARR(A_JSV);
  double* g = ARR(A_G); double* s = ARR(A_S); double* y = ARR(A_Y); double* zU = ARR(A_ZU);
  double* dsc = ARR(A_DSC); double* sU = ARR(A_SU); double* ds = ARR(A_DS); double* dy = ARR(A_DY);
  double* dzU = ARR(A_DZU); double* gt = ARR(A_GT); double* st = ARR(A_ST); double* wv = ARR(A_WV);
  double* zL = ARR(A_ZL); double* sL = ARR(A_SL); double* dzL = ARR(A_DZL); double* beq = ARR(A_BEQ);
#undef ARR
  int* I = A.iscr + (size_t)blockIdx.x * A.iscr_stride;
  int* eqidx = I;           int* eqrow = eqidx + m;

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
          const int4 f = __ldg(reinterpret_cast<const int4*>(T.tape_fac) + t);
          acc += T.tape_coef[t] * V[f.x] * V[f.y] * V[f.z] * V[f.w];
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
    if (XL) {
      for (int l = tid; l < T.n_mid; l += NT) { const int2 r = T.midg[l]; xe[n + 1 + l] = eval_range(T.Gt, r.x, r.y, V, xe); }
      __syncthreads();
      jac_xl(T, V, xe, jx, jval, nullptr, true);
    }

    // ---- S3: row classification, scaling, starting point -----------------------
    double fmaxv = 0.0;
    for (int j = tid; j < n; j += NT)
      fmaxv = fmax(fmaxv, fabs(eval_range(T.DFt, T.dfptr[j], T.dfptr[j + 1], V, xe)));
    {
      double r1[1] = {fmaxv}; const int o1[1] = {OP_MAX};
      block_reduce<1>(r1, o1, red);
      fmaxv = r1[0];
    }
    const double smg = O.scaling_max_gradient;
    const double fsc = (fmaxv > smg) ? fmax(smg / fmaxv, 1e-8) : 1.0;
    for (int i = tid; i < m; i += NT) {
      const RowRec rr = T.rowrec[i];
      // Jacobian row: max |J| for the gradient-based scaling
      double gm = 0.0;
      if (XL) {       // jval = unscaled Jacobian (jac_xl)
        for (int k = 0; k < rr.ns; ++k) gm = fmax(gm, fabs(jval[rr.s0 + k]));
      } else {
        double acc = 0.0; int cur = 0, aux;
        for (int k = rr.jt0; k < rr.jt1; ++k) {
          const double v = term_value(T.Jt + k, V, xe, &aux);
          if (aux != cur) { gm = fmax(gm, fabs(acc)); acc = 0.0; cur = aux; }
          acc += v;
        }
        gm = fmax(gm, fabs(acc));
      }
      const double d = (gm > smg) ? fmax(smg / gm, 1e-8) : 1.0;
      dsc[i] = d;
      const double lb = lbg[i], ub = ubg[i];
      const bool eq = (lb == ub);
      const bool hL = (lb > -INF_BOUND) && !eq, hU = (ub < INF_BOUND) && !eq;
      rt[i] = (unsigned char)((hL ? 1 : 0) | (hU ? 2 : 0) | (eq ? 4 : 0));
      double l = lb * d, uu = ub * d;
      beq[i] = l;
      if (hL) l -= O.bound_relax_factor * fmax(1.0, fabs(l));
      if (hU) uu += O.bound_relax_factor * fmax(1.0, fabs(uu));
      sL[i] = l; sU[i] = uu;
      const double gi = d * eval_range(T.Gt, rr.g0, rr.g1, V, xe);
      g[i] = gi;
      double si = gi;
      const double k1 = O.bound_push, k2 = O.bound_frac;
      double pl = k1 * fmax(1.0, fabs(l)), pu = k1 * fmax(1.0, fabs(uu));
      if (hL && hU) { pl = fmin(pl, k2 * (uu - l)); pu = fmin(pu, k2 * (uu - l)); }
      if (hL) si = fmax(si, l + pl);
      if (hU) si = fmin(si, uu - pu);
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
      ctl.fsc = fsc; ctl.alpha = 0.0; ctl.delta_w = 0.0; ctl.n_restart = 0;
      ctl.f = fsc * eval_range(T.Ft, 0, T.n_f, V, xe);
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
      // ---- I1: rows: Jacobian values, residual terms (g kept from the trial) ------
      double rv[NRED];
      // 0 cinf(max) 1 maxprod(max) 2 minprod(min) 3 viol(max) 4 rsinf(max)
      // 5 rsinf_un(max) 6 ysum 7 zsum 8 theta 9 logsum 10 rxinf(max)
      const int rop[NRED] = {OP_MAX, OP_MAX, OP_MIN, OP_MAX, OP_MAX, OP_MAX,
                             OP_SUM, OP_SUM, OP_SUM, OP_SUM, OP_MAX, OP_SUM};
      for (int r = 0; r < NRED; ++r) rv[r] = 0.0;
      rv[2] = 1e300;
      if (XL) jac_xl(T, V, xe, jx, jval, dsc, false);
      for (int i = tid; i < m; i += NT) {
        const RowRec rr = T.rowrec[i];
        const int r = rt[i];
        const double d = dsc[i];
        if (!XL) {
          double acc = 0.0; int cur = 0, aux;
          double* jv = jval + rr.s0;
#pragma unroll 4
          for (int k = rr.jt0; k < rr.jt1; ++k) {
            const double v = term_value(T.Jt + k, V, xe, &aux);
            if (aux != cur) { jv[cur] = d * acc; acc = 0.0; cur = aux; }
            acc += v;
          }
          if (rr.ns > 0) jv[cur] = d * acc;
        }
        const double gi = g[i], si = s[i], yi = y[i];
        const double ci = (r & 4) ? gi - beq[i] : gi - si;
        rv[0] = fmax(rv[0], fabs(ci));
        rv[8] += fabs(ci);
        double zl = 0.0, zu = 0.0;
        if (r & 1) { zl = zL[i]; const double dl = si - sL[i]; const double pz = dl * zl;
          rv[1] = fmax(rv[1], pz); rv[2] = fmin(rv[2], pz); rv[9] += log(dl); rv[7] += zl; }
        if (r & 2) { zu = zU[i]; const double du = sU[i] - si; const double pz = du * zu;
          rv[1] = fmax(rv[1], pz); rv[2] = fmin(rv[2], pz); rv[9] += log(du); rv[7] += zu; }
        const double gun = gi / d;
        if (r & 6) rv[3] = fmax(rv[3], gun - ubg[i]);
        if (r & 5) rv[3] = fmax(rv[3], lbg[i] - gun);
        if (!(r & 4)) { const double rs = fabs(-yi - zl + zu);
          rv[4] = fmax(rv[4], rs); rv[5] = fmax(rv[5], rs * d); }
        rv[6] += fabs(yi);
      }
      __syncthreads();   // jval visible to the column pass
      TICK(1);   // row pass
      // ---- I2: columns: grad f, dual residual ------------------------------------
      for (int j = tid; j < n; j += NT) {
        const double gj = ctl.fsc * eval_range(T.DFt, T.dfptr[j], T.dfptr[j + 1], V, xe);
        gf[j] = gj;
        double rx = gj;
        const int q0 = T.colptr[j], q1 = T.colptr[j + 1];
#pragma unroll 4
        for (int q = q0; q < q1; ++q) {
          const unsigned cr = __ldg(T.colrec + q);
          rx += jval[cr & 0xffffu] * y[cr >> 16];
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
      if (!isfinite(E0) || !isfinite(theta)) status = OMG_INVALID_NUMBER_DETECTED;
      else if (E0 <= O.tol && dinf_un <= O.dual_inf_tol && viol <= O.constr_viol_tol &&
               cmpl0 / ctl.fsc <= O.compl_inf_tol) status = OMG_SOLVE_SUCCEEDED;
      else if (iter >= O.max_iter) status = OMG_MAX_ITER_EXCEEDED;
      if (tracing && tid == 0 && iter < TRACE_ROWS - 2) {
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
      // ---- I4: Sigma, w = Sigma r_d + phi_s, sigma-scaled Jacobian ----------------
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
        const RowRec rr = T.rowrec[i];
        for (int k = 0; k < rr.ns; ++k) jsv[rr.s0 + k] = sg * jval[rr.s0 + k];
      }
      __syncthreads();
      TICK(4);   // sigma pass

      if (XL) {   // multipliers of the mids' Hessians: mu = A^T (y*dsc), A raw (unscaled)
        for (int l = tid; l < T.n_mid; l += NT) {
          double acc = 0.0;
          for (int e = T.mu_ptr[l]; e < T.mu_ptr[l + 1]; ++e) {
            const int i = T.mu_row[e];
            acc += y[i] * dsc[i] * jx[T.mu_slot[e]];
          }
          mu_mid[l] = acc;
        }
        __syncthreads();
      }
      // ---- I7/I8: assemble + factorise, with inertia correction -----------------
      bool assembled = false;   // the assembled K (delta = 0) is kept in scratch for the retries
      for (;;) {
        if (assembled) {
          const double2* C2 = reinterpret_cast<const double2*>(Kc);
          double2* K2 = reinterpret_cast<double2*>(K);
          const int h2 = (T.env_size + 1) >> 1;
          for (int q = tid; q < h2; q += NT) K2[q] = C2[q];
          __syncthreads();
          for (int j = tid; j < n; j += NT) {
            const int pj = T.pos_var[j];
            const double v = K[T.kdiag[pj]] + ctl.delta_w;
            K[T.kdiag[pj]] = v; diag0[pj] = fabs(v);
          }
          for (int k = tid; k < n_eq; k += NT) {
            const int pk = T.pos_eq[k];
            K[T.kdiag[pk]] = -ctl.delta_c; diag0[pk] = ctl.delta_c;
          }
          if (tid == 0) { ctl.fail = 0; ctl.eq_fail = 0; }
          __syncthreads();
        } else {
        {
          double2* K2 = reinterpret_cast<double2*>(K);
          const int h2 = (T.env_size + 1) >> 1;
          for (int q = tid; q < h2; q += NT) K2[q] = make_double2(0.0, 0.0);
        }
        __syncthreads();
        // H positions: gather J^T Sigma J (+ delta_w on the diagonal)
        if (XL && T.hc_nchunk > 0) {
          // the Jacobian values live in the L2-resident scratch: a pair would cost two scattered
          // 8-byte loads (32-byte sectors).  Instead the rows are staged chunk by chunk into the
          // panel buffers (free until the factorisation) with coalesced loads, and every thread
          // adds the pairs of ITS H positions that fall into the chunk, from a coalesced record
          // stream (positions are owned by one virtual thread, so no atomics).
          double* Ach = sm + S.hst; double* Bch = Ach + S.hch;
          for (int c = 0; c < T.hc_nchunk; ++c) {
            const int s0 = T.hc_slot[c], nsl = T.hc_slot[c + 1] - s0;
            for (int k = tid; k < nsl; k += NT) { Ach[k] = jval[s0 + k]; Bch[k] = jsv[s0 + k]; }
            __syncthreads();
            const uint2* rp = T.hc_rec + (size_t)T.hc_base[c];
            for (int vt = tid; vt < T.hc_vt; vt += NT) {
              const int cnt = T.hc_cnt[c * T.hc_vt + vt];
              unsigned cur = 0xffffffffu;
              double acc = 0.0;
              for (int k0 = 0; k0 < cnt; k0 += 8) {        // eight record loads in flight
                uint2 r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) if (k0 + u < cnt) r[u] = __ldg(rp + (size_t)(k0 + u) * T.hc_vt + vt);
#pragma unroll
                for (int u = 0; u < 8; ++u) if (k0 + u < cnt) {
                  if (r[u].x != cur) { if (cur != 0xffffffffu) K[cur] += acc; cur = r[u].x; acc = 0.0; }
                  acc += Bch[r[u].y & 0xffffu] * Ach[r[u].y >> 16];
                }
              }
              if (cur != 0xffffffffu) K[cur] += acc;
            }
            __syncthreads();
          }
          for (int j = tid; j < n; j += NT) K[T.kdiag[T.pos_var[j]]] += ctl.delta_w;
        } else
        for (int q = tid; q < T.nnz_h; q += NT) {
          const HqRec h = T.hq[q];
          double acc = 0.0;
#pragma unroll 4
          for (int e = h.p0; e < h.p1; ++e) {
            const unsigned pk = __ldg(T.hpack + e);
            acc += jsv[pk & 0xffffu] * jval[pk >> 16];
          }
          if (h.diag) acc += ctl.delta_w;
          K[h.dst] = acc;
        }
        __syncthreads();
        TICK(5);   // zero + H gather
        // Lagrangian Hessian W (lambda = y*dsc, objective factor fsc)
        if (XL) {   // slots differ widely in term count: one warp per slot
          const int lane = tid & 31;
          for (int q = tid >> 5; q < T.nnz_w + T.nnz_wx; q += NWARP) {
            const WRec w = T.wrec[q];
            double acc = 0.0;
            for (int t = w.t0 + lane; t < w.t1; t += 32) {
              int lr;
              double v = term_value(T.Wt + t, V, xe, &lr);
              v *= (lr < m) ? (y[lr] * dsc[lr]) : (lr == m ? ctl.fsc : mu_mid[lr - m - 1]);
              acc += v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL, acc, o);
            if (lane == 0) { if (q < T.nnz_w) K[w.dst] += acc; else wx[w.dst] = acc; }
          }
          if (T.nnz_wx) {   // X^T C + C^T X + C^T M C, one thread per H position (deterministic)
            __syncthreads();
            for (int e = tid; e < T.n_xq; e += NT) {
              const HqRec h = T.xq[e];
              double acc = 0.0;
              for (int r = h.p0; r < h.p1; ++r) {
                const int4 pr = T.xqp[r];
                const double v = wx[pr.x] * jx[pr.y];
                acc += (pr.z >= 0) ? v * jx[pr.z] : v;
              }
              K[h.dst] += acc;
            }
          }
        } else
        for (int q = tid; q < T.nnz_w; q += NT) {
          const WRec w = T.wrec[q];
          double acc = 0.0;
          for (int t = w.t0; t < w.t1; ++t) {
            int lr;
            double v = term_value(T.Wt + t, V, xe, &lr);
            if (XL) v *= (lr < m) ? (y[lr] * dsc[lr]) : (lr == m ? ctl.fsc : mu_mid[lr - m - 1]);
            else v *= (lr < m) ? (y[lr] * dsc[lr]) : ctl.fsc;
            acc += v;
          }
          K[w.dst] += acc;
        }
        __syncthreads();
        for (int j = tid; j < n; j += NT) {
          const int pj = T.pos_var[j];
          diag0[pj] = fabs(K[T.kdiag[pj]]);
        }
        // equality border + right-hand-side row
        const int rhs0 = KS_EPTR[N];
        for (int k = tid; k < n_eq; k += NT) {
          const int i = eqrow[k], pk = T.pos_eq[k];
          const RowRec rr = T.rowrec[i];
          for (int q = 0; q < rr.ns; ++q) K[T.jdst[rr.s0 + q]] = jval[rr.s0 + q];
          K[T.kdiag[pk]] = -ctl.delta_c;
          diag0[pk] = ctl.delta_c;
          K[rhs0 + pk] = -(g[i] - beq[i]);
        }
        for (int j = tid; j < n; j += NT) {
          double acc = gf[j];
          const int q0 = T.colptr[j], q1 = T.colptr[j + 1];
#pragma unroll 4
          for (int q = q0; q < q1; ++q) {
            const unsigned cr = __ldg(T.colrec + q);
            acc += jval[cr & 0xffffu] * wv[cr >> 16];
          }
          K[rhs0 + T.pos_var[j]] = -acc;
        }
        if (tid == 0) { ctl.fail = 0; ctl.eq_fail = 0; }
        __syncthreads();
        {
          double2* C2 = reinterpret_cast<double2*>(Kc);
          const double2* K2 = reinterpret_cast<const double2*>(K);
          const int h2 = (T.env_size + 1) >> 1;
          for (int q = tid; q < h2; q += NT) C2[q] = K2[q];
          assembled = true;
        }
        }
        TICK(6);   // W + border + rhs
        factor_env(T, S, K, &ctl, tracing ? phase_cyc : nullptr, O.inertia_mode);
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
      for (int j = tid; j < N; j += NT) u[j] = K[KS_EPTR[N] + j];
      __syncthreads();
      back_solve_env(T, S, K);
      for (int j = tid; j < n; j += NT) dx[j] = u[T.pos_var[j]];
      for (int k = tid; k < n_eq; k += NT) dx[n + k] = u[T.pos_eq[k]];
      __syncthreads();
      TICK(10);  // back substitution
      // ---- I10: ds, dy, dz, fraction to the boundary --------------------------
      double sv[4];   // 0 a_p(min) 1 a_d(min) 2 gphi(sum)
      const int sop[4] = {OP_MIN, OP_MIN, OP_SUM, OP_SUM};
      sv[0] = 1.0; sv[1] = 1.0; sv[2] = 0.0; sv[3] = 0.0;
      for (int i = tid; i < m; i += NT) {
        const int r = rt[i];
        const RowRec rr = T.rowrec[i];
        double jd = 0.0;
#pragma unroll 4
        for (int k = 0; k < rr.ns; ++k) jd += jval[rr.s0 + k] * dx[T.jcol16[rr.s0 + k]];
        if (r & 4) {
          ds[i] = 0.0; dy[i] = dx[n + eqidx[i]]; dzL[i] = 0.0; dzU[i] = 0.0;
        } else {
          const double si = s[i];
          const double dsi = jd + (g[i] - si);
          double ph = 0.0, a = 0.0, b = 0.0;
          if (r & 1) { const double dl = si - sL[i]; const double z = zL[i]; ph -= mu / dl;
            a = mu / dl - z - (z / dl) * dsi;
            if (dsi < 0.0) sv[0] = fmin(sv[0], -tau * dl / dsi);
            if (a < 0.0) sv[1] = fmin(sv[1], -tau * z / a); dzL[i] = a; }
          if (r & 2) { const double du = sU[i] - si; const double z = zU[i]; ph += mu / du;
            b = mu / du - z + (z / du) * dsi;
            if (dsi > 0.0) sv[0] = fmin(sv[0], tau * du / dsi);
            if (b < 0.0) sv[1] = fmin(sv[1], -tau * z / b); }
          ds[i] = dsi; dzU[i] = b;
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
        if (XL) {
          for (int l = tid; l < T.n_mid; l += NT) { const int2 r = T.midg[l]; xt[n + 1 + l] = eval_range(T.Gt, r.x, r.y, V, xt); }
          __syncthreads();
        }
        double tv[3];  // 0 theta 1 logsum 2 f
        const int top[3] = {OP_SUM, OP_SUM, OP_SUM};
        tv[0] = 0.0; tv[1] = 0.0; tv[2] = 0.0;
        for (int i = tid; i < m; i += NT) {
          const RowRec rr = T.rowrec[i];
          const int r = rt[i];
          const double gi = dsc[i] * eval_range(T.Gt, rr.g0, rr.g1, V, xt);
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
        // objective terms spread over the block (summed by the same reduction)
        for (int t = tid; t < T.n_f; t += NT) { int aux; tv[2] += term_value(T.Ft + t, V, xt, &aux); }
        block_reduce<3>(tv, top, red);
        ft = ctl.fsc * tv[2];
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
      bool soft = false;
      if (!accepted && O.soft_resto) {
        // ---- soft restoration (IPOPT): the filter rejected every trial step; accept a step
        // along the same direction if it reduces the primal-dual error of the barrier problem
        __syncthreads();
        double pv[1]; const int pop[1] = {OP_SUM};
        pv[0] = 0.0;
        for (int i = tid; i < m; i += NT) {
          const int r = rt[i];
          if (r & 4) { pv[0] += fabs(g[i] - beq[i]); continue; }
          const double si = s[i];
          double zl = 0.0, zu = 0.0, acc = fabs(g[i] - si);
          if (r & 1) { zl = zL[i]; acc += fabs((si - sL[i]) * zl - mu); }
          if (r & 2) { zu = zU[i]; acc += fabs((sU[i] - si) * zu - mu); }
          pv[0] += acc + fabs(-y[i] - zl + zu);
        }
        for (int j = tid; j < n; j += NT) {
          double rx = gf[j];
          for (int q = T.colptr[j]; q < T.colptr[j + 1]; ++q) {
            const unsigned cr = __ldg(T.colrec + q);
            rx += jval[cr & 0xffffu] * y[cr >> 16];
          }
          pv[0] += fabs(rx);
        }
        block_reduce<1>(pv, pop, red);
        const double pd0 = pv[0];
        alpha = a_p;
        for (int n_try = 0; n_try < 12; ++n_try) {
          for (int j = tid; j < n; j += NT) xt[j] = xe[j] + alpha * dx[j];
          __syncthreads();
          if (XL) {
            for (int l = tid; l < T.n_mid; l += NT) { const int2 r = T.midg[l]; xt[n + 1 + l] = eval_range(T.Gt, r.x, r.y, V, xt); }
            __syncthreads();
            jac_xl(T, V, xt, jx, jsv, dsc, false);       // trial Jacobian -> jsv (free until the next sigma pass)
          }
          const double az = fmin(alpha, a_d);
          pv[0] = 0.0;
          for (int i = tid; i < m; i += NT) {
            const RowRec rr = T.rowrec[i];
            const int r = rt[i];
            const double d = dsc[i];
            if (!XL) {
              double acc = 0.0; int cur = 0, aux;
              double* jv = jsv + rr.s0;
              for (int k = rr.jt0; k < rr.jt1; ++k) {
                const double v = term_value(T.Jt + k, V, xt, &aux);
                if (aux != cur) { jv[cur] = d * acc; acc = 0.0; cur = aux; }
                acc += v;
              }
              if (rr.ns > 0) jv[cur] = d * acc;
            }
            const double gi = d * eval_range(T.Gt, rr.g0, rr.g1, V, xt);
            gt[i] = gi;
            const double yt = y[i] + alpha * dy[i];
            wv[i] = yt;
            if (r & 4) { pv[0] += fabs(gi - beq[i]); continue; }
            const double si = s[i] + alpha * ds[i];
            st[i] = si;
            double zl = 0.0, zu = 0.0, acc = fabs(gi - si);
            if (r & 1) { zl = zL[i] + az * dzL[i]; acc += fabs((si - sL[i]) * zl - mu); }
            if (r & 2) { zu = zU[i] + az * dzU[i]; acc += fabs((sU[i] - si) * zu - mu); }
            pv[0] += acc + fabs(-yt - zl + zu);
          }
          __syncthreads();
          for (int j = tid; j < n; j += NT) {
            double rx = ctl.fsc * eval_range(T.DFt, T.dfptr[j], T.dfptr[j + 1], V, xt);
            for (int q = T.colptr[j]; q < T.colptr[j + 1]; ++q) {
              const unsigned cr = __ldg(T.colrec + q);
              rx += jsv[cr & 0xffffu] * wv[cr >> 16];
            }
            pv[0] += fabs(rx);
          }
          block_reduce<1>(pv, pop, red);
          if (isfinite(pv[0]) && pv[0] <= SOFT_RESTO_FACTOR * pd0) {
            double fv[1]; fv[0] = 0.0;
            for (int t = tid; t < T.n_f; t += NT) { int aux; fv[0] += term_value(T.Ft + t, V, xt, &aux); }
            block_reduce<1>(fv, pop, red);
            ft = ctl.fsc * fv[0];
            accepted = true; soft = true; ftype = true;
            break;
          }
          alpha *= 0.5;
        }
      }
      if (!accepted) {
        if (ctl.n_restart < O.max_restarts) {
          // feasibility restart (stand-in for IPOPT's restoration phase, oracle/ipm_ref.py):
          // keep x, re-centre the slacks, zero the multipliers, clear the filter, mu = restart_mu
          __syncthreads();
          const double mu_r = O.restart_mu;
          for (int i = tid; i < m; i += NT) {
            const int r = rt[i];
            double si = g[i];
            if (r & 1) si = fmax(si, sL[i] + O.restart_push * fmax(1.0, fabs(sL[i])));
            if (r & 2) si = fmin(si, sU[i] - O.restart_push * fmax(1.0, fabs(sU[i])));
            s[i] = si; y[i] = 0.0;
            if (r & 1) zL[i] = mu_r / (si - sL[i]);
            if (r & 2) zU[i] = mu_r / (sU[i] - si);
          }
          if (tid == 0) {
            ctl.n_restart += 1; ctl.mu = mu_r; ctl.tau = fmax(TAU_MIN, 1.0 - mu_r);
            ctl.nfilt = 0; ctl.theta_max = -1.0; ctl.delta_w_last = 0.0;
          }
          __syncthreads();
          continue;
        }
        if (tid == 0) { ctl.status = OMG_RESTORATION_FAILED; ctl.iter = iter; }
        __syncthreads();
        break;
      }
      __syncthreads();
      if (tid == 0) {
        if (soft) ctl.nfilt = 0;
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
      for (int j = tid; j < n_xe; j += NT) xe[j] = xt[j];
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

__global__ void __launch_bounds__(512, 1)
omg_ipm_kernel(const DevTab T, const omg_options O, const Batch A, const Smem S) { ipm_body<false>(T, O, A, S); }

// sparse variant (omg_sp.cuh): L D L^T on the minimum-degree structure, thread streams, 128 or
// 256 threads per instance, as many blocks per SM as shared memory allows (config 2: 3)
#include "omg_sp.cuh"
__global__ void __launch_bounds__(128, 4)
omg_ipm_kernel_sp(const DevTab T, const SpTab P, const omg_options O, const Batch A, const SpSmem S) {
  ipm_body_sp(T, P, O, A, S);
}

__global__ void __launch_bounds__(256, 2)
omg_ipm_kernel_2cta(const DevTab T, const omg_options O, const Batch A, const Smem S) { ipm_body<false>(T, O, A, S); }

__global__ void __launch_bounds__(512, 1)
omg_ipm_kernel_xl(const DevTab T, const omg_options O, const Batch A, const Smem S) { ipm_body<true>(T, O, A, S); }

// XL with K in scratch: the block needs little shared memory, and the kernel is bound by the
// latency of its L2 streams, so two blocks per SM overlap better than one wide block
__global__ void __launch_bounds__(256, 2)
omg_ipm_kernel_xl_2cta(const DevTab T, const omg_options O, const Batch A, const Smem S) { ipm_body<true>(T, O, A, S); }

// warm-start shift: x[b, off + c*len + i] <- sum_k T[i,k] x[b, off + c*len + k]
__global__ void omg_shift_kernel(double* x, int B, int n, int n_blocks, const int* offs,
                                 const int* lens, const int* ncols, const int* toffs,
                                 const double* Tm) {
  OMG_DYN_SHARED(xs);
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

// vehicle models for the batched state prediction (reference Vehicle.ode of each class)
enum { OMG_ODE_INTEGRATOR = 0, OMG_ODE_QUADROTOR3D = 1, OMG_ODE_QUADROTOR2D = 2 };
#define OMG_ODE_MAX_STATE 8

__device__ __forceinline__ void ode_rhs(int model, int ns, const double* st, const double* u, double* d) {
  if (model == OMG_ODE_QUADROTOR3D) {          // quadrotor3d.py:308-312
    const double phi = st[6], theta = st[7], g = 9.81;
    d[0] = st[3]; d[1] = st[4]; d[2] = st[5];
    d[3] = u[0] * sin(theta) * cos(phi); d[4] = -u[0] * sin(phi);
    d[5] = -g + u[0] * cos(phi) * cos(theta); d[6] = u[1]; d[7] = u[2];
  } else if (model == OMG_ODE_QUADROTOR2D) {   // quadrotor.py:154-157
    const double theta = st[4], g = 9.81;
    d[0] = st[2]; d[1] = st[3]; d[2] = u[0] * sin(theta); d[3] = u[0] * cos(theta) - g; d[4] = u[1];
  } else {                                     // holonomic*.py: state' = input
    for (int j = 0; j < ns; ++j) d[j] = u[j];
  }
}

// Non-ideal prediction: integrate the vehicle ODE over `steps` samples of the planned input
// with classical RK4 (reference Vehicle.predict / integrate_ode, vehicle.py:302-337, 412-423;
// C++ twin Vehicle::integrate, export/vehicles/Vehicle.cpp:80-110: k1..k3 with input[i], k4
// with input[i+1]).  One thread per instance.
__global__ void omg_rk4_kernel(int model, int B, int ns, int ni, const double* __restrict__ state0,
                               const double* __restrict__ inputs, double dt, int steps,
                               double* __restrict__ stateT) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x[OMG_ODE_MAX_STATE], k1[OMG_ODE_MAX_STATE], k2[OMG_ODE_MAX_STATE], k3[OMG_ODE_MAX_STATE],
         k4[OMG_ODE_MAX_STATE], st[OMG_ODE_MAX_STATE];
  for (int j = 0; j < ns; ++j) x[j] = state0[(size_t)b * ns + j];
  const double* U = inputs + (size_t)b * (steps + 1) * ni;
  for (int i = 0; i < steps; ++i) {
    const double* u0 = U + (size_t)i * ni;
    const double* u1 = u0 + ni;
    ode_rhs(model, ns, x, u0, k1);
    for (int j = 0; j < ns; ++j) st[j] = x[j] + 0.5 * dt * k1[j];
    ode_rhs(model, ns, st, u0, k2);
    for (int j = 0; j < ns; ++j) st[j] = x[j] + 0.5 * dt * k2[j];
    ode_rhs(model, ns, st, u0, k3);
    for (int j = 0; j < ns; ++j) st[j] = x[j] + dt * k3[j];
    ode_rhs(model, ns, st, u1, k4);
    for (int j = 0; j < ns; ++j) x[j] += (dt / 6.0) * (k1[j] + 2.0 * k2[j] + 2.0 * k3[j] + k4[j]);
  }
  for (int j = 0; j < ns; ++j) stateT[(size_t)b * ns + j] = x[j];
}

// trajectory sampling: out[b, blk, c, s] = sum_k S_blk[s,k] * x[b, off_blk + c*len_blk + k]
// (batched Cox-de Boor evaluation with precomputed basis rows; reference
//  Vehicle.store -> sample_splines, vehicle.py:250-300, spline_extra.py:406-410;
//  C++ twin Vehicle::sampleSplines, Vehicle.cpp:112-190)
__global__ void omg_sample_kernel(const double* __restrict__ x, int B, int n, int n_blocks,
                                  const int* __restrict__ offs, const int* __restrict__ lens,
                                  const int* __restrict__ ncols, const int* __restrict__ nsamp,
                                  const int* __restrict__ soffs, const int* __restrict__ ooffs,
                                  const double* __restrict__ Sm, double* __restrict__ out, int n_out) {
  OMG_DYN_SHARED(xs);
  const int b = blockIdx.x;
  if (b >= B) return;
  const double* xb = x + (size_t)b * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = xb[i];
  __syncthreads();
  double* ob = out + (size_t)b * n_out;
  for (int blk = 0; blk < n_blocks; ++blk) {
    const int L = lens[blk], nc = ncols[blk], ns = nsamp[blk], off = offs[blk];
    const double* Sb = Sm + soffs[blk];
    for (int e = threadIdx.x; e < ns * nc; e += blockDim.x) {
      const int c = e / ns, sidx = e - c * ns;
      double acc = 0.0;
      for (int k = 0; k < L; ++k) acc += Sb[sidx * L + k] * xs[off + c * L + k];
      ob[ooffs[blk] + c * ns + sidx] = acc;
    }
  }
}

// ---------------------------------------------------------------------------
// ADMM consensus step of one agent per block (reference admm.py:117-168 z-update,
// 248-266 lambda-update, 268-307 residuals), in first-knot-shifted coordinates:
//   v   = Tf (x + l/rho)            for the own copy and every neighbour copy
//   z~  = P v + c                   P = I - A^T (A A^T)^-1 A,  c = A^T (A A^T)^-1 b
//   z   = Tb z~ ;  l += rho (x - z)
//   pr  = |Tf (x - z)|^2 ; dr = rho |Tf (z - z_prev)|^2 ; cr = rho pr + dr
// Tf / Tb (L x L, row-major) restrict a spline to the future / undo it.
// ---------------------------------------------------------------------------
__global__ void omg_admm_zl_kernel(int nsh, int nn, int L, const double* __restrict__ PzT,
                                   const double* __restrict__ c, const double* __restrict__ Tf,
                                   const double* __restrict__ Tb, double rho,
                                   const double* __restrict__ x_i, const double* __restrict__ x_j,
                                   double* __restrict__ z_i, double* __restrict__ z_ij,
                                   double* __restrict__ l_i, double* __restrict__ l_ij,
                                   double* __restrict__ res) {
  OMG_DYN_SHARED(sh);
  const int nz = nsh * (1 + nn);
  double* xs = sh;            // x  (own, neighbours)      [nz]
  double* ls = xs + nz;       // l                          [nz]
  double* zp = ls + nz;       // previous z                 [nz]
  double* v = zp + nz;        // Tf (x + l/rho)             [nz]
  double* zt = v + nz;        // z~ then z                  [nz]
  double* red = zt + nz;      // [2 * blockDim/32]
  const int a = blockIdx.x, tid = threadIdx.x;
  for (int k = tid; k < nz; k += blockDim.x) {
    const bool own = k < nsh;
    xs[k] = own ? x_i[(size_t)a * nsh + k] : x_j[(size_t)a * nn * nsh + (k - nsh)];
    ls[k] = own ? l_i[(size_t)a * nsh + k] : l_ij[(size_t)a * nn * nsh + (k - nsh)];
    zp[k] = own ? z_i[(size_t)a * nsh + k] : z_ij[(size_t)a * nn * nsh + (k - nsh)];
  }
  __syncthreads();
  for (int k = tid; k < nz; k += blockDim.x) {       // v = Tf (x + l/rho), spline by spline
    const int blk = k / L, r = k - blk * L;
    double acc = 0.0;
    for (int q = 0; q < L; ++q) acc += Tf[r * L + q] * (xs[blk * L + q] + ls[blk * L + q] / rho);
    v[k] = acc;
  }
  __syncthreads();
  for (int k = tid; k < nz; k += blockDim.x) {       // z~ = P v + c
    double acc = c[(size_t)a * nz + k];
    for (int q = 0; q < nz; ++q) acc += PzT[(size_t)q * nz + k] * v[q];
    zt[k] = acc;
  }
  __syncthreads();
  double znew = 0.0;
  for (int k = tid; k < nz; k += blockDim.x) {       // z = Tb z~ (one entry per thread, nz <= blockDim)
    const int blk = k / L, r = k - blk * L;
    double acc = 0.0;
    for (int q = 0; q < L; ++q) acc += Tb[r * L + q] * zt[blk * L + q];
    znew = acc;
  }
  __syncthreads();
  for (int k = tid; k < nz; k += blockDim.x) zt[k] = znew;
  __syncthreads();
  double pr = 0.0, dr = 0.0;
  for (int k = tid; k < nz; k += blockDim.x) {
    const bool own = k < nsh;
    const double zk = zt[k];
    const double lk = ls[k] + rho * (xs[k] - zk);
    if (own) { z_i[(size_t)a * nsh + k] = zk; l_i[(size_t)a * nsh + k] = lk; }
    else { z_ij[(size_t)a * nn * nsh + (k - nsh)] = zk; l_ij[(size_t)a * nn * nsh + (k - nsh)] = lk; }
    const int blk = k / L, r = k - blk * L;
    double e1 = 0.0, e2 = 0.0;
    for (int q = 0; q < L; ++q) {
      const double t = Tf[r * L + q];
      e1 += t * (xs[blk * L + q] - zt[blk * L + q]);
      e2 += t * (zt[blk * L + q] - zp[blk * L + q]);
    }
    pr += e1 * e1; dr += rho * e2 * e2;
  }
  for (int off = 16; off > 0; off >>= 1) { pr += __shfl_down_sync(FULL, pr, off); dr += __shfl_down_sync(FULL, dr, off); }
  const int nw = blockDim.x >> 5;
  if ((tid & 31) == 0) { red[tid >> 5] = pr; red[nw + (tid >> 5)] = dr; }
  __syncthreads();
  if (tid == 0) {
    double p = 0.0, d = 0.0;
    for (int w = 0; w < nw; ++w) { p += red[w]; d += red[nw + w]; }
    res[(size_t)a * 3 + 0] = p; res[(size_t)a * 3 + 1] = d; res[(size_t)a * 3 + 2] = rho * p + d;
  }
}

// ---------------------------------------------------------------------------
// Feasibility phase (fallback after Restoration_Failed; oracle/ipm_ref.py feasibility_lm):
// Levenberg-Marquardt on the constraint violation v(x) = g - clip(g, lbg, ubg),
//   (Jv^T Jv + lam I) dx = -Jv^T v,   Jv = the rows with v != 0,
// accept x + dx when it lowers 1/2 |v|^2 (then lam /= 10), else lam *= 10 (12 tries).
// One block per instance at a time; everything lives in the block's L2-resident scratch:
//   V[n_v] | jx[nnz_jx] | v[m] | vt[m] | A[n*n] | L[(n+1)*n] | rhs[n] | dx[n] | xe[n_xe] | xt[n_xe]
// A is assembled one thread per column through the CSC view of the Jacobian (no atomics:
// the sums are in a fixed order), the dense Cholesky carries the right-hand side as row n.
// ---------------------------------------------------------------------------
struct FeasArgs {
  int B, bounds_shared, max_steps;
  const double *x0, *p, *lbg, *ubg;
  double *x, *viol; int* steps;
  double* scr; size_t stride;
};

__device__ __forceinline__ double feas_viol(double g, double lb, double ub) {
  return (g < lb) ? g - lb : ((g > ub) ? g - ub : 0.0);
}

// v[i] for the point xq (mids of xq refreshed first); returns 1/2 |v|^2 and max |v| to all threads
__device__ __forceinline__ void feas_residual(const DevTab& T, const double* V, double* xq, const double* lbg,
                                              const double* ubg, double* v, double* red, double* phi, double* vmax) {
  const int tid = threadIdx.x;
  for (int l = tid; l < T.n_mid; l += NT) { const int2 r = T.midg[l]; xq[T.n + 1 + l] = eval_range(T.Gt, r.x, r.y, V, xq); }
  __syncthreads();
  double ss = 0.0, mx = 0.0;
  for (int i = tid; i < T.m; i += NT) {
    const RowRec rr = T.rowrec[i];
    const double vi = feas_viol(eval_range(T.Gt, rr.g0, rr.g1, V, xq), lbg[i], ubg[i]);
    v[i] = vi; ss += vi * vi; mx = fmax(mx, fabs(vi));
  }
  double r2[2] = {ss, mx}; const int o2[2] = {OP_SUM, OP_MAX};
  block_reduce<2>(r2, o2, red);
  *phi = 0.5 * r2[0]; *vmax = r2[1];
}

__global__ void __launch_bounds__(256, 2)
omg_feas_kernel(const DevTab T, const FeasArgs F) {
  __shared__ double red[MAX_NWARP * 2];
  __shared__ double piv_s;
  const int tid = threadIdx.x;
  const int n = T.n, m = T.m, n_xe = T.n + 1 + T.n_mid;
  double* D = F.scr + (size_t)blockIdx.x * F.stride;
  double* V = D;            double* jx = V + T.n_v;     double* v = jx + T.nnz_jx;
  double* vt = v + m;       double* Am = vt + m;        double* L = Am + (size_t)n * n;
  double* rhs = L + (size_t)(n + 1) * n;  double* dx = rhs + n;
  double* xe = dx + n;      double* xt = xe + n_xe;
  for (int inst = blockIdx.x; inst < F.B; inst += gridDim.x) {
    const double* par = F.p + (size_t)inst * T.n_par;
    const double* lbg = F.lbg + (F.bounds_shared ? 0 : (size_t)inst * m);
    const double* ubg = F.ubg + (F.bounds_shared ? 0 : (size_t)inst * m);
    for (int i = tid; i < 1 + T.n_par; i += NT) V[i] = (i == 0) ? 1.0 : par[i - 1];
    __syncthreads();
    for (int l = 0; l < T.n_levels; ++l) {      // parameter tape (as ipm_body S1)
      for (int e = T.level_ptr[l] + tid; e < T.level_ptr[l + 1]; e += NT) {
        double acc = 0.0;
        for (int t = T.tape_ptr[e]; t < T.tape_ptr[e + 1]; ++t) {
          const int4 f = __ldg(reinterpret_cast<const int4*>(T.tape_fac) + t);
          acc += T.tape_coef[t] * V[f.x] * V[f.y] * V[f.z] * V[f.w];
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
    for (int i = tid; i < n_xe; i += NT) {
      const double xi = (i < n) ? F.x0[(size_t)inst * n + i] : ((i == n) ? 1.0 : 0.0);
      xe[i] = xi; xt[i] = xi;
    }
    __syncthreads();
    double phi, vmax, lam = 1e-3;
    feas_residual(T, V, xe, lbg, ubg, v, red, &phi, &vmax);
    int steps = 0;
    while (steps < F.max_steps && vmax > 1e-8) {
      jac_xl(T, V, xe, jx, jx, nullptr, true);           // jx[0..nnz_j) = Jacobian slots (jval aliases jx)
      // normal equations, thread c owns row c of A: sum over the active rows of column c
      for (int c = tid; c < n; c += NT) {
        double* Ac = Am + (size_t)c * n;
        for (int k = 0; k < n; ++k) Ac[k] = 0.0;
        double bc = 0.0;
        for (int q = T.colptr[c]; q < T.colptr[c + 1]; ++q) {
          const unsigned cr = __ldg(T.colrec + q);
          const int s1 = (int)(cr & 0xffffu), r = (int)(cr >> 16);
          const double vr = v[r];
          if (vr == 0.0) continue;
          const double j1 = jx[s1];
          bc -= j1 * vr;
          const RowRec rr = T.rowrec[r];
          for (int k = 0; k < rr.ns; ++k) Ac[T.jcol16[rr.s0 + k]] += j1 * jx[rr.s0 + k];
        }
        rhs[c] = bc;
      }
      __syncthreads();
      bool accepted = false;
      for (int attempt = 0; attempt < 12 && !accepted; ++attempt) {
        // L = lower(A) + lam I, row n = rhs
        for (int e = tid; e < (n + 1) * n; e += NT) {
          const int i = e / n, k = e - i * n;
          L[e] = (i == n) ? rhs[k] : ((k <= i) ? Am[e] + ((k == i) ? lam : 0.0) : 0.0);
        }
        __syncthreads();
        bool ok = true;
        for (int j = 0; j < n; ++j) {
          if (tid == 0) {
            const double d = L[(size_t)j * n + j];
            const double pv = (d > 0.0 && d < 1e300) ? sqrt(d) : -1.0;
            piv_s = pv;
            if (pv > 0.0) L[(size_t)j * n + j] = pv;
          }
          __syncthreads();
          const double pv = piv_s;
          if (!(pv > 0.0)) { ok = false; break; }
          const double inv = 1.0 / pv;
          for (int i = j + 1 + tid; i <= n; i += NT) L[(size_t)i * n + j] *= inv;
          __syncthreads();
          for (int i = j + 1 + tid; i <= n; i += NT) {
            double* Li = L + (size_t)i * n;
            const double lij = Li[j];
            const int kend = (i < n) ? i : n - 1;
            for (int k = j + 1; k <= kend; ++k) Li[k] -= lij * L[(size_t)k * n + j];
          }
          __syncthreads();
        }
        __syncthreads();
        double pt = 0.0, vmt = 0.0;
        if (ok) {
          // back substitution L^T dx = w (w = row n)
          double* w = L + (size_t)n * n;
          for (int j = n - 1; j >= 0; --j) {
            if (tid == 0) dx[j] = w[j] / L[(size_t)j * n + j];
            __syncthreads();
            const double dj = dx[j];
            for (int k = tid; k < j; k += NT) w[k] -= L[(size_t)j * n + k] * dj;
            __syncthreads();
          }
          for (int i = tid; i < n; i += NT) xt[i] = xe[i] + dx[i];
          __syncthreads();
          feas_residual(T, V, xt, lbg, ubg, vt, red, &pt, &vmt);
        }
        if (ok && pt < phi) {          // (a NaN pt compares false)
          for (int i = tid; i < n_xe; i += NT) xe[i] = xt[i];
          for (int i = tid; i < m; i += NT) v[i] = vt[i];
          phi = pt; vmax = vmt; lam = fmax(lam / 10.0, 1e-12);
          accepted = true;
        } else {
          lam *= 10.0;
        }
        __syncthreads();
      }
      if (!accepted) break;
      ++steps;
    }
    for (int i = tid; i < n; i += NT) F.x[(size_t)inst * n + i] = xe[i];
    if (tid == 0) { F.viol[inst] = vmax; F.steps[inst] = steps; }
    __syncthreads();
  }
}


// ===========================================================================
// host side: C ABI
// ===========================================================================
static thread_local std::string g_err;
static void set_err(const std::string& s) { g_err = s; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  set_err(std::string(#call) + ": " + cudaGetErrorString(e_)); return -1; } } while (0)

// Device buffers for the block descriptors of the shift / sampling calls.  They only GROW (a
// receding-horizon loop passes the same block layout every step, the sampling matrix changes
// with the time): after the first call there is no cudaMalloc / cudaFree, and the small
// host-to-device copies are enqueued on the CALLER's stream every time, so calls on different
// streams are ordered correctly and stay asynchronous.
struct DescCache {
  int device = -1;
  size_t cap_i = 0, cap_d = 0;
  int* d_i = nullptr; double* d_d = nullptr;
  ~DescCache() { if (d_i) cudaFree(d_i); if (d_d) cudaFree(d_d); }
};

static void free_desc(DescCache* c) { delete c; }

struct omg_problem {
  int device = 0;
  DevTab T;
  Smem S;
  omg_options opt;
  std::vector<void*> allocs;
  int n_sm = 0, ctas_per_sm = 1, nt = 512, target_ctas = 1;
  bool xl = false;
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
  // scratch of the feasibility phase (omg_feas_batch), sized on first use
  double* fscr = nullptr; int fscr_ctas = 0; size_t fscr_stride = 0;
  DescCache* shift_desc = nullptr;   // device copy of the last omg_shift_batch block descriptor
  // sparse kernel variant (omg_sp.cuh / omg_sp_host.cuh)
  bool sp = false; SpTab P; SpSmem SS; size_t sp_smem_bytes = 0; int sp_ctas = 0, sp_dscr_stride = 0;
  std::string sp_info, sp_info_extra;
  // launch configuration of the envelope kernels (kept: inertia_mode = 1 is tied to the
  // envelope's elimination order and always runs there)
  int env_nt = 0, env_ctas = 0; size_t env_smem_bytes = 0;
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

#include "omg_sp_host.cuh"

// pack a term list into 32-byte records; aux = lrow (W) / slot offset within the row (J)
static const PTerm* upload_terms(omg_problem* h, const omg_termlist& L, int n_one,
                                 const std::vector<int>* aux, bool* ok) {
  std::vector<PTerm> pk((size_t)L.n_terms + 1);
  memset(pk.data(), 0, pk.size() * sizeof(PTerm));
  for (int k = 0; k < L.n_terms; ++k) {
    PTerm& q = pk[k];
    q.coef = L.coef[k]; q.cidx = (unsigned short)L.cidx[k];
    q.aux = (unsigned short)(aux ? (*aux)[k] : (L.lrow ? L.lrow[k] : 0));
    for (int w = 0; w < MAXW; ++w)
      q.x[w] = (unsigned short)(w < L.width ? L.xi[(size_t)k * L.width + w] : n_one);
  }
  for (int w = 0; w < MAXW; ++w) pk[L.n_terms].x[w] = (unsigned short)n_one;
  return upload(h, pk.data(), pk.size(), ok);
}

extern "C" {

int omg_abi_version(void) { return OMG_ABI_VERSION; }
#ifdef OMG_CPU_EMU
// only the CPU emulation build (tools/cpu_emu) exports this: its "device" pointers are host
// pointers, which lets the tests drive the device-pointer API with CPU tensors
int omg_is_emulation(void) { return 1; }
#endif

// ---- table files ---------------------------------------------------------------
namespace {
struct TabField { const char* name; int dtype; size_t off; int scalar; };
#define TF_S(f)      {#f, 0, offsetof(omg_tables, f), 1}
#define TF_I(f)      {#f, 0, offsetof(omg_tables, f), 0}
#define TF_D(f)      {#f, 1, offsetof(omg_tables, f), 0}
#define TF_LS(l, f)  {#l "." #f, 0, offsetof(omg_tables, l) + offsetof(omg_termlist, f), 1}
#define TF_LI(l, f)  {#l "." #f, 0, offsetof(omg_tables, l) + offsetof(omg_termlist, f), 0}
#define TF_LD(l, f)  {#l "." #f, 1, offsetof(omg_tables, l) + offsetof(omg_termlist, f), 0}
#define TF_LIST(l)   TF_LS(l, n_out), TF_LS(l, n_terms), TF_LS(l, width), TF_LI(l, ptr), TF_LD(l, coef), \
                     TF_LI(l, cidx), TF_LI(l, xi), TF_LI(l, lrow)
const TabField kTabFields[] = {
  TF_S(n), TF_S(m), TF_S(n_par), TF_S(n_v), TF_S(degree), TF_S(n_tape), TF_S(n_tape_terms), TF_S(n_levels),
  TF_I(tape_func), TF_I(tape_ptr), TF_D(tape_coef), TF_I(tape_fac), TF_I(level_ptr),
  TF_LIST(G), TF_LIST(F), TF_LIST(DF), TF_LIST(J), TF_LIST(W),
  TF_S(nnz_j), TF_I(jrow), TF_I(jcol), TF_I(jrow_ptr),
  TF_S(n_mid), TF_S(nnz_jx), TF_S(n_jp), TF_S(n_mu), TF_I(jp_ptr), TF_I(jp_a), TF_I(jp_c),
  TF_I(mu_ptr), TF_I(mu_row), TF_I(mu_slot),
  TF_S(nnz_w), TF_I(wrow), TF_I(wcol), TF_I(w2h),
  TF_S(nnz_h), TF_S(n_hp), TF_I(hrow), TF_I(hcol), TF_I(hp_ptr), TF_I(hp_s1), TF_I(hp_s2), TF_I(hp_row),
  TF_D(lbg), TF_D(ubg),
  TF_S(kkt_n), TF_S(kkt_n_eq), TF_S(env_size), TF_S(n_panel_rows), TF_S(max_panel_rows),
  TF_I(kkt_eq_rows), TF_I(kkt_pos_var), TF_I(kkt_pos_eq), TF_I(kkt_sign), TF_I(env_first), TF_I(env_ptr),
  TF_I(kkt_hdst), TF_I(kkt_jdst), TF_I(kkt_diag), TF_I(kkt_panel_ptr), TF_I(kkt_panel_rows),
  TF_S(nnz_wx), TF_S(n_xq), TF_S(n_xp), TF_I(xq_h), TF_I(xq_ptr), TF_I(xq_w), TF_I(xq_a), TF_I(xq_b),
};
struct OwnedTables { omg_tables T; std::vector<void*> blocks; };
}  // namespace

omg_tables* omg_tables_read(const char* path) {
  FILE* fp = path ? fopen(path, "rb") : nullptr;
  if (!fp) { set_err(std::string("cannot open table file ") + (path ? path : "(null)")); return nullptr; }
  OwnedTables* O = new OwnedTables();
  memset(&O->T, 0, sizeof(O->T));
  bool ok = true;
  char magic[8]; int32_t ver = 0, nrec = 0;
  if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, "OMGTBL\0\0", 8) != 0) { set_err("not an omg table file"); ok = false; }
  if (ok && (fread(&ver, 4, 1, fp) != 1 || fread(&nrec, 4, 1, fp) != 1)) { set_err("truncated table file"); ok = false; }
  if (ok && ver != OMG_ABI_VERSION) { set_err("table file written for another ABI version"); ok = false; }
  O->T.abi_version = ver;
  const int nf = (int)(sizeof(kTabFields) / sizeof(kTabFields[0]));
  std::vector<char> seen(nf, 0);
  for (int r = 0; ok && r < nrec; ++r) {
    char name[24]; int32_t dtype = 0, pad = 0; int64_t count = 0;
    if (fread(name, 1, 24, fp) != 24 || fread(&dtype, 4, 1, fp) != 1 || fread(&pad, 4, 1, fp) != 1 ||
        fread(&count, 8, 1, fp) != 1 || count < 0) { set_err("truncated table file"); ok = false; break; }
    name[23] = 0;
    int k = -1;
    for (int q = 0; q < nf; ++q) if (strcmp(kTabFields[q].name, name) == 0) { k = q; break; }
    const size_t esz = dtype ? 8 : 4;
    if (k < 0 || kTabFields[k].dtype != dtype) { set_err(std::string("unknown record in table file: ") + name); ok = false; break; }
    char* base = reinterpret_cast<char*>(&O->T) + kTabFields[k].off;
    if (kTabFields[k].scalar) {
      if (count != 1 || fread(base, 4, 1, fp) != 1) { set_err("bad scalar record"); ok = false; break; }
    } else {
      void* blk = malloc((size_t)(count > 0 ? count : 1) * esz);
      O->blocks.push_back(blk);
      if (!blk || (count > 0 && fread(blk, esz, (size_t)count, fp) != (size_t)count)) { set_err("truncated table file"); ok = false; break; }
      *reinterpret_cast<void**>(base) = (count > 0 || strcmp(name + strlen(name) - 4, "lrow") != 0) ? blk : nullptr;
    }
    seen[k] = 1;
  }
  fclose(fp);
  // lrow of the lists without multipliers is legitimately absent; everything else is required
  for (int q = 0; ok && q < nf; ++q)
    if (!seen[q] && !strstr(kTabFields[q].name, ".lrow")) { set_err(std::string("table file lacks ") + kTabFields[q].name); ok = false; }
  if (!ok) { omg_tables_free(&O->T); return nullptr; }
  return &O->T;
}

void omg_tables_free(omg_tables* tables) {
  if (!tables) return;
  OwnedTables* O = reinterpret_cast<OwnedTables*>(tables);   // T is the first member
  for (void* b : O->blocks) free(b);
  delete O;
}
const char* omg_last_error(void) { return g_err.c_str(); }

void omg_default_options(omg_options* o) {
  o->tol = 1e-3; o->constr_viol_tol = 1e-4; o->dual_inf_tol = 1.0; o->compl_inf_tol = 1e-4;
  o->mu_init = 0.1; o->bound_push = 1e-3; o->bound_frac = 1e-3; o->mult_bound_push = 1e-3;
  o->bound_relax_factor = 1e-8; o->scaling_max_gradient = 100.0;
  o->max_iter = 3000; o->trace = 0;
  o->max_restarts = 5; o->soft_resto = 1; o->restart_mu = 1.0; o->restart_push = 0.1;
  o->inertia_mode = 0; o->reserved = 0;
}

omg_problem* omg_problem_create(const omg_tables* tb, const omg_options* opt, int device) {
  if (!tb || tb->abi_version != OMG_ABI_VERSION) { set_err("omg_tables ABI version mismatch"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_err("no CUDA device available: libomgb200 has no CPU fallback"); return nullptr; }
  if (device < 0 || device >= ndev) { set_err("invalid device index"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); return nullptr; }
  g_err.clear();
  omg_problem* h = new omg_problem();
  h->device = device;
  if (opt) h->opt = *opt; else omg_default_options(&h->opt);
  bool ok = true;
  DevTab& T = h->T;
  memset(&T, 0, sizeof(T));
  const int n = tb->n, m = tb->m;
  T.n = n; T.m = m; T.n_par = tb->n_par; T.n_v = tb->n_v;
  T.n_tape = tb->n_tape; T.n_levels = tb->n_levels;
  T.nnz_j = tb->nnz_j; T.nnz_w = tb->nnz_w; T.nnz_h = tb->nnz_h;
  T.n_eq = tb->kkt_n_eq; T.N = tb->kkt_n;
  T.env_size = tb->env_size; T.max_panel_rows = tb->max_panel_rows;
  T.n_panel_rows = tb->n_panel_rows;
  T.n_panels = (tb->kkt_n + NB - 1) / NB;
  if (tb->kkt_n != n + tb->kkt_n_eq) { set_err("inconsistent KKT structure"); ok = false; }
  const int n_mid = tb->n_mid;
  T.n_mid = n_mid; T.nnz_jx = n_mid ? tb->nnz_jx : tb->nnz_j;
  if (!(n + 1 + n_mid < 65535 && tb->n_v < 65536 && m + 1 + n_mid < 65535 && tb->nnz_j < 65536)) {
    set_err("problem too large for 16-bit packed indices"); ok = false; }
  const omg_termlist* lists[5] = {&tb->G, &tb->F, &tb->DF, &tb->J, &tb->W};
  for (int k = 0; k < 5 && ok; ++k)
    if (lists[k]->width > MAXW) { set_err("term degree exceeds the record width (5 factors)"); ok = false; }
  for (int i = 0; ok && i <= tb->kkt_n; ++i)
    if (tb->env_first[i] % NB != 0) {
      set_err("envelope rows must start at multiples of the panel width"); ok = false; }
  if (!ok) { delete h; return nullptr; }

  T.tape_func = upload(h, tb->tape_func, tb->n_tape, &ok);
  T.tape_ptr = upload(h, tb->tape_ptr, (size_t)tb->n_tape + 1, &ok);
  T.tape_coef = upload(h, tb->tape_coef, tb->n_tape_terms, &ok);
  T.tape_fac = upload(h, tb->tape_fac, (size_t)tb->n_tape_terms * 4, &ok);
  T.level_ptr = upload(h, tb->level_ptr, (size_t)tb->n_levels + 1, &ok);
  // term records
  T.Gt = upload_terms(h, tb->G, n, nullptr, &ok);
  T.Ft = upload_terms(h, tb->F, n, nullptr, &ok); T.n_f = tb->F.n_terms;
  T.DFt = upload_terms(h, tb->DF, n, nullptr, &ok);
  T.dfptr = upload(h, tb->DF.ptr, (size_t)n + 1, &ok);
  T.Wt = upload_terms(h, tb->W, n, nullptr, &ok);
  {
    std::vector<int> aux(tb->J.n_terms > 0 ? tb->J.n_terms : 1);
    for (int s = 0; s < tb->nnz_j; ++s) {
      const int off = s - tb->jrow_ptr[tb->jrow[s]];
      for (int t = tb->J.ptr[s]; t < tb->J.ptr[s + 1]; ++t) aux[t] = off;
    }
    T.Jt = upload_terms(h, tb->J, n, &aux, &ok);
  }
  {  // per-row records
    std::vector<RowRec> rr(m > 0 ? m : 1);
    for (int i = 0; i < m; ++i) {
      RowRec& r = rr[i];
      r.g0 = tb->G.ptr[i]; r.g1 = tb->G.ptr[i + 1];
      r.s0 = tb->jrow_ptr[i]; r.ns = tb->jrow_ptr[i + 1] - tb->jrow_ptr[i];
      r.jt0 = tb->J.ptr[r.s0]; r.jt1 = tb->J.ptr[r.s0 + r.ns];
      r.pad0 = r.pad1 = 0;
      // every slot of the row must own at least one term (aux bookkeeping); with
      // intermediates the chain-rule pass initialises every slot instead
      for (int s = r.s0; s < r.s0 + r.ns && !n_mid; ++s)
        if (tb->J.ptr[s + 1] == tb->J.ptr[s]) { set_err("empty Jacobian slot"); ok = false; }
    }
    T.rowrec = upload(h, rr.data(), rr.size(), &ok);
  }
  if (n_mid) {  // intermediates: term ranges and chain-rule lists
    std::vector<int2> mg(n_mid);
    for (int l = 0; l < n_mid; ++l) mg[l] = make_int2(tb->G.ptr[m + l], tb->G.ptr[m + l + 1]);
    T.midg = upload(h, mg.data(), mg.size(), &ok);

    T.jp_ptr = upload(h, tb->jp_ptr, (size_t)tb->nnz_j + 1, &ok);
    T.jp_a = upload(h, tb->jp_a, tb->n_jp, &ok);
    T.jp_c = upload(h, tb->jp_c, tb->n_jp, &ok);
    T.mu_ptr = upload(h, tb->mu_ptr, (size_t)n_mid + 1, &ok);
    T.mu_row = upload(h, tb->mu_row, tb->n_mu, &ok);
    T.mu_slot = upload(h, tb->mu_slot, tb->n_mu, &ok);
    std::vector<int> xv;   // extra slots with at least one x factor (xi != n, the constant 1)
    for (int s = tb->nnz_j; s < tb->nnz_jx; ++s) {
      bool dep = false;
      for (int t = tb->J.ptr[s]; t < tb->J.ptr[s + 1] && !dep; ++t)
        for (int w = 0; w < tb->J.width; ++w) if (tb->J.xi[(size_t)t * tb->J.width + w] != n) { dep = true; break; }
      if (dep) xv.push_back(s);
    }
    T.n_jxvar = (int)xv.size();
    if (xv.empty()) xv.push_back(0);
    T.jxvar = upload(h, xv.data(), xv.size(), &ok);
  }
  T.jtptr = upload(h, tb->J.ptr, (size_t)T.nnz_jx + 1, &ok);
  T.jrow = upload(h, tb->jrow, tb->nnz_j, &ok);
  {  // CSC view of the Jacobian pattern: slot | row << 16
    std::vector<int> cptr(n + 1, 0);
    std::vector<unsigned> crec(tb->nnz_j > 0 ? tb->nnz_j : 1);
    std::vector<unsigned short> jc(tb->nnz_j > 0 ? tb->nnz_j : 1);
    for (int s = 0; s < tb->nnz_j; ++s) { cptr[tb->jcol[s] + 1]++; jc[s] = (unsigned short)tb->jcol[s]; }
    for (int j = 0; j < n; ++j) cptr[j + 1] += cptr[j];
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (int s = 0; s < tb->nnz_j; ++s)
      crec[fill[tb->jcol[s]]++] = (unsigned)s | ((unsigned)tb->jrow[s] << 16);
    T.colptr = upload(h, cptr.data(), cptr.size(), &ok);
    T.colrec = upload(h, crec.data(), crec.size(), &ok);
    T.jcol16 = upload(h, jc.data(), jc.size(), &ok);
  }
  {  // H positions sorted by descending pair count (balanced warps) + packed pairs
    std::vector<int> order(tb->nnz_h);
    for (int q = 0; q < tb->nnz_h; ++q) order[q] = q;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      return (tb->hp_ptr[a + 1] - tb->hp_ptr[a]) > (tb->hp_ptr[b + 1] - tb->hp_ptr[b]); });
    std::vector<HqRec> hq(tb->nnz_h > 0 ? tb->nnz_h : 1);
    std::vector<unsigned> pack((size_t)tb->n_hp + 1);
    int pos = 0;
    std::vector<char> has(n, 0);
    for (int k = 0; k < tb->nnz_h; ++k) {
      const int q = order[k];
      HqRec& r = hq[k];
      r.dst = tb->kkt_hdst[q]; r.p0 = pos; r.diag = (tb->hrow[q] == tb->hcol[q]) ? 1 : 0;
      if (r.diag) has[tb->hrow[q]] = 1;
      for (int e = tb->hp_ptr[q]; e < tb->hp_ptr[q + 1]; ++e)
        pack[pos++] = (unsigned)tb->hp_s1[e] | ((unsigned)tb->hp_s2[e] << 16);
      r.p1 = pos;
      if (r.p1 - r.p0 >= 64) T.n_hq_heavy = k + 1;
    }
    for (int j = 0; j < n; ++j) if (!has[j]) {
      set_err("H pattern lacks a diagonal entry (variable without constraint)"); ok = false; break; }
    T.hq = upload(h, hq.data(), hq.size(), &ok);
    T.hpack = upload(h, pack.data(), pack.size(), &ok);
  }
  const int nnz_wx = n_mid ? tb->nnz_wx : 0;
  T.nnz_wx = nnz_wx; T.n_xq = nnz_wx ? tb->n_xq : 0;
  if (tb->W.n_out != tb->nnz_w + nnz_wx) { set_err("W term list does not match nnz_w + nnz_wx"); ok = false; }
  if (ok) {  // Hessian slots: destination in the envelope (cross slots: index into Wx) + term range
    std::vector<WRec> wr(tb->nnz_w + nnz_wx > 0 ? tb->nnz_w + nnz_wx : 1);
    for (int q = 0; q < tb->nnz_w + nnz_wx; ++q) {
      wr[q].dst = (q < tb->nnz_w) ? tb->kkt_hdst[tb->w2h[q]] : q - tb->nnz_w;
      wr[q].t0 = tb->W.ptr[q]; wr[q].t1 = tb->W.ptr[q + 1]; wr[q].pad = 0;
    }
    T.wrec = upload(h, wr.data(), wr.size(), &ok);
  }
  if (ok && nnz_wx) {  // cross-Hessian gather records
    std::vector<HqRec> xq(T.n_xq > 0 ? T.n_xq : 1);
    std::vector<int4> xp(tb->n_xp > 0 ? tb->n_xp : 1);
    for (int e = 0; e < T.n_xq && ok; ++e) {
      const int q = tb->xq_h[e];
      if (q < 0 || q >= tb->nnz_h) { set_err("cross-Hessian position out of range"); ok = false; break; }
      xq[e].dst = tb->kkt_hdst[q]; xq[e].p0 = tb->xq_ptr[e]; xq[e].p1 = tb->xq_ptr[e + 1];
      xq[e].diag = (tb->hrow[q] == tb->hcol[q]) ? 1 : 0;
    }
    for (int r = 0; r < tb->n_xp && ok; ++r) {
      const int a = tb->xq_a[r], b = tb->xq_b[r];
      if (tb->xq_w[r] < 0 || tb->xq_w[r] >= nnz_wx || a < tb->nnz_j || a >= tb->nnz_jx ||
          (b >= 0 && (b < tb->nnz_j || b >= tb->nnz_jx))) {
        set_err("extra Hessian product out of range"); ok = false; break; }
      xp[r] = make_int4(tb->xq_w[r], a, b, 0);
    }
    T.xq = upload(h, xq.data(), xq.size(), &ok);
    T.xqp = upload(h, xp.data(), xp.size(), &ok);
  }
  T.eq_rows = upload(h, tb->kkt_eq_rows, tb->kkt_n_eq, &ok);
  T.pos_var = upload(h, tb->kkt_pos_var, n, &ok);
  T.pos_eq = upload(h, tb->kkt_pos_eq, tb->kkt_n_eq, &ok);
  T.ksign = upload(h, tb->kkt_sign, tb->kkt_n, &ok);
  T.env_first = upload(h, tb->env_first, (size_t)tb->kkt_n + 1, &ok);
  T.env_ptr = upload(h, tb->env_ptr, (size_t)tb->kkt_n + 2, &ok);
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

  // ---- shared-memory layout: mandatory part, then per-instance arrays by priority
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) ok = false;
  h->n_sm = prop.multiProcessorCount;
  Smem& S = h->S;
  const int N = T.N;
  int off = 0;
  auto take = [&](int cnt) { int o = off; off += (cnt + 1) & ~1; return o; };
  const int xe_len = n + 1 + n_mid;
  // mandatory shared-memory part; K and the parameter tape V optionally in scratch
  auto layout = [&](bool k_smem, bool v_smem) {
    off = 0;
    S.K = k_smem ? take(T.env_size + 2) : -1;
    S.LDP = (T.max_panel_rows + 2 + 3) & ~3;
    S.Pt = take(NB * S.LDP); S.PtS = take(NB * S.LDP); S.Ld = take(NB * NB);
    S.rbase = take(S.LDP);               // 2*LDP ints: row base offsets + row ids
    S.xe = take(xe_len); S.xt = take(xe_len); S.dx = take(N + 1); S.u = take(N + 1);
    S.gf = take(n);
    S.diag0 = take(N + 1); S.invd = take(N + 1); S.V = v_smem ? take(T.n_v) : -1;
    S.red = take(MAX_NWARP * NRED); S.filt = take(2 * MAXF);
    S.rt8 = take((m + 7) / 8);
    S.sgn = take(N); S.eptr = take((N + 2 + 1) / 2); S.efirst = take((N + 1 + 1) / 2);
    S.pptr = take((T.n_panels + 1 + 1) / 2); S.prow = take((tb->n_panel_rows + 1) / 2);
    S.pcmin = take((T.n_panels + 1) / 2);
    return (size_t)off * 8;
  };
  // standard kernels keep K and V in shared memory; the XL kernel (intermediates, or a
  // structure too large for that) keeps V in scratch, and K too if it leaves no room
  bool k_in_smem = true;
  {
    cudaFuncAttributes f0, fx;
    size_t b0 = 0, bx = 0;
    if (cudaFuncGetAttributes(&f0, (const void*)omg_ipm_kernel) == cudaSuccess)
      b0 = (size_t)prop.sharedMemPerBlockOptin - f0.sharedSizeBytes;
    if (cudaFuncGetAttributes(&fx, (const void*)omg_ipm_kernel_xl) == cudaSuccess)
      bx = (size_t)prop.sharedMemPerBlockOptin - fx.sharedSizeBytes;
    h->xl = (n_mid > 0) || layout(true, true) > b0;
    const char* kg = getenv("OMG_B200_XL_KGLOBAL");   // tuning knob: 1 = K in scratch even if it fits
    const bool kglob = kg && atoi(kg) == 1;
    if (h->xl) {
      // what is read at random stays on the SM in this order: K, then the parameter tape V (one
      // load per TERM of every stream); the m-vectors are streamed one thread per row and go last
      if (!kglob && layout(true, true) <= bx) { }
      else if (!kglob && layout(true, false) <= bx) { }
      else { k_in_smem = false; if (layout(false, true) > bx) layout(false, false); }
    } else layout(true, true);
  }
  // blocks per SM: 2 x 256 threads overlap one block's serial pivots with the other's
  // parallel phases; 1 x 512 keeps every per-instance array in shared memory.
  cudaFuncAttributes fa;
  if (ok && cudaFuncGetAttributes(&fa, h->xl ? (const void*)omg_ipm_kernel_xl : (const void*)omg_ipm_kernel) != cudaSuccess) { set_err("cudaFuncGetAttributes failed"); ok = false; }
  const size_t budget1 = ok ? (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes : 0;
  const size_t budget2 = ok ? ((size_t)prop.sharedMemPerMultiprocessor - 2 * 1024) / 2 - fa.sharedSizeBytes : 0;
  {
    const char* e = getenv("OMG_B200_CTAS");
    const int want = (e && atoi(e) == 1) ? 1 : 2;
    h->target_ctas = (want == 2 && (!h->xl || !k_in_smem) && (size_t)off * 8 <= budget2) ? 2 : 1;
    h->nt = (h->target_ctas == 1) ? 512 : 256;
  }
  const void* kfn = h->xl ? ((h->target_ctas == 1) ? (const void*)omg_ipm_kernel_xl : (const void*)omg_ipm_kernel_xl_2cta)
                          : (h->target_ctas == 1) ? (const void*)omg_ipm_kernel : (const void*)omg_ipm_kernel_2cta;
  const size_t budget = (h->target_ctas == 1) ? budget1 : budget2;
  if (ok && (size_t)off * 8 > budget) {
    char buf[256];
    snprintf(buf, sizeof buf, "KKT envelope does not fit shared memory: need %zu B, have %zu B (n=%d, n_eq=%d)",
             (size_t)off * 8, budget, n, T.n_eq);
    set_err(buf); ok = false;
  }
  int goff = 0;   // global scratch offset (doubles)
  S.Kg = S.Vg = S.jxg = S.mug = S.Kcg = S.wxg = 0;
  if (h->xl) {
    auto gtake = [&](int cnt) { int o = goff; goff += (cnt + 1) & ~1; return o; };
    S.Vg = gtake(T.n_v);
    S.jxg = gtake(T.nnz_jx - T.nnz_j + 1);
    S.mug = gtake(n_mid + 1);
    S.wxg = gtake(T.nnz_wx + 1);
    if (!k_in_smem) S.Kg = gtake(T.env_size + 2);
  }
  S.Kcg = goff; goff += (T.env_size + 2 + 1) & ~1;
  const int sizes[N_ARR] = {tb->nnz_j, m, tb->nnz_j, m, m, m, m, m, m, m, m, m, m, m, m, m, m, m, m};
  // staging buffer of the chunked J^T Sigma J gather (XL kernel with the Jacobian values in
  // scratch): the panel buffers by default; three quarters of the free shared memory (up to
  // 2 x 4096 doubles) when that is more -- fewer, longer chunks; the streamed m-vectors take the rest
  S.hst = S.Pt; S.hch = NB * S.LDP;
  if (h->xl && getenv("OMG_B200_HCHUNK") && atoi(getenv("OMG_B200_HCHUNK")) == 1 &&
      (size_t)(off + 2 * ((tb->nnz_j + 1) & ~1)) * 8 > budget) {
    const long long free_d = ((long long)budget - (long long)off * 8) / 8;
    int ch = (int)std::min<long long>(4096, (free_d * 3 / 4) / 2) & ~1;
    if (ch > S.hch) { S.hst = off; S.hch = ch; off += 2 * ch; }
  }
  for (int k = 0; k < N_ARR; ++k) {
    const int cnt = (sizes[k] + 1) & ~1;
    const bool optional_low = (k >= A_ZL);     // lower-bound / equality-only arrays: rarely touched
    if (!optional_low && (size_t)(off + cnt) * 8 <= budget) { S.arr[k] = off; off += cnt; }
    else { S.arr[k] = -(goff + 1); goff += cnt; }
  }
  S.total = off;
  T.hc_nchunk = 0; T.hc_vt = 0;
  // (experimental, off by default: OMG_B200_HCHUNK=1.  Measured on config 4: the gather phase -21 %
  //  and the solve -4 % at n = 238 with 16 chunks, nothing at n = 406 where only the panel buffers
  //  are free (69 chunks); it changes the summation order, which the 300-600-iteration cold starts
  //  of HolonomicOrient amplify to 3e-4 -- not worth that without a larger gain)
  const char* hce = getenv("OMG_B200_HCHUNK");
  if (ok && hce && atoi(hce) == 1 && h->xl && S.arr[A_JVAL] < 0 && S.arr[A_JSV] < 0 && tb->nnz_h > 0) {
    // chunked J^T Sigma J gather (see the kernel): chunks of whole rows that fit one panel buffer
    const int CH = std::min(S.hch, 65535), VT = 512;
    std::vector<int> cslot(1, 0), chunk_of_row(m, 0);
    bool fits = true;
    int cur0 = 0;
    for (int i = 0; i < m && fits; ++i) {
      const int a = tb->jrow_ptr[i], b = tb->jrow_ptr[i + 1];
      if (b - a > CH) fits = false;
      if (b - cur0 > CH) { cslot.push_back(a); cur0 = a; }
      chunk_of_row[i] = (int)cslot.size() - 1;
    }
    cslot.push_back(tb->nnz_j);
    if (fits) {
      const int nc = (int)cslot.size() - 1;
      // per chunk: the pairs of every H position that fall into it, positions dealt to the
      // virtual threads by pair count (LPT).  A position may change hands between chunks: the
      // chunks are separated by barriers.
      std::vector<std::vector<std::vector<uint2>>> per(nc);     // [chunk][position in chunk] -> records
      {
        std::vector<int> slot_in_chunk(tb->nnz_h, -1);
        for (int c = 0; c < nc; ++c) per[c].clear();
        std::vector<int> last_chunk(tb->nnz_h, -1);
        for (int q = 0; q < tb->nnz_h; ++q)
          for (int e = tb->hp_ptr[q]; e < tb->hp_ptr[q + 1]; ++e) {
            const int c = chunk_of_row[tb->hp_row[e]];
            if (last_chunk[q] != c) { last_chunk[q] = c; slot_in_chunk[q] = (int)per[c].size(); per[c].push_back({}); }
            per[c][slot_in_chunk[q]].push_back(make_uint2((unsigned)tb->kkt_hdst[q],
                (unsigned)(tb->hp_s1[e] - cslot[c]) | ((unsigned)(tb->hp_s2[e] - cslot[c]) << 16)));
          }
      }
      std::vector<std::vector<uint2>> lists((size_t)nc * VT);
      typedef std::pair<long long, int> LT;
      for (int c = 0; c < nc; ++c) {
        std::vector<int> ord(per[c].size());
        for (size_t k = 0; k < ord.size(); ++k) ord[k] = (int)k;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return per[c][a].size() > per[c][b].size(); });
        std::priority_queue<LT, std::vector<LT>, std::greater<LT>> heap;
        for (int t = 0; t < VT; ++t) heap.push(LT(0, t));
        for (int k : ord) {
          LT top = heap.top(); heap.pop();
          std::vector<uint2>& dst = lists[(size_t)c * VT + top.second];
          dst.insert(dst.end(), per[c][k].begin(), per[c][k].end());
          heap.push(LT(top.first + (long long)per[c][k].size(), top.second));
        }
      }
      std::vector<int> base(nc), cnt((size_t)nc * VT);
      size_t total = 0;
      for (int c = 0; c < nc; ++c) {
        size_t mx = 0;
        for (int t = 0; t < VT; ++t) { cnt[(size_t)c * VT + t] = (int)lists[(size_t)c * VT + t].size(); mx = std::max(mx, lists[(size_t)c * VT + t].size()); }
        base[c] = (int)total; total += mx * VT;
      }
      std::vector<uint2> rec(std::max<size_t>(total, 1), make_uint2(0u, 0u));
      for (int c = 0; c < nc; ++c)
        for (int t = 0; t < VT; ++t) {
          const std::vector<uint2>& l = lists[(size_t)c * VT + t];
          for (size_t k = 0; k < l.size(); ++k) rec[(size_t)base[c] + k * VT + t] = l[k];
        }
      if (total < (size_t)1 << 31) {
        T.hc_slot = upload(h, cslot.data(), cslot.size(), &ok);
        T.hc_base = upload(h, base.data(), base.size(), &ok);
        T.hc_cnt = upload(h, cnt.data(), cnt.size(), &ok);
        T.hc_rec = upload(h, rec.data(), rec.size(), &ok);
        T.hc_nchunk = nc; T.hc_vt = VT;
        if (getenv("OMG_B200_VERBOSE"))
          fprintf(stderr, "[omg_b200] chunked H gather: %d chunks of <= %d slots, %d pairs, %zu record slots\n", nc, CH, tb->n_hp, total);
      }
    }
  }
  h->smem_bytes = (size_t)off * sizeof(double);
  if (ok && cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes)) != cudaSuccess) {
    set_err(std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(cudaGetLastError())); ok = false; }
  if (ok) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, h->nt, h->smem_bytes);
    h->ctas_per_sm = occ > 0 ? occ : 1;
    h->dscr_stride = goff + 8;
    h->iscr_stride = 2 * m + 8;
    {  // sparse variant: preferred whenever the problem is inside its coverage
      const char* e = getenv("OMG_B200_KERNEL");   // "envelope": force the envelope kernels
      std::string why;
      h->env_nt = h->nt; h->env_ctas = h->ctas_per_sm; h->env_smem_bytes = h->smem_bytes;
      if (!(e && strcmp(e, "envelope") == 0) && sp_setup(h, tb, prop, &why)) {
        h->sp = true;
        h->nt = h->P.nt; h->ctas_per_sm = h->sp_ctas;
        h->smem_bytes = h->sp_smem_bytes;
        if (h->sp_dscr_stride > h->dscr_stride) h->dscr_stride = h->sp_dscr_stride;
      } else h->sp_info = "envelope kernels (" + (why.empty() ? std::string("forced") : why) + ")";
      if (getenv("OMG_B200_VERBOSE")) {
        fprintf(stderr, "[omg_b200] %s\n", h->sp_info.c_str());
        if (!h->sp) {   // what the sparse ordering would give on this structure (diagnostic only)
          SpSym Y; std::string w2;
          const bool sy = sp_symbolic(tb, Y, &w2);
          int pairs = 0;
          for (int j = 0; j < Y.R0 && j < (int)Y.st.size(); ++j) pairs += (int)(Y.st[j].size() * (Y.st[j].size() + 1) / 2);
          fprintf(stderr, "[omg_b200]   symbolic (%s): N=%d Lsize=%d levels=%d root=%d pairs(non-root)=%d env_size=%d\n",
                  sy ? "ok" : w2.c_str(), Y.N, Y.Lsize, Y.n_lev, Y.nr, pairs, T.env_size);
        }
      }
    }
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
  if (h->fscr) cudaFree(h->fscr);
  free_desc(h->shift_desc);
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

const char* omg_structure_info(omg_problem* h) { return h ? h->sp_info.c_str() : ""; }

int omg_solve_batch(omg_problem* h, int32_t B, const double* x0, const double* p,
                    const double* lbg, const double* ubg, int32_t bounds_shared,
                    const double* lam_g0, double* x, double* lam_g, double* f,
                    int32_t* status, int32_t* iters, void* stream_) {
  if (!h) { set_err("null handle"); return -1; }
  if (B <= 0) return 0;
  if (!x0 || !p || !lbg || !ubg || !x || !lam_g || !f || !status || !iters) { set_err("null buffer"); return -1; }
  cudaStream_t stream = (cudaStream_t)stream_;
  CK(cudaSetDevice(h->device));
  const bool use_sp = h->sp && h->opt.inertia_mode == 0;
  int grid = h->n_sm * (use_sp ? h->sp_ctas : h->env_ctas);
  if (grid > B) grid = B;
  if (grid > h->scr_ctas) {
    if (h->dscr) cudaFree(h->dscr);
    if (h->iscr) cudaFree(h->iscr);
    h->dscr = nullptr; h->iscr = nullptr;
    const int want = h->n_sm * std::max(h->ctas_per_sm, h->env_ctas);
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
  if (use_sp) OMG_LAUNCH(omg_ipm_kernel_sp, grid, h->P.nt, h->sp_smem_bytes, stream, h->T, h->P, h->opt, A, h->SS);
  else if (h->xl && h->target_ctas == 1) OMG_LAUNCH(omg_ipm_kernel_xl, grid, 512, h->env_smem_bytes, stream, h->T, h->opt, A, h->S);
  else if (h->xl) OMG_LAUNCH(omg_ipm_kernel_xl_2cta, grid, 256, h->env_smem_bytes, stream, h->T, h->opt, A, h->S);
  else if (h->target_ctas == 1) OMG_LAUNCH(omg_ipm_kernel, grid, 512, h->env_smem_bytes, stream, h->T, h->opt, A, h->S);
  else OMG_LAUNCH(omg_ipm_kernel_2cta, grid, 256, h->env_smem_bytes, stream, h->T, h->opt, A, h->S);
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

static int ensure_feas_scratch(omg_problem* h, int grid) {
  const DevTab& T = h->T;
  const size_t n = T.n, m = T.m, n_xe = n + 1 + T.n_mid;
  const size_t stride = ((size_t)T.n_v + T.nnz_jx + 2 * m + n * n + (n + 1) * n + 2 * n + 2 * n_xe + 7) & ~(size_t)7;
  if (grid > h->fscr_ctas) {
    if (h->fscr) cudaFree(h->fscr);
    h->fscr = nullptr; h->fscr_ctas = 0;
    CK(cudaMalloc(&h->fscr, (size_t)grid * stride * sizeof(double)));
    h->fscr_ctas = grid;
  }
  h->fscr_stride = stride;
  return 0;
}

int omg_feas_batch(omg_problem* h, int32_t B, const double* x0, const double* p, const double* lbg,
                   const double* ubg, int32_t bounds_shared, int32_t max_steps, double* x, double* viol,
                   int32_t* steps, void* stream_) {
  if (!h) { set_err("null handle"); return -1; }
  if (B <= 0) return 0;
  if (!x0 || !p || !lbg || !ubg || !x || !viol || !steps) { set_err("null buffer"); return -1; }
  cudaStream_t stream = (cudaStream_t)stream_;
  CK(cudaSetDevice(h->device));
  int grid = h->n_sm * 2;
  if (grid > B) grid = B;
  if (ensure_feas_scratch(h, grid)) return -1;
  FeasArgs F;
  F.B = B; F.bounds_shared = bounds_shared; F.max_steps = max_steps;
  F.x0 = x0; F.p = p; F.lbg = lbg; F.ubg = ubg; F.x = x; F.viol = viol; F.steps = steps;
  F.scr = h->fscr; F.stride = h->fscr_stride;
  OMG_LAUNCH(omg_feas_kernel, grid, 256, 0, stream, h->T, F);
  CK(cudaGetLastError());
  return 0;
}

int omg_feas_batch_host(omg_problem* h, int32_t B, const double* x0, const double* p, const double* lbg,
                        const double* ubg, int32_t bounds_shared, int32_t max_steps, double* x, double* viol,
                        int32_t* steps) {
  if (!h) { set_err("null handle"); return -1; }
  if (B <= 0) return 0;
  CK(cudaSetDevice(h->device));
  if (ensure_staging(h, B, bounds_shared ? 1 : 0)) return -1;
  const size_t n = h->T.n, m = h->T.m, np_ = h->T.n_par, b = B;
  CK(cudaMemcpyAsync(h->hx0, x0, b * n * 8, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(h->hp, p, b * np_ * 8, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(h->hlb, lbg, (bounds_shared ? 1 : b) * m * 8, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(h->hub, ubg, (bounds_shared ? 1 : b) * m * 8, cudaMemcpyHostToDevice, 0));
  if (omg_feas_batch(h, B, h->hx0, h->hp, h->hlb, h->hub, bounds_shared, max_steps, h->hx, h->hf, h->hit, nullptr)) return -1;
  CK(cudaMemcpyAsync(x, h->hx, b * n * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(viol, h->hf, b * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(steps, h->hit, b * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return 0;
}

static int desc_upload(DescCache& c, int device, const std::vector<int>& iv, const double* dv, size_t nd,
                       cudaStream_t stream) {
  if (c.device != device || iv.size() > c.cap_i || nd > c.cap_d) {   // (re)allocate: first call / growth only
    if (c.d_i) cudaFree(c.d_i);
    if (c.d_d) cudaFree(c.d_d);
    c.d_i = nullptr; c.d_d = nullptr; c.device = -1;
    c.cap_i = std::max(iv.size(), (size_t)64) * 2; c.cap_d = std::max(nd, (size_t)64) * 2;
    CK(cudaMalloc(&c.d_i, sizeof(int) * c.cap_i));
    CK(cudaMalloc(&c.d_d, sizeof(double) * c.cap_d));
    c.device = device;
  }
  // pageable sources: the runtime stages them before returning, the caller's arrays are free again
  CK(cudaMemcpyAsync(c.d_i, iv.data(), sizeof(int) * iv.size(), cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(c.d_d, dv, sizeof(double) * nd, cudaMemcpyHostToDevice, stream));
  return 0;
}

int omg_shift_batch(omg_problem* h, int32_t B, double* x, int32_t n_blocks, const int32_t* offs,
                    const int32_t* lens, const int32_t* ncols, const double* Tm, void* stream_) {
  if (!h || !x || !offs || !lens || !ncols || !Tm) { set_err("null argument"); return -1; }
  if (B <= 0 || n_blocks <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  CK(cudaSetDevice(h->device));
  std::vector<int> iv(4 * (size_t)n_blocks);
  int tot = 0;
  for (int b = 0; b < n_blocks; ++b) {
    iv[b] = offs[b]; iv[n_blocks + b] = lens[b]; iv[2 * n_blocks + b] = ncols[b]; iv[3 * n_blocks + b] = tot;
    tot += lens[b] * lens[b];
  }
  if (!h->shift_desc) h->shift_desc = new DescCache();
  DescCache& c = *h->shift_desc;
  if (desc_upload(c, h->device, iv, Tm, (size_t)tot, stream)) return -1;
  OMG_LAUNCH(omg_shift_kernel, B, 128, sizeof(double) * h->T.n, stream, x, B, h->T.n, n_blocks, c.d_i, c.d_i + n_blocks,
                                                               c.d_i + 2 * n_blocks, c.d_i + 3 * n_blocks, c.d_d);
  CK(cudaGetLastError());
  return 0;
}

int omg_sample_batch(int32_t B, int32_t n, const double* x, int32_t n_blocks, const int32_t* offs,
                     const int32_t* lens, const int32_t* ncols, const int32_t* nsamp,
                     const double* Sm, double* out, void* stream_) {
  if (!x || !offs || !lens || !ncols || !nsamp || !Sm || !out) { set_err("null argument"); return -1; }
  if (B <= 0 || n_blocks <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  std::vector<int> iv(6 * (size_t)n_blocks);
  int stot = 0, otot = 0;
  for (int b = 0; b < n_blocks; ++b) {
    iv[b] = offs[b]; iv[n_blocks + b] = lens[b]; iv[2 * n_blocks + b] = ncols[b]; iv[3 * n_blocks + b] = nsamp[b];
    iv[4 * n_blocks + b] = stot; stot += nsamp[b] * lens[b];
    iv[5 * n_blocks + b] = otot; otot += nsamp[b] * ncols[b];
  }
  int device = 0;
  CK(cudaGetDevice(&device));
  static thread_local DescCache cache;         // (no handle in this call: one descriptor per host thread)
  if (desc_upload(cache, device, iv, Sm, (size_t)stot, stream)) return -1;
  const int* d_i = cache.d_i;
  OMG_LAUNCH(omg_sample_kernel, B, 128, sizeof(double) * n, stream, x, B, n, n_blocks, d_i, d_i + n_blocks, d_i + 2 * n_blocks,
                                                           d_i + 3 * n_blocks, d_i + 4 * n_blocks, d_i + 5 * n_blocks,
                                                           cache.d_d, out, otot);
  CK(cudaGetLastError());
  return 0;
}

int omg_integrate_rk4(int32_t model, int32_t B, int32_t n_state, int32_t n_input, const double* state0,
                      const double* inputs, double sample_time, int32_t steps, double* stateT, void* stream_) {
  if (B <= 0) return 0;
  if (!state0 || !inputs || !stateT) { set_err("null buffer"); return -1; }
  const int want_s[3] = {n_input, 8, 5}, want_i[3] = {n_input, 3, 2};
  if (model < 0 || model > 2 || n_state < 1 || n_state > OMG_ODE_MAX_STATE || steps < 0 ||
      n_state != want_s[model] || n_input != want_i[model]) { set_err("bad vehicle model / sizes"); return -1; }
  cudaStream_t stream = (cudaStream_t)stream_;
  OMG_LAUNCH(omg_rk4_kernel, (B + 127) / 128, 128, 0, stream, model, B, n_state, n_input, state0, inputs, sample_time, steps, stateT);
  CK(cudaGetLastError());
  return 0;
}

int omg_admm_zl_update(int32_t n_agents, int32_t nsh, int32_t n_nghb, int32_t L,
                       const double* PzT, const double* c, const double* Tf, const double* Tb,
                       double rho, const double* x_i, const double* x_j, double* z_i, double* z_ij,
                       double* l_i, double* l_ij, double* res, void* stream_) {
  if (n_agents <= 0) return 0;
  if (!PzT || !c || !Tf || !Tb || !x_i || !x_j || !z_i || !z_ij || !l_i || !l_ij || !res) { set_err("null buffer"); return -1; }
  const int nz = nsh * (1 + n_nghb);
  int nt = 32; while (nt < nz) nt <<= 1;
  if (nt > 1024 || nsh % L != 0) { set_err("unsupported consensus block size"); return -1; }
  const size_t smem = sizeof(double) * (5 * (size_t)nz + 64);
  OMG_LAUNCH(omg_admm_zl_kernel, n_agents, nt, smem, (cudaStream_t)stream_, nsh, n_nghb, L, PzT, c, Tf, Tb, rho,
                                                                   x_i, x_j, z_i, z_ij, l_i, l_ij, res);
  CK(cudaGetLastError());
  return 0;
}

// ---- multi-GPU ADMM exchange (NCCL bound at run time) -----------------------------------------
struct omg_comm {
  int n_ranks = 1, rank = 0, device = 0;
  void* nccl = nullptr;                 // ncclComm_t
  double* gx = nullptr; double* gz = nullptr; double* gl = nullptr; size_t cap_x = 0, cap_z = 0;
};

}  // extern "C" (reopened below)

#ifndef OMG_CPU_EMU
#include <dlfcn.h>
#endif
struct OmgNcclId { char internal[128]; };      // layout of ncclUniqueId
namespace {
struct NcclApi {
  bool ok = false;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, OmgNcclId /* ncclUniqueId by value */, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
const int kNcclFloat64 = 8, kNcclSum = 0;

bool nccl_bind() {
#ifdef OMG_CPU_EMU
  return false;
#else
  if (g_nccl.ok) return true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the host process' copy, if any
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { set_err("NCCL not found (dlopen libnccl.so.2)"); return false; }
  *(void**)(&g_nccl.GetUniqueId) = dlsym(lib, "ncclGetUniqueId");
  *(void**)(&g_nccl.CommInitRank) = dlsym(lib, "ncclCommInitRank");
  *(void**)(&g_nccl.CommDestroy) = dlsym(lib, "ncclCommDestroy");
  *(void**)(&g_nccl.AllGather) = dlsym(lib, "ncclAllGather");
  *(void**)(&g_nccl.AllReduce) = dlsym(lib, "ncclAllReduce");
  *(void**)(&g_nccl.GetErrorString) = dlsym(lib, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather &&
              g_nccl.AllReduce && g_nccl.GetErrorString;
  if (!g_nccl.ok) set_err("libnccl lacks a required symbol");
  return g_nccl.ok;
#endif
}
#define NK(call) do { int r_ = (call); if (r_ != 0) { set_err(std::string(#call) + ": " + g_nccl.GetErrorString(r_)); return -1; } } while (0)

// dst[i][k][:] = src[idx_a[i][k]] [ idx_b ? idx_b[i][k] : - ] [:]  (rows of `len` doubles)
__global__ void omg_gather_rows_kernel(int n_rows, int len, int stride_a, const int* __restrict__ ia,
                                       const int* __restrict__ ib, const double* __restrict__ src,
                                       double* __restrict__ dst) {
  const int r = blockIdx.x;
  if (r >= n_rows) return;
  const size_t base = (size_t)ia[r] * stride_a + (ib ? (size_t)ib[r] * len : 0);
  for (int q = threadIdx.x; q < len; q += blockDim.x) dst[(size_t)r * len + q] = src[base + q];
}
__global__ void omg_colsum3_kernel(int n, const double* __restrict__ res, double* __restrict__ out) {
  double a = 0.0;
  for (int i = 0; i < n; ++i) a += res[(size_t)i * 3 + threadIdx.x];   // fixed order: deterministic
  out[threadIdx.x] = a;
}
}  // namespace

extern "C" {

int omg_comm_unique_id(void* id128) {
  if (!id128) { set_err("null buffer"); return -1; }
  if (!nccl_bind()) return -1;
  NK(g_nccl.GetUniqueId(id128));
  return 0;
}

omg_comm* omg_comm_create(const void* id128, int32_t n_ranks, int32_t rank, int32_t device) {
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_err("bad rank / n_ranks"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice failed"); return nullptr; }
  omg_comm* c = new omg_comm();
  c->n_ranks = n_ranks; c->rank = rank; c->device = device;
  if (n_ranks > 1) {
    if (!id128 || !nccl_bind()) { if (!id128) set_err("null unique id"); delete c; return nullptr; }
    OmgNcclId id; memcpy(&id, id128, sizeof id);
    const int r = g_nccl.CommInitRank(&c->nccl, n_ranks, id, rank);
    if (r != 0) { set_err(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); delete c; return nullptr; }
  }
  return c;
}

void omg_comm_destroy(omg_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->nccl && g_nccl.ok) g_nccl.CommDestroy(c->nccl);
  if (c->gx) cudaFree(c->gx);
  if (c->gz) cudaFree(c->gz);
  if (c->gl) cudaFree(c->gl);
  delete c;
}

static int comm_reserve(omg_comm* c, size_t nx, size_t nz) {
  if (nx > c->cap_x) { if (c->gx) cudaFree(c->gx); c->gx = nullptr; CK(cudaMalloc(&c->gx, nx * 8)); c->cap_x = nx; }
  if (nz > c->cap_z) {
    if (c->gz) cudaFree(c->gz); if (c->gl) cudaFree(c->gl); c->gz = c->gl = nullptr;
    CK(cudaMalloc(&c->gz, nz * 8)); CK(cudaMalloc(&c->gl, nz * 8)); c->cap_z = nz;
  }
  return 0;
}

int omg_admm_exchange_x(omg_comm* c, int32_t n_local, int32_t nsh, int32_t n_nghb, const int32_t* nghb,
                        const double* x_i, double* x_j, void* stream_) {
  if (!c || !nghb || !x_i || !x_j) { set_err("null argument"); return -1; }
  if (n_local <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  CK(cudaSetDevice(c->device));
  const double* all = x_i;
  if (c->n_ranks > 1) {
    if (comm_reserve(c, (size_t)c->n_ranks * n_local * nsh, 0)) return -1;
    NK(g_nccl.AllGather(x_i, c->gx, (size_t)n_local * nsh, kNcclFloat64, c->nccl, stream));
    all = c->gx;
  }
  OMG_LAUNCH(omg_gather_rows_kernel, n_local * n_nghb, 32, 0, stream, n_local * n_nghb, nsh, nsh, nghb,
             (const int*)nullptr, all, x_j);
  CK(cudaGetLastError());
  return 0;
}

int omg_admm_zl_update_dist(omg_comm* c, int32_t n_local, int32_t nsh, int32_t n_nghb, int32_t L,
                            const double* PzT, const double* cvec, const double* Tf, const double* Tb,
                            double rho, const double* x_i, const double* x_j, double* z_i, double* z_ij,
                            double* l_i, double* l_ij, double* res, const int32_t* nghb, const int32_t* back,
                            double* z_ji, double* l_ji, double* res_total, void* stream_) {
  if (!c || !nghb || !back || !z_ji || !l_ji || !res_total) { set_err("null argument"); return -1; }
  if (n_local <= 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  CK(cudaSetDevice(c->device));
  if (omg_admm_zl_update(n_local, nsh, n_nghb, L, PzT, cvec, Tf, Tb, rho, x_i, x_j, z_i, z_ij, l_i, l_ij,
                         res, stream_)) return -1;
  OMG_LAUNCH(omg_colsum3_kernel, 1, 3, 0, stream, n_local, res, res_total);
  CK(cudaGetLastError());
  const double* allz = z_ij; const double* alll = l_ij;
  if (c->n_ranks > 1) {
    const size_t cnt = (size_t)n_local * n_nghb * nsh;
    if (comm_reserve(c, 0, (size_t)c->n_ranks * cnt)) return -1;
    NK(g_nccl.AllReduce(res_total, res_total, 3, kNcclFloat64, kNcclSum, c->nccl, stream));
    NK(g_nccl.AllGather(z_ij, c->gz, cnt, kNcclFloat64, c->nccl, stream));
    NK(g_nccl.AllGather(l_ij, c->gl, cnt, kNcclFloat64, c->nccl, stream));
    allz = c->gz; alll = c->gl;
  }
  OMG_LAUNCH(omg_gather_rows_kernel, n_local * n_nghb, 32, 0, stream, n_local * n_nghb, nsh, n_nghb * nsh, nghb, back, allz, z_ji);
  OMG_LAUNCH(omg_gather_rows_kernel, n_local * n_nghb, 32, 0, stream, n_local * n_nghb, nsh, n_nghb * nsh, nghb, back, alll, l_ji);
  CK(cudaGetLastError());
  return 0;
}

}  // extern "C"
