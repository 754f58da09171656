/* omg_b200.h -- C ABI of libomgb200.so: batched spline-NLP interior-point solve
 * on NVIDIA B200 (sm_100a).
 *
 * This library replaces, for OMG-tools' per-MPC-step solve, the CasADi+IPOPT
 * call of the reference:
 *
 *   omg_problem_create   <->  nlpsol('solver','ipopt',{x,p,f,g},opts)
 *                              omgtools/basics/optilayer.py:49-60 (create_nlp);
 *                              C++ twin: Point2Point.cpp:80-91 (nlpsol(..."nlp.so"))
 *   omg_solve_batch      <->  result = self.problem(x0=var,p=par,lbg=lb,ubg=ub)
 *                              omgtools/problems/problem.py:113 (and admm.py:390);
 *                              C++ twin: Point2Point::solve, Point2Point.cpp:207-231
 *   status/iters arrays  <->  self.problem.stats()['return_status']
 *                              omgtools/problems/problem.py:119-128
 *   omg_feas_batch       <->  (IPOPT's restoration phase inside the same nlpsol call)
 *   omg_shift_batch      <->  father.transform_primal_splines(T.dot(coeffs))
 *                              omgtools/problems/point2point.py:187-198,
 *                              optilayer.py:470-490; C++: transformSplines,
 *                              export.py:406-444
 *   omg_problem_destroy  <->  (garbage collection of the casadi Function)
 *
 * Conventions: plain C, no C++/torch types.  Unless a function name ends in
 * _host, all data pointers are DEVICE pointers owned by the caller; the library
 * owns only the uploaded tables and its workspaces.  Calls on one handle are
 * stream-ordered and not re-entrant.  Every function returns 0 on success or a
 * negative error code; omg_last_error() gives the message (thread-local).
 * Nothing throws across this boundary.
 */
#ifndef OMG_B200_H
#define OMG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMG_ABI_VERSION 6

/* CSR list of polynomial terms per output slot:
 *   out[s] = sum_{t in [ptr[s],ptr[s+1])} coef[t] * V[cidx[t]]
 *            * prod_{k<width} x_ext[xi[t*width+k]]   ( * lam_ext[lrow[t]] if lrow )
 * x_ext = [x, 1, mids], lam_ext = [lambda (m rows), obj_factor, mu (n_mid)].
 *
 * Intermediates ("mids", n_mid >= 0): polynomials of x shared by many rows (the
 * graph nodes CasADi shares in the reference's expression graph; Quadrotor3D's
 * acceleration product splines, quadrotor3d.py:91-126).  They are the outputs
 * m .. m+n_mid-1 of G; rows are affine in them.  The J term list then has
 * nnz_jx >= nnz_j slots: [0,nnz_j) the constraint Jacobian (direct part),
 * then A = d row/d mid, then C = d mid/d x, and
 *     J[s] += sum_{e in [jp_ptr[s],jp_ptr[s+1])} Jx[jp_a[e]] * Jx[jp_c[e]]
 *     mu_l  = sum_{e in [mu_ptr[l],mu_ptr[l+1])} lambda[mu_row[e]] * Jx[mu_slot[e]]
 * give the chain rule for the Jacobian and the multipliers of the mids' own
 * Hessians (W terms with lrow = m+1+l).  The coefficient of a mid in a row may
 * depend on x (a hyperplane normal times an integrated position: Dubins, AGV, trailer;
 * dubins.py:235-251) or on another mid (the steering-rate rows of the bicycle,
 * bicycle.py:115-124): the A slots are then functions of x_ext and the W term list
 * carries nnz_wx extra slots after its nnz_w regular ones, with values Wx:
 *     X[l,k]   = sum_i lambda_i d2 row_i / d mid_l d x_k       (cross slots)
 *     M[l1,l2] = sum_i lambda_i d2 row_i / d mid_l1 d mid_l2   (l1 >= l2)
 * Their contribution X^T C + C^T X + C^T M C to the Hessian is gathered per H position:
 *     H[xq_h[e]] += sum_{r in [xq_ptr[e],xq_ptr[e+1])} Wx[xq_w[r]] * Jx[xq_a[r]] * Jx[xq_b[r]]
 * where xq_b[r] = -1 stands for a factor 1 (cross terms; listed twice on the diagonal). */
typedef struct omg_termlist {
  int32_t n_out, n_terms, width;
  const int32_t* ptr;   /* [n_out+1] */
  const double*  coef;  /* [n_terms] */
  const int32_t* cidx;  /* [n_terms] index into the parameter tape V */
  const int32_t* xi;    /* [n_terms*width] indices into x_ext (n = constant 1) */
  const int32_t* lrow;  /* [n_terms] or NULL */
} omg_termlist;

/* Lowered NLP  min f(x,p)  s.t.  lbg <= g(x,p) <= ubg  (HOST pointers; copied
 * by omg_problem_create).  Built by omg_tools_b200/basics/lowering.py. */
typedef struct omg_tables {
  int32_t abi_version;
  int32_t n, m, n_par, n_v, degree;
  /* parameter tape: V[0]=1, V[1..n_par]=p, entry e -> V[1+n_par+e] =
   * func( sum_t coef[t]*V[f0]*V[f1]*V[f2]*V[f3] ), evaluated level by level */
  int32_t n_tape, n_tape_terms, n_levels;
  const int32_t* tape_func;   /* [n_tape] 0 id,1 inv,2 ge0,3 gt0,4 sin,5 cos,6 sqrt */
  const int32_t* tape_ptr;    /* [n_tape+1] */
  const double*  tape_coef;   /* [n_tape_terms] */
  const int32_t* tape_fac;    /* [n_tape_terms*4] */
  const int32_t* level_ptr;   /* [n_levels+1] entry ranges per level */
  omg_termlist G, F, DF, J, W;
  /* Jacobian pattern, slots sorted by (row, col) */
  int32_t nnz_j;
  const int32_t* jrow; const int32_t* jcol; const int32_t* jrow_ptr; /* [m+1] */
  /* intermediates (see omg_termlist); n_mid = 0: all of this is unused */
  int32_t n_mid, nnz_jx, n_jp, n_mu;
  const int32_t* jp_ptr; const int32_t* jp_a; const int32_t* jp_c;   /* [nnz_j+1],[n_jp] */
  const int32_t* mu_ptr; const int32_t* mu_row; const int32_t* mu_slot; /* [n_mid+1],[n_mu] */
  /* Lagrangian-Hessian pattern (lower triangle) and its position in H */
  int32_t nnz_w;
  const int32_t* wrow; const int32_t* wcol; const int32_t* w2h;
  /* condensed KKT pattern H = W + J^T Sigma J: gather lists of slot pairs */
  int32_t nnz_h, n_hp;
  const int32_t* hrow; const int32_t* hcol; const int32_t* hp_ptr; /* [nnz_h+1] */
  const int32_t* hp_s1; const int32_t* hp_s2; const int32_t* hp_row; /* [n_hp] */
  /* default bounds: rows with lbg==ubg are the structural equality rows */
  const double* lbg; const double* ubg;
  /* condensed KKT K = [[H,Jc^T],[Jc,-dc I]] = L S L^T: fill-reducing symmetric
   * permutation, signs S, lower envelope (row i stored from env_first[i], a
   * multiple of 8, to i; row kkt_n = right-hand side) and the rows reached by
   * each 8-column panel (lowering.build_kkt_structure) */
  int32_t kkt_n, kkt_n_eq, env_size, n_panel_rows, max_panel_rows;
  const int32_t* kkt_eq_rows;   /* [kkt_n_eq] constraint rows that are equalities */
  const int32_t* kkt_pos_var;   /* [n] permuted index of variable j */
  const int32_t* kkt_pos_eq;    /* [kkt_n_eq] permuted index of equality row k */
  const int32_t* kkt_sign;      /* [kkt_n] +1 / -1 */
  const int32_t* env_first;     /* [kkt_n+1] */
  const int32_t* env_ptr;       /* [kkt_n+2] */
  const int32_t* kkt_hdst;      /* [nnz_h] envelope offset of H position q */
  const int32_t* kkt_jdst;      /* [nnz_j] envelope offset of J slot (eq rows) or -1 */
  const int32_t* kkt_diag;      /* [kkt_n] envelope offset of the diagonal */
  const int32_t* kkt_panel_ptr; /* [n_panels+1] */
  const int32_t* kkt_panel_rows;/* [n_panel_rows] */
  /* extra Hessian products (see omg_termlist); nnz_wx = 0: unused.  W.n_out = nnz_w + nnz_wx */
  int32_t nnz_wx, n_xq, n_xp;
  const int32_t* xq_h;          /* [n_xq] H position */
  const int32_t* xq_ptr;        /* [n_xq+1] */
  const int32_t* xq_w;          /* [n_xp] extra W slot (0-based within the extra slots) */
  const int32_t* xq_a;          /* [n_xp] J slot of C = d mid / d x */
  const int32_t* xq_b;          /* [n_xp] second J slot of C, or -1 */
} omg_tables;

/* Interior-point options; defaults = the reference's IPOPT settings
 * (problem.py:57-60) + IPOPT defaults.  Fill with omg_default_options first. */
typedef struct omg_options {
  double tol, constr_viol_tol, dual_inf_tol, compl_inf_tol;
  double mu_init, bound_push, bound_frac, mult_bound_push, bound_relax_factor;
  double scaling_max_gradient;
  int32_t max_iter;
  int32_t trace;          /* 1: record per-iteration diagnostics (debug) */
  /* feasibility restart, the stand-in for IPOPT's restoration phase: when the
   * filter line search fails, keep x, re-centre the slacks (push restart_push),
   * zero the multipliers, clear the filter and continue from mu = restart_mu;
   * at most max_restarts times, then Restoration_Failed. */
  int32_t max_restarts;
  int32_t soft_resto;     /* 1: IPOPT's soft restoration -- when the filter rejects every trial step,
                           * accept the step if it reduces the primal-dual error by 1e-4 (default 1) */
  double restart_mu, restart_push;
  /* inertia test of the factorisation K = L S L^T.  0 (default): IPOPT's -- S takes the
   * sign of every pivot as it comes and the step is accepted when the NUMBER of negative
   * pivots equals the number of equality rows (Sylvester); 1: the stricter positional
   * test (variables +, equality rows -) of the first kernel versions, which over-
   * regularises problems with non-convex constraints. */
  int32_t inertia_mode;
  int32_t reserved;
} omg_options;

/* per-instance status codes (mapped to IPOPT strings in solver/b200.py) */
enum {
  OMG_SOLVE_SUCCEEDED = 0,
  OMG_MAX_ITER_EXCEEDED = 1,
  OMG_RESTORATION_FAILED = 2,
  OMG_ERROR_IN_STEP_COMPUTATION = 3,
  OMG_INVALID_NUMBER_DETECTED = 4,
  OMG_INFEASIBLE_PROBLEM_DETECTED = 5
};

typedef struct omg_problem omg_problem;   /* opaque handle */

void omg_default_options(omg_options* opt);

/* Upload tables to `device`, size workspaces.  Returns NULL on error. */
omg_problem* omg_problem_create(const omg_tables* tables,
                                const omg_options* opt, int device);
void omg_problem_destroy(omg_problem* h);
int  omg_set_options(omg_problem* h, const omg_options* opt);

/* Solve B independent instances of the structure (DEVICE pointers, row-major
 * [B][.] ; lbg/ubg are [m] if bounds_shared else [B][m]; lam_g0 may be NULL).
 * stream: a cudaStream_t (NULL = default stream).  Asynchronous. */
int omg_solve_batch(omg_problem* h, int32_t B,
                    const double* x0, const double* p,
                    const double* lbg, const double* ubg, int32_t bounds_shared,
                    const double* lam_g0,
                    double* x, double* lam_g, double* f,
                    int32_t* status, int32_t* iters, void* stream);

/* Same call with HOST buffers: H2D of inputs, solve, D2H of results,
 * synchronous.  This is the call the reference-facing plugin times end to end. */
int omg_solve_batch_host(omg_problem* h, int32_t B,
                         const double* x0, const double* p,
                         const double* lbg, const double* ubg,
                         int32_t bounds_shared, const double* lam_g0,
                         double* x, double* lam_g, double* f,
                         int32_t* status, int32_t* iters);

/* Feasibility phase, the fallback the host runs on instances that came back with
 * OMG_RESTORATION_FAILED before solving them once more (stand-in for the feasibility
 * part of IPOPT's restoration phase, which the reference relies on implicitly through
 * nlpsol(...,'ipopt',...), optilayer.py:55-60 / problem.py:113-128): up to max_steps
 * Levenberg-Marquardt steps on the violation v(x) = g(x,p) - clip(g(x,p), lbg, ubg),
 * (Jv^T Jv + lam I) dx = -Jv^T v, lam from 1e-3, /10 on an accepted step, x10 (at most
 * 12 times) on a rejected one; stops when max|v| <= 1e-8.  DEVICE pointers; x [B][n],
 * viol [B] (max |v| at the returned point), steps [B].  Asynchronous on `stream`. */
int omg_feas_batch(omg_problem* h, int32_t B,
                   const double* x0, const double* p,
                   const double* lbg, const double* ubg, int32_t bounds_shared,
                   int32_t max_steps, double* x, double* viol, int32_t* steps,
                   void* stream);
/* Same call with HOST buffers (synchronous). */
int omg_feas_batch_host(omg_problem* h, int32_t B,
                        const double* x0, const double* p,
                        const double* lbg, const double* ubg, int32_t bounds_shared,
                        int32_t max_steps, double* x, double* viol, int32_t* steps);

/* Receding-horizon warm start: x[b, off:off+len*ncol] <- T (len x len) applied
 * to each of the ncol columns, for n_blocks spline variables (DEVICE x,
 * in place).  offs/lens/ncols: HOST int arrays [n_blocks]; T: HOST array of the
 * n_blocks row-major matrices, concatenated.  Asynchronous on `stream`; the block
 * descriptor is uploaded on the first call and whenever it changes (the host arrays are
 * copied before the call returns). */
int omg_shift_batch(omg_problem* h, int32_t B, double* x,
                    int32_t n_blocks, const int32_t* offs, const int32_t* lens,
                    const int32_t* ncols, const double* T, void* stream);

/* Debug trace of the last omg_solve_batch with opt.trace=1: per iteration of
 * instance 0, 8 doubles {iter, f, constr_inf, dual_inf, mu, E0, alpha, delta_w}.
 * Copies up to max_rows rows to HOST `out`; returns the number of rows. */
int omg_get_trace(omg_problem* h, double* out, int32_t max_rows);

/* Introspection */
int omg_get_info(omg_problem* h, int32_t* n, int32_t* m, int32_t* n_par,
                 int32_t* smem_bytes, int32_t* ctas_per_sm, int32_t* n_sm);
/* Which kernel family serves this problem and its structure, as one line of text (e.g.
 * "sparse LDL^T: N=200 nnz(L)=3861 levels=29 root=36 pairs=41088 nt=128 ctas/SM=4 smem=53104"
 * or "envelope kernels (intermediates)").  Valid until the handle is destroyed. */
const char* omg_structure_info(omg_problem* h);
/* Device time (ms) and kernel-launch count of the last omg_solve_batch,
 * measured with CUDA events on the caller's stream. */
int omg_last_timing(omg_problem* h, float* kernel_ms, int32_t* launches);

/* Batched spline sampling (post-solve trajectory extraction, reference
 * Vehicle.store -> sample_splines, omgtools/vehicles/vehicle.py:250-300,
 * spline_extra.py:406-410; C++ twin Vehicle.cpp:112-190): for each of n_blocks
 * spline variables (offset, basis length, columns) apply the HOST matrix
 * S_blk [nsamp x len] (precomputed basis / derivative rows) to every column:
 * out[b] = concat_blk( [col][sample] ), DEVICE x [B][n] and out [B][sum nsamp*ncols].
 * Asynchronous on `stream`; descriptor and S are uploaded on the first call of the host
 * thread and whenever they change (copied before the call returns). */
int omg_sample_batch(int32_t B, int32_t n, const double* x, int32_t n_blocks,
                     const int32_t* offs, const int32_t* lens, const int32_t* ncols,
                     const int32_t* nsamp, const double* S, double* out, void* stream);

/* Non-ideal state prediction for a batch (DEVICE pointers): integrate the vehicle ODE from
 * state0 [B x n_state] over `steps` samples of the planned input trajectory
 * inputs [B x (steps+1) x n_input] with classical RK4, result in stateT [B x n_state].
 * Reference: Vehicle.predict / integrate_ode (vehicle.py:302-337, 412-423), C++ twin
 * Vehicle::predict / integrate (export/vehicles/Vehicle.cpp:61-110; stages 1-3 use
 * input[i], stage 4 input[i+1]).  Unlike the C++ twin, which evaluates every step's
 * stages at the initial state, the running state is used (identical for the holonomic
 * integrator model).  model: 0 integrator (state' = input: Holonomic, Holonomic1D/3D),
 * 1 Quadrotor3D (8 states, 3 inputs; quadrotor3d.py:308-312), 2 planar Quadrotor
 * (5 states, 2 inputs; quadrotor.py:154-157). */
int omg_integrate_rk4(int32_t model, int32_t B, int32_t n_state, int32_t n_input,
                      const double* state0, const double* inputs, double sample_time,
                      int32_t steps, double* stateT, void* stream);

/* ADMM consensus step for n_agents agents on the current device (DEVICE pointers):
 * closed-form z-update, lambda-update and squared residuals of the reference's
 * ADMM updater (omgtools/problems/admm.py:117-168 construct_upd_z/update_z,
 * 248-266 upd_l, 268-307 upd_res; C++ twin ADMMPoint2Point::update2,
 * ADMMPoint2Point.cpp:213-265).  nsh shared coefficients per agent (n_spl
 * splines of L coefficients), n_nghb neighbours; PzT [nz x nz] is the TRANSPOSED
 * consensus projector (nz = nsh*(1+n_nghb)), c [n_agents x nz] its affine part,
 * Tf/Tb [L x L] the first-knot shift and its inverse.  z_*, l_* are updated in
 * place, res [n_agents x 3] receives (pr, dr, cr) squared.  The neighbour
 * exchange itself is the caller's NCCL step (problems/admm_gpu.py). */
int omg_admm_zl_update(int32_t n_agents, int32_t nsh, int32_t n_nghb, int32_t L,
                       const double* PzT, const double* c, const double* Tf,
                       const double* Tb, double rho,
                       const double* x_i, const double* x_j,
                       double* z_i, double* z_ij, double* l_i, double* l_ij,
                       double* res, void* stream);

/* ---- multi-GPU ADMM: the consensus exchange inside the boundary ----------------------------
 * The reference "communicates" by copying neighbours' fields in one process
 * (omgtools/problems/admm.py:468-475); its C++ export leaves the transport to the user
 * (update1 returns x_var and takes z_ji/l_ji, update2 takes x_j and returns z_ij/l_ij/
 * residuals: export/point2point/admm/ADMMPoint2Point.cpp:107-117, 213-265; the test shuffles
 * the vectors by hand, export/tests/formation/test.cpp:159-187).  Here one process per GPU
 * holds a contiguous slice of the agents and the transport is NCCL over NVLink:
 *
 *   omg_comm_unique_id   rank 0: 128 opaque bytes to hand to every rank (ncclGetUniqueId)
 *   omg_comm_create      every rank: join the communicator on `device` (ncclCommInitRank);
 *                        n_ranks = 1 needs no id and no NCCL
 *   omg_admm_exchange_x  after the x-update: x_j[i][k] <- x_i of agent nghb[i][k]
 *                        (ncclAllGather of the local x_i + an index kernel)
 *   omg_admm_zl_update_dist  omg_admm_zl_update, then ncclAllReduce of the three residual
 *                        sums and the second exchange z_ji[i][k] <- z_ij[nghb[i][k]][back[i][k]]
 *                        (likewise l) -- one stream-ordered call, no host synchronisation
 *
 * nghb / back are DEVICE int32 [n_local x n_nghb]: global agent id of neighbour k of local agent
 * i, and the position of agent i in that neighbour's own list.  Every rank holds n_local agents
 * (rank r: agents r*n_local ..).  NCCL is bound at run time (dlopen of libnccl.so.2, the copy a
 * host process such as PyTorch already loaded if any): the library has no link-time dependency
 * on it, and a single-GPU caller never touches it. */
typedef struct omg_comm omg_comm;
int omg_comm_unique_id(void* id128);
omg_comm* omg_comm_create(const void* id128, int32_t n_ranks, int32_t rank, int32_t device);
void omg_comm_destroy(omg_comm* c);
int omg_admm_exchange_x(omg_comm* c, int32_t n_local, int32_t nsh, int32_t n_nghb,
                        const int32_t* nghb, const double* x_i, double* x_j, void* stream);
int omg_admm_zl_update_dist(omg_comm* c, int32_t n_local, int32_t nsh, int32_t n_nghb, int32_t L,
                            const double* PzT, const double* cvec, const double* Tf,
                            const double* Tb, double rho, const double* x_i, const double* x_j,
                            double* z_i, double* z_ij, double* l_i, double* l_ij, double* res,
                            const int32_t* nghb, const int32_t* back, double* z_ji, double* l_ji,
                            double* res_total, void* stream);

/* Table files: the on-disk form of omg_tables, the stand-in for the nlp.c / nlp.so
 * bundle that the reference's exporter writes for its C++ runtime
 * (omgtools/export/export_p2p.py:43-60 -> Point2Point.cpp:80-91 nlpsol("problem",
 * "ipopt","nlp.so")).  A file is "OMGTBL\0\0", int32 abi_version, int32 record
 * count, then named records {char name[24]; int32 dtype (0 int32, 1 float64);
 * int32 pad; int64 count; data}.  Written by omg_tools_b200.solver.b200.save_tables;
 * omg_tables_read returns a heap object that owns its arrays (release it with
 * omg_tables_free) or NULL (omg_last_error).  Host only, no GPU needed. */
omg_tables* omg_tables_read(const char* path);
void omg_tables_free(omg_tables* tables);

const char* omg_last_error(void);
int omg_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OMG_B200_H */
