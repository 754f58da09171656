// Native caller of libomgb200.so: the C++ twin of the reference's exported runtime
// (omgtools/export/point2point/Point2Point.cpp:80-91 loads "nlp.so" with nlpsol and
// calls problem(args) in Point2Point::solve, lines 207-231).  Here the deployable
// artefact is a table file written by omg_tools_b200.solver.b200.save_tables; no Python
// and no CasADi at run time.
//
//   g++ -O2 -I include examples/native/native_solve.cpp -o native_solve \
//       -L omg_tools_b200/csrc -lomgb200 -Wl,-rpath,$PWD/omg_tools_b200/csrc
//   ./native_solve problem.omgtbl x0.f64 p.f64 B x_out.f64
//
// x0.f64 / p.f64: raw little-endian doubles, B rows of n / n_par values.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "omg_b200.h"

static bool read_doubles(const char* path, std::vector<double>& v) {
  FILE* fp = fopen(path, "rb");
  if (!fp) return false;
  const size_t got = fread(v.data(), sizeof(double), v.size(), fp);
  fclose(fp);
  return got == v.size();
}

int main(int argc, char** argv) {
  if (argc != 6) { fprintf(stderr, "usage: %s tables x0 p B x_out\n", argv[0]); return 2; }
  omg_tables* tb = omg_tables_read(argv[1]);
  if (!tb) { fprintf(stderr, "tables: %s\n", omg_last_error()); return 1; }
  const int B = atoi(argv[4]);
  std::vector<double> x0((size_t)B * tb->n), p((size_t)B * tb->n_par), x((size_t)B * tb->n),
      lam((size_t)B * tb->m), f(B);
  std::vector<int32_t> status(B), iters(B);
  if (!read_doubles(argv[2], x0) || !read_doubles(argv[3], p)) { fprintf(stderr, "bad input files\n"); return 1; }
  omg_options opt;
  omg_default_options(&opt);
  omg_problem* h = omg_problem_create(tb, &opt, 0);
  if (!h) { fprintf(stderr, "create: %s\n", omg_last_error()); return 1; }
  // bounds: the structural defaults of the tables, shared by all instances
  if (omg_solve_batch_host(h, B, x0.data(), p.data(), tb->lbg, tb->ubg, 1, nullptr, x.data(),
                           lam.data(), f.data(), status.data(), iters.data()) != 0) {
    fprintf(stderr, "solve: %s\n", omg_last_error()); return 1;
  }
  // instances whose line search gave up: feasibility phase from where they stopped, one
  // more solve from there (what B200Solver.solve_batch does; IPOPT's restoration phase in
  // the reference's runtime happens inside the same nlpsol call, Point2Point.cpp:219)
  std::vector<int> bad;
  for (int b = 0; b < B; ++b) if (status[b] == OMG_RESTORATION_FAILED) bad.push_back(b);
  if (!bad.empty()) {
    const size_t nb = bad.size(), n = tb->n, np_ = tb->n_par, m = tb->m;
    std::vector<double> xs(nb * n), ps(nb * np_), x1(nb * n), viol(nb), x2(nb * n), lam2(nb * m), f2(nb);
    std::vector<int32_t> steps(nb), st2(nb), it2(nb);
    for (size_t k = 0; k < nb; ++k) {
      std::copy(x.begin() + bad[k] * n, x.begin() + (bad[k] + 1) * n, xs.begin() + k * n);
      std::copy(p.begin() + bad[k] * np_, p.begin() + (bad[k] + 1) * np_, ps.begin() + k * np_);
    }
    if (omg_feas_batch_host(h, (int)nb, xs.data(), ps.data(), tb->lbg, tb->ubg, 1, 30, x1.data(), viol.data(),
                            steps.data()) != 0 ||
        omg_solve_batch_host(h, (int)nb, x1.data(), ps.data(), tb->lbg, tb->ubg, 1, nullptr, x2.data(),
                             lam2.data(), f2.data(), st2.data(), it2.data()) != 0) {
      fprintf(stderr, "feasibility phase: %s\n", omg_last_error()); return 1;
    }
    for (size_t k = 0; k < nb; ++k) {
      std::copy(x2.begin() + k * n, x2.begin() + (k + 1) * n, x.begin() + bad[k] * n);
      status[bad[k]] = st2[k]; iters[bad[k]] += it2[k]; f[bad[k]] = f2[k];
    }
  }
  for (int b = 0; b < B; ++b) printf("instance %d status %d iters %d f %.12g\n", b, status[b], iters[b], f[b]);
  FILE* fo = fopen(argv[5], "wb");
  if (fo) { fwrite(x.data(), sizeof(double), x.size(), fo); fclose(fo); }
  omg_problem_destroy(h);
  omg_tables_free(tb);
  return 0;
}
