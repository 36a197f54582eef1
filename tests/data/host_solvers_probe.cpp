// Test helper: the PRODUCT's host-side solver templates (mg_ic_code_b200/host/ChomboSolvers.H: BiCGStabSolver<T>, MultiGrid<T>)
// instantiated on the CPU over a mock operator -- a 1-D variable-coefficient Helmholtz problem on std::vector<double> -- so
// that their control flow is exercised without a GPU.  Prints one JSON line; tests/test_host_solvers_cpu.py compares it with
// the numpy restatement of the same algorithms (tests/amr_twin.py).
//   L phi = a_i phi_i - (phi_{i-1} - 2 phi_i + phi_{i+1}) / h^2, homogeneous Dirichlet (ghost = -near), n cells
#include <cstdio>
#include <vector>

#include "ChomboSolvers.H"

typedef std::vector<double> Vec;

class Helm1D : public MGLevelOp<Vec> {
public:
  int n;
  double h;
  Vec a, lambda;
  Helm1D(int n_, double h_, const Vec &a_) : n(n_), h(h_), a(a_), lambda(n_) {
    for (int i = 0; i < n; i++) lambda[i] = 1.0 / (a[i] + 2.0 / (h * h));
  }
  double at(const Vec &x, int i) const { return (i < 0) ? -x[0] : (i >= n ? -x[n - 1] : x[i]); }
  double L(const Vec &x, int i) const { return a[i] * x[i] - ((at(x, i - 1) - 2.0 * x[i]) + at(x, i + 1)) / (h * h); }
  void residual(Vec &lhs, const Vec &phi, const Vec &rhs, bool) override { for (int i = 0; i < n; i++) lhs[i] = rhs[i] - L(phi, i); }
  void applyOp(Vec &lhs, const Vec &phi, bool) override { for (int i = 0; i < n; i++) lhs[i] = L(phi, i); }
  void preCond(Vec &cor, const Vec &res) override {          // lambda * r, then two sweeps (VariableCoeffPoissonOperator.cpp:72-104)
    for (int i = 0; i < n; i++) cor[i] = res[i] * lambda[i];
    relax(cor, res, 2);
  }
  void relax(Vec &e, const Vec &r, int iterations) override {   // red-black Gauss-Seidel
    for (int it = 0; it < iterations; it++)
      for (int colour = 0; colour < 2; colour++)
        for (int i = colour; i < n; i += 2) e[i] = e[i] - lambda[i] * (L(e, i) - r[i]);
  }
  void restrictResidual(Vec &resCoarse, Vec &phiFine, const Vec &rhsFine) override {
    for (int i = 0; i < n / 2; i++) resCoarse[i] = 0.5 * ((rhsFine[2 * i] - L(phiFine, 2 * i)) + (rhsFine[2 * i + 1] - L(phiFine, 2 * i + 1)));
  }
  void prolongIncrement(Vec &phi, const Vec &coarse) override { for (int i = 0; i < n; i++) phi[i] += coarse[i / 2]; }
  void createCoarser(Vec &coarse, const Vec &fine, bool) override { coarse.assign(fine.size() / 2, 0.0); }
  void create(Vec &lhs, const Vec &rhs) override { lhs.assign(rhs.size(), 0.0); }
  void assign(Vec &lhs, const Vec &rhs) override { lhs = rhs; }
  Real dotProduct(const Vec &x, const Vec &y) override { double s = 0; for (size_t i = 0; i < x.size(); i++) s += x[i] * y[i]; return s; }
  void incr(Vec &lhs, const Vec &x, Real s) override { for (size_t i = 0; i < x.size(); i++) lhs[i] += s * x[i]; }
  void axby(Vec &lhs, const Vec &x, const Vec &y, Real p, Real q) override { for (size_t i = 0; i < x.size(); i++) lhs[i] = p * x[i] + q * y[i]; }
  void scale(Vec &lhs, const Real &s) override { for (double &v : lhs) v *= s; }
  Real norm(const Vec &x, int ord) override {
    double s = 0;
    for (double v : x) s = ord == 0 ? std::max(s, std::fabs(v)) : s + (ord == 1 ? std::fabs(v) : v * v);
    return ord == 2 ? std::sqrt(s) : s;
  }
  void setToZero(Vec &lhs) override { std::fill(lhs.begin(), lhs.end(), 0.0); }
};

// coefficients of depth d = arithmetic mean of 2^d fine cells (the factory's CoarseAverage); stops at minCells cells
class Helm1DFactory : public MGLevelOpFactory<Vec> {
public:
  int n, minCells;
  double h;
  Vec a;
  MGLevelOp<Vec> *MGnewOp(const ProblemDomain &, int depth, bool) override {
    const int c = 1 << depth;
    if (n % c || n / c < minCells) return NULL;
    Vec ac(n / c);
    for (int i = 0; i < n / c; i++) { double s = 0; for (int q = 0; q < c; q++) s += a[i * c + q]; ac[i] = s / c; }
    return new Helm1D(n / c, h * c, ac);
  }
};

static void print_vec(const char *name, const Vec &v) {
  std::printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); i++) std::printf("%s%.17g", i ? ", " : "", v[i]);
  std::printf("]");
}

int main() {
  const int n = 64;
  Helm1DFactory F;
  F.n = n; F.minCells = 4; F.h = 1.0 / n;
  F.a.resize(n);
  Vec rhs(n);
  for (int i = 0; i < n; i++) { F.a[i] = 1.0 + 30.0 * ((i * 37) % 11) / 11.0; rhs[i] = std::sin(0.3 * i) + 0.01 * i; }
  std::printf("{");
  {  // BiCGStab preconditioned by the operator's preCond, the way the bottom solver runs (max-norm, like the outer solver)
    Helm1D op(n, F.h, F.a);
    BiCGStabSolver<Vec> s;
    s.define(&op, false);
    s.m_verbosity = 0; s.m_normType = 0; s.m_eps = 1e-10; s.m_imax = 100;
    Vec phi(n, 0.0);
    s.solve(phi, rhs);
    std::printf("\"bicgstab_iterations\": %d, \"bicgstab_status\": %d, ", s.m_iterations, s.m_exitStatus);
    print_vec("bicgstab_history", s.m_history); std::printf(", ");
    print_vec("bicgstab_phi", phi); std::printf(", ");
    // an unreachable tolerance: the hang / restart logic has to give up with status 3 after m_numRestarts restarts
    BiCGStabSolver<Vec> g;
    g.define(&op, false);
    g.m_verbosity = 0; g.m_normType = 0; g.m_eps = 1e-30; g.m_reps = 1e-30; g.m_imax = 400;
    Vec phi2(n, 0.0);
    g.solve(phi2, rhs);
    std::printf("\"giveup_iterations\": %d, \"giveup_status\": %d, ", g.m_iterations, g.m_exitStatus);
  }
  {  // MultiGrid<T>::oneCycle, V(2,2), bottom = relax(2) + BiCGStab
    BiCGStabSolver<Vec> bottom;
    bottom.m_verbosity = 0;
    MultiGrid<Vec> mg;
    ProblemDomain dom;
    mg.define(F, &bottom, dom);
    mg.m_pre = mg.m_post = mg.m_bottom = 2;
    Vec e(n, 0.0), r(n);
    Vec hist;
    for (int cyc = 0; cyc < 6; cyc++) {
      mg.m_op[0]->residual(r, e, rhs, true);
      hist.push_back(mg.m_op[0]->norm(r, 0));
      mg.oneCycle(e, rhs);
    }
    mg.m_op[0]->residual(r, e, rhs, true);
    hist.push_back(mg.m_op[0]->norm(r, 0));
    std::printf("\"mg_depth\": %d, ", mg.m_depth);
    print_vec("mg_history", hist); std::printf(", ");
    print_vec("mg_e", e); std::printf(", ");
  }
  print_vec("a", F.a); std::printf(", ");
  print_vec("rhs", rhs);
  std::printf("}\n");
  return 0;
}
