/*
 * mgic_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Line-faithful CPU restatement of the multigrid hot path of
 * eugenealim/MG_IC_code, on the reference's own BOXED layout (one ghosted
 * FAB per box, explicit exchange + BC fill before every colour pass), so it
 * pays the same structural costs as the reference and doubles as the timed
 * "reference CPU path" (OpenMP over boxes == the reference's one-MPI-rank-per
 * -core box parallelism, jobscript.pbs:3,13).
 *
 * PARITY (see mgic_oracle.h): source terms, parameters, ParseBC, the operator
 * class's orchestration and the factory are pinned bit for bit to the reference's
 * own C++ and to a mechanical translation of its .ChF kernels (oracle/_ref);
 * UNPINNED is everything tagged [Chombo], which restates Chombo 3.2
 * (GNUmakefile:12) -- not under /root/reference -- from its published algorithm.
 *
 * Build: g++ -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp (see Makefile).
 * -ffp-contract=off keeps the source evaluation order (no FMA contraction) so
 * that the CUDA kernels (compiled -fmad=false) can be compared BIT-EXACTLY.
 */
#include "mgic_oracle.h"

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double Real;

namespace {

// ---------------------------------------------------------------------------
// [Chombo] BoxTools minimal restatement: Box, FArrayBox, layout, LevelData
// ---------------------------------------------------------------------------
static inline int coarsen1(int i, int r) { return (i < 0) ? -((-i - 1) / r) - 1 : i / r; }

struct Box {
  int lo[3], hi[3];
  Box() { for (int d = 0; d < 3; d++) { lo[d] = 0; hi[d] = -1; } }
  Box(const int *l, const int *h) { for (int d = 0; d < 3; d++) { lo[d] = l[d]; hi[d] = h[d]; } }
  Box(int l0, int l1, int l2, int h0, int h1, int h2) {
    lo[0] = l0; lo[1] = l1; lo[2] = l2; hi[0] = h0; hi[1] = h1; hi[2] = h2;
  }
  bool empty() const { return hi[0] < lo[0] || hi[1] < lo[1] || hi[2] < lo[2]; }
  int size(int d) const { return hi[d] - lo[d] + 1; }
  long numPts() const { return empty() ? 0 : (long)size(0) * size(1) * size(2); }
  bool contains(const Box &o) const {
    for (int d = 0; d < 3; d++) if (o.lo[d] < lo[d] || o.hi[d] > hi[d]) return false;
    return true;
  }
  bool contains(int i, int j, int k) const { return i >= lo[0] && i <= hi[0] && j >= lo[1] && j <= hi[1] && k >= lo[2] && k <= hi[2]; }
  Box operator&(const Box &o) const {
    Box r;
    for (int d = 0; d < 3; d++) { r.lo[d] = std::max(lo[d], o.lo[d]); r.hi[d] = std::min(hi[d], o.hi[d]); }
    return r;
  }
  Box grown(int g) const { Box r = *this; for (int d = 0; d < 3; d++) { r.lo[d] -= g; r.hi[d] += g; } return r; }
  Box coarsened(int r) const {
    Box c; for (int d = 0; d < 3; d++) { c.lo[d] = coarsen1(lo[d], r); c.hi[d] = coarsen1(hi[d], r); } return c;
  }
  Box shifted(int d, int s) const { Box r = *this; r.lo[d] += s; r.hi[d] += s; return r; }
  // [Chombo] Box::coarsenable(refrat): refine(coarsen(b)) == b
  bool coarsenable(int r) const {
    for (int d = 0; d < 3; d++) {
      int cl = coarsen1(lo[d], r), ch = coarsen1(hi[d], r);
      if (cl * r != lo[d] || (ch + 1) * r - 1 != hi[d]) return false;
    }
    return true;
  }
};

// [Chombo] adjCellBox(b, dir, side, len): the len-thick cell box just outside b
static Box adjCellBox(const Box &b, int dir, int side /*-1 lo,+1 hi*/, int len) {
  Box r = b;
  if (side < 0) { r.hi[dir] = b.lo[dir] - 1; r.lo[dir] = b.lo[dir] - len; }
  else { r.lo[dir] = b.hi[dir] + 1; r.hi[dir] = b.hi[dir] + len; }
  return r;
}

struct FAB {
  Box b; int nc = 0; long s1 = 0, s2 = 0, sc = 0; std::vector<Real> d;
  void define(const Box &bx, int ncomp) {
    b = bx; nc = ncomp; s1 = bx.size(0); s2 = s1 * bx.size(1); sc = s2 * bx.size(2);
    d.assign((size_t)sc * nc, 0.0);
  }
  inline long idx(int i, int j, int k) const { return (i - b.lo[0]) + s1 * (j - b.lo[1]) + s2 * (long)(k - b.lo[2]); }
  inline Real &operator()(int i, int j, int k, int c = 0) { return d[idx(i, j, k) + sc * c]; }
  inline const Real &operator()(int i, int j, int k, int c = 0) const { return d[idx(i, j, k) + sc * c]; }
  void setVal(Real v) { std::fill(d.begin(), d.end(), v); }
};

// region is in the DESTINATION's index space; the source cell is region - shift (shift != 0: periodic image)
struct CopyItem { int from, to; Box region; int shift[3]; };

struct Layout {
  std::vector<Box> boxes;
  Box domain;
  bool periodic = false;
  std::vector<CopyItem> exFace1;  // exchangeDefine(grids, Unit) + trimEdges  (Factory.cpp:83-84)
  std::vector<CopyItem> exFull3;  // exchangeDefine(grids, 3*Unit)            (Main_PoissonSolver.cpp:200-201)

  void buildCopier(std::vector<CopyItem> &out, int ghost, bool facesOnly) const {
    out.clear();
    // spatial bins of size = min box extent
    int bs = 1 << 30;
    for (auto &b : boxes) for (int d = 0; d < 3; d++) bs = std::min(bs, b.size(d));
    if (boxes.empty()) return;
    auto key = [&](int bi, int bj, int bk) -> long long {
      return ((long long)(bi + 4096) << 42) | ((long long)(bj + 4096) << 21) | (long long)(bk + 4096);
    };
    std::unordered_map<long long, std::vector<int>> bins;
    for (int n = 0; n < (int)boxes.size(); n++) {
      const Box &b = boxes[n];
      for (int bk = coarsen1(b.lo[2], bs); bk <= coarsen1(b.hi[2], bs); bk++)
        for (int bj = coarsen1(b.lo[1], bs); bj <= coarsen1(b.hi[1], bs); bj++)
          for (int bi = coarsen1(b.lo[0], bs); bi <= coarsen1(b.hi[0], bs); bi++)
            bins[key(bi, bj, bk)].push_back(n);
    }
    for (int to = 0; to < (int)boxes.size(); to++) {
      std::vector<Box> regions;
      if (facesOnly) {
        for (int d = 0; d < 3; d++) { regions.push_back(adjCellBox(boxes[to], d, -1, ghost)); regions.push_back(adjCellBox(boxes[to], d, +1, ghost)); }
      } else {
        regions.push_back(boxes[to].grown(ghost));
      }
      for (auto &g : regions) {
        std::vector<int> cand;
        for (int bk = coarsen1(g.lo[2], bs); bk <= coarsen1(g.hi[2], bs); bk++)
          for (int bj = coarsen1(g.lo[1], bs); bj <= coarsen1(g.hi[1], bs); bj++)
            for (int bi = coarsen1(g.lo[0], bs); bi <= coarsen1(g.hi[0], bs); bi++) {
              auto it = bins.find(key(bi, bj, bk));
              if (it != bins.end()) for (int n : it->second) cand.push_back(n);
            }
        std::sort(cand.begin(), cand.end());
        cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
        for (int from : cand) {
          if (from == to) continue;
          Box r = g & boxes[from];
          if (!r.empty()) out.push_back({from, to, r, {0, 0, 0}});
        }
        if (periodic) {
          // [Chombo] Copier::exchangeDefine on a periodic ProblemDomain: ghost cells outside the domain are filled from
          // the periodic image of the valid cells (every non-zero shift by whole domain lengths)
          for (int s2 = -1; s2 <= 1; s2++)
            for (int s1 = -1; s1 <= 1; s1++)
              for (int s0 = -1; s0 <= 1; s0++) {
                if (!s0 && !s1 && !s2) continue;
                const int sh[3] = {s0 * domain.size(0), s1 * domain.size(1), s2 * domain.size(2)};
                Box gs = g;  // the ghost region seen in the source's index space
                for (int d = 0; d < 3; d++) { gs.lo[d] -= sh[d]; gs.hi[d] -= sh[d]; }
                if ((gs & domain).empty()) continue;
                for (int from = 0; from < (int)boxes.size(); from++) {
                  Box r = gs & boxes[from];
                  if (r.empty()) continue;
                  for (int d = 0; d < 3; d++) { r.lo[d] += sh[d]; r.hi[d] += sh[d]; }
                  out.push_back({from, to, r, {sh[0], sh[1], sh[2]}});
                }
              }
        }
      }
    }
  }
};

struct LevelData {
  const Layout *lay = nullptr; int nc = 0, ng = 0; std::vector<FAB> fab;
  void define(const Layout *l, int ncomp, int ghost) {
    lay = l; nc = ncomp; ng = ghost; fab.resize(l->boxes.size());
#pragma omp parallel for schedule(static)
    for (int n = 0; n < (int)fab.size(); n++) fab[n].define(l->boxes[n].grown(ghost), ncomp);
  }
  bool defined() const { return lay != nullptr; }
  int size() const { return (int)fab.size(); }
};

// [Chombo] LevelData::exchange(interval, copier): valid -> neighbour ghost
static void exchange(LevelData &ld, const std::vector<CopyItem> &items) {
#pragma omp parallel for schedule(static)
  for (int m = 0; m < (int)items.size(); m++) {
    const CopyItem &it = items[m];
    const FAB &src = ld.fab[it.from]; FAB &dst = ld.fab[it.to];
    for (int c = 0; c < ld.nc; c++)
      for (int k = it.region.lo[2]; k <= it.region.hi[2]; k++)
        for (int j = it.region.lo[1]; j <= it.region.hi[1]; j++)
          for (int i = it.region.lo[0]; i <= it.region.hi[0]; i++)
            dst(i, j, k, c) = src(i - it.shift[0], j - it.shift[1], k - it.shift[2], c);
  }
}

// ---------------------------------------------------------------------------
// Fortran kernels (Source/VariableCoeffPoissonOperatorF.ChF, SetLevelDataF.ChF)
// ---------------------------------------------------------------------------
struct View {  // a (const) FRA argument
  Real *p; int lo[3]; long s1, s2, sc;
  inline Real &operator()(int i, int j, int k, int n = 0) const {
    return p[(i - lo[0]) + s1 * (j - lo[1]) + s2 * (long)(k - lo[2]) + sc * n];
  }
};
static inline View mkview(const Real *p, int l0, int l1, int l2, int h0, int h1, int h2) {
  View v; v.p = const_cast<Real *>(p); v.lo[0] = l0; v.lo[1] = l1; v.lo[2] = l2;
  v.s1 = h0 - l0 + 1; v.s2 = v.s1 * (h1 - l1 + 1); v.sc = v.s2 * (h2 - l2 + 1); return v;
}
#define VIEW(a) mkview(a, *a##lo0, *a##lo1, *a##lo2, *a##hi0, *a##hi1, *a##hi2)
static inline View fabview(const FAB &f, const int *shift = nullptr) {
  View v; v.p = const_cast<Real *>(f.d.data());
  for (int d = 0; d < 3; d++) v.lo[d] = f.b.lo[d] - (shift ? shift[d] : 0);
  v.s1 = f.s1; v.s2 = f.s2; v.sc = f.sc; return v;
}

// The 7-point "laplacian term" CHF_DTERM block common to all four kernels
// (VariableCoeffPoissonOperatorF.ChF:111-120, 219-228, 322-330, 415-424):
// each parenthesised group left to right, groups added in x,y,z order.
static inline Real S7(const View &u, int i, int j, int k, int n) {
  const Real two = 2.0;
  return ((u(i + 1, j, k, n) + u(i - 1, j, k, n)) - two * u(i, j, k, n)) +
         ((u(i, j + 1, k, n) + u(i, j - 1, k, n)) - two * u(i, j, k, n)) +
         ((u(i, j, k + 1, n) + u(i, j, k - 1, n)) - two * u(i, j, k, n));
}

// GSRBHELMHOLTZVC3D  -- VariableCoeffPoissonOperatorF.ChF:56-139
static void k_gsrb(const View &dpsi, const View &rhs, const Box &region, Real dx, Real alpha,
                   const View &aCoef, Real beta, const View &bCoef, const View &lambda, int redBlack,
                   int ncomp) {
  const Real one = 1.0;
  Real dxinv = one / (dx * dx);                                        // :89
  for (int n = 0; n < ncomp; n++)
    for (int k = region.lo[2]; k <= region.hi[2]; k++)                 // :93
      for (int j = region.lo[1]; j <= region.hi[1]; j++) {             // :96
        int imin = region.lo[0];                                       // :98
        int indtot = imin + j + k;                                     // :99
        imin = imin + std::abs((indtot + redBlack) % 2);               // :104 (Fortran mod keeps sign; abs)
        int imax = region.hi[0];
        for (int i = imin; i <= imax; i += 2) {                        // :106
          Real lofdpsi = alpha * aCoef(i, j, k, n) * dpsi(i, j, k, n); // :107-108
          Real ldpsi = S7(dpsi, i, j, k, n);                           // :111-120
          ldpsi = ldpsi * dxinv * bCoef(i, j, k, n);                   // :122
          lofdpsi = lofdpsi - beta * ldpsi;                            // :124
          dpsi(i, j, k, n) = dpsi(i, j, k, n) - lambda(i, j, k, n) * (lofdpsi - rhs(i, j, k, n));  // :127-128
        }
      }
}

// VCCOMPUTEOP3D -- VariableCoeffPoissonOperatorF.ChF:181-237
static void k_op(const View &lof, const View &dpsi, Real alpha, const View &aCoef, Real beta,
                 const View &bCoef, const Box &region, Real dx, int ncomp) {
  Real dxinv = 1.0 / (dx * dx);                                        // :208
  for (int n = 0; n < ncomp; n++)
    for (int k = region.lo[2]; k <= region.hi[2]; k++)
      for (int j = region.lo[1]; j <= region.hi[1]; j++)
        for (int i = region.lo[0]; i <= region.hi[0]; i++) {
          lof(i, j, k, n) = alpha * aCoef(i, j, k, n) * dpsi(i, j, k, n);  // :211-212
          Real ldpsi = S7(dpsi, i, j, k, n);                           // :216-225
          ldpsi = ldpsi * dxinv * beta * bCoef(i, j, k, n);            // :227
          lof(i, j, k, n) = lof(i, j, k, n) - ldpsi;                   // :229
        }
}

// VCCOMPUTERES3D -- VariableCoeffPoissonOperatorF.ChF:283-339
static void k_res(const View &res, const View &dpsi, const View &rhs, Real alpha, const View &aCoef,
                  Real beta, const View &bCoef, const Box &region, Real dx, int ncomp) {
  Real dxinv = 1.0 / (dx * dx);                                        // :311
  for (int n = 0; n < ncomp; n++)
    for (int k = region.lo[2]; k <= region.hi[2]; k++)
      for (int j = region.lo[1]; j <= region.hi[1]; j++)
        for (int i = region.lo[0]; i <= region.hi[0]; i++) {
          res(i, j, k, n) = rhs(i, j, k, n) - alpha * aCoef(i, j, k, n) * dpsi(i, j, k, n);  // :314-316
          Real ldpsi = S7(dpsi, i, j, k, n);                           // :320-329
          ldpsi = ldpsi * dxinv * beta * bCoef(i, j, k, n);            // :331
          res(i, j, k, n) = res(i, j, k, n) + ldpsi;                   // :333
        }
}

// RESTRICTRESVC3D -- VariableCoeffPoissonOperatorF.ChF:379-437 (indices already shifted by caller)
static void k_restrict(const View &res, const View &dpsi, const View &rhs, Real alpha, const View &aCoef,
                       Real beta, const View &bCoef, const Box &region, Real dx, int ncomp) {
  Real dxinv = 1.0 / (dx * dx);                                        // :401
  Real denom = 2 * 2 * 2;                                              // :402
  for (int n = 0; n < ncomp; n++)
    for (int k = region.lo[2]; k <= region.hi[2]; k++)
      for (int j = region.lo[1]; j <= region.hi[1]; j++)
        for (int i = region.lo[0]; i <= region.hi[0]; i++) {
          int ii = i / 2, jj = j / 2, kk = k / 2;                      // :406-409 (non-negative after shift)
          Real lofdpsi = alpha * aCoef(i, j, k, n) * dpsi(i, j, k, n); // :411-412
          Real ldpsi = S7(dpsi, i, j, k, n);                           // :416-425
          ldpsi = ldpsi * dxinv * beta * bCoef(i, j, k, n);            // :427
          lofdpsi = lofdpsi - ldpsi;                                   // :429
          res(ii, jj, kk, n) = res(ii, jj, kk, n) + (rhs(i, j, k, n) - lofdpsi) / denom;  // :431-432
        }
}

// GETLAPLACIANPSIF -- SetLevelDataF.ChF:15-58
static void k_lap(const View &l, const View &psi, Real dx, const Box &box) {
  for (int k = box.lo[2]; k <= box.hi[2]; k++)
    for (int j = box.lo[1]; j <= box.hi[1]; j++)
      for (int i = box.lo[0]; i <= box.hi[0]; i++) {
        l(i, j, k) = 0.0;                                              // :26
        for (int d0 = 0; d0 < 3; d0++) {                               // :28
          int ii0 = (d0 == 0), ii1 = (d0 == 1), ii2 = (d0 == 2);
          Real dpsidxdx = 1.0 / dx / dx *                              // :35-39
                          (+1.0 * psi(i - ii0, j - ii1, k - ii2) - 2.0 * psi(i, j, k) +
                           1.0 * psi(i + ii0, j + ii1, k + ii2));
          l(i, j, k) = l(i, j, k) + dpsidxdx;                          // :52
        }
      }
}

// GETRHOGRADPHIF -- SetLevelDataF.ChF:65-103
static void k_rho(const View &r, const View &phi, Real dx, const Box &box) {
  for (int k = box.lo[2]; k <= box.hi[2]; k++)
    for (int j = box.lo[1]; j <= box.hi[1]; j++)
      for (int i = box.lo[0]; i <= box.hi[0]; i++) {
        r(i, j, k) = 0.0;                                              // :75
        for (int d0 = 0; d0 < 3; d0++) {
          int ii0 = (d0 == 0), ii1 = (d0 == 1), ii2 = (d0 == 2);
          Real dphidx = 0.5 / dx * (+phi(i + ii0, j + ii1, k + ii2) - phi(i - ii0, j - ii1, k - ii2));  // :83-86
          r(i, j, k) = r(i, j, k) + 0.5 * dphidx * dphidx;             // :98
        }
      }
}

// [Chombo] AMRPoissonOpF.ChF PROLONG: phi(i,j,k) += coarse(i/m, j/m, k/m)  (indices shifted >= 0)
static void k_prolong(const View &phi, const View &coarse, const Box &region, int m, int ncomp) {
  for (int n = 0; n < ncomp; n++)
    for (int k = region.lo[2]; k <= region.hi[2]; k++)
      for (int j = region.lo[1]; j <= region.hi[1]; j++)
        for (int i = region.lo[0]; i <= region.hi[0]; i++)
          phi(i, j, k, n) = phi(i, j, k, n) + coarse(i / m, j / m, k / m, n);
}

// ---------------------------------------------------------------------------
// Physical BCs: Source/SetBCs.cpp:49-131 (+ [Chombo] BCFunc DiriBC/NeumBC, order 1)
// ---------------------------------------------------------------------------
struct BCSpec { int lo[3], hi[3]; Real value; };

// [Chombo] DiriBC order 1: ghost = 2*v - near ; NeumBC: ghost = near + sign*dx*v
static void bcFill(FAB &state, const Box &valid, int dir, int side, int type, Real dx, bool homog, Real value) {
  Box toRegion = adjCellBox(valid, dir, side, 1) & state.b;
  if (toRegion.empty()) return;
  Real v = homog ? 0.0 : value;
  int off[3] = {0, 0, 0}; off[dir] = -side;
  for (int n = 0; n < state.nc; n++)
    for (int k = toRegion.lo[2]; k <= toRegion.hi[2]; k++)
      for (int j = toRegion.lo[1]; j <= toRegion.hi[1]; j++)
        for (int i = toRegion.lo[0]; i <= toRegion.hi[0]; i++) {
          Real nearVal = state(i + off[0], j + off[1], k + off[2], n);
          if (type == 0) state(i, j, k, n) = 2 * v - nearVal;          // linearInterp
          else state(i, j, k, n) = nearVal + side * dx * v;            // NeumBC
        }
}

// ParseBC -- Source/SetBCs.cpp:49-131
static void ParseBC(FAB &state, const Box &valid, const Box &domain, bool periodic, const BCSpec &bc, Real dx,
                    bool homog) {
  if (domain.contains(state.b)) return;                                // :51
  for (int i = 0; i < 3; i++) {
    if (periodic) continue;                                            // :63
    Box gLo = adjCellBox(valid, i, -1, 1), gHi = adjCellBox(valid, i, +1, 1);
    if (!domain.contains(gLo)) {                                       // :66
      if (bc.lo[i] == 1) bcFill(state, valid, i, -1, 1, dx, homog, bc.value);
      else if (bc.lo[i] == 0) bcFill(state, valid, i, -1, 0, dx, homog, bc.value);
      else if (bc.lo[i] == 2) {}
      else { fprintf(stderr, "bogus bc flag low side\n"); abort(); }   // :94
    }
    if (!domain.contains(gHi)) {                                       // :98
      if (bc.hi[i] == 1) bcFill(state, valid, i, +1, 1, dx, homog, bc.value);
      else if (bc.hi[i] == 0) bcFill(state, valid, i, +1, 0, dx, homog, bc.value);
      else if (bc.hi[i] == 2) {}
      else { fprintf(stderr, "bogus bc flag high side\n"); abort(); }  // :123
    }
  }
}

// ---------------------------------------------------------------------------
// VariableCoeffPoissonOperator (Source/VariableCoeffPoissonOperator.cpp) on one MG depth
// ---------------------------------------------------------------------------
struct Op {
  Layout lay; Real dx = 0, alpha = 0, beta = 0; BCSpec bc;
  LevelData *aCoef = nullptr, *bCoef = nullptr;   // shared at depth 0 (Factory.cpp:194-197)
  LevelData aOwn, bOwn, lambda;
  bool lambdaNeedsResetting = true;
  // an AMR level > 0 (a union of boxes that does not cover the domain): the coarser level's spacing
  bool hasCoarser = false; Real dxCrse = 0;

  // [Chombo] AMRPoissonOp::homogeneousCFInterp -> AMRPoissonOpF.ChF INTERPHOMO (restated from the published
  // algorithm; Chombo is not vendored): every face ghost cell on a coarse-fine interface gets the value at x = 2*dx of
  // the parabola through the second interior cell (at 0), the first interior cell (at dx) and a ZERO coarse value at
  // x2 = (3*dx + dxCrse)/2.  All face ghost cells inside the domain are filled; the exchange that follows at every call
  // site (VariableCoeffPoissonOperator.cpp:163,301) overwrites the ones a neighbouring box of the level covers.
  void homogeneousCFInterp(LevelData &phi) {
    if (!hasCoarser) return;
    const Real x1 = dx;
    const Real x2 = 0.5 * (3. * x1 + dxCrse);
    const Real denom = 1.0 - ((x1 + x2) / x1);
    const Real idenom = 1 / (denom);
    const Real x = 2. * x1;
    const Real xsquared = x * x;
    const Real m1 = 1 / (x1 * x1);
    const Real m2 = 1 / (x1 * (x1 - x2));
    const Real q1 = 1 / (x1 - x2);
    const Real q2 = x1 + x2;
#pragma omp parallel for schedule(static)
    for (int n = 0; n < phi.size(); n++) {
      FAB &f = phi.fab[n];
      for (int dir = 0; dir < 3; dir++)
        for (int side = -1; side <= 1; side += 2) {
          const Box region = adjCellBox(lay.boxes[n], dir, side, 1) & lay.domain & f.b;
          if (region.empty()) continue;
          int ii[3] = {0, 0, 0}; ii[dir] = side;
          for (int c = 0; c < phi.nc; c++)
            for (int k = region.lo[2]; k <= region.hi[2]; k++)
              for (int j = region.lo[1]; j <= region.hi[1]; j++)
                for (int i = region.lo[0]; i <= region.hi[0]; i++) {
                  const Real pa = f(i - 2 * ii[0], j - 2 * ii[1], k - 2 * ii[2], c);
                  const Real pb = f(i - ii[0], j - ii[1], k - ii[2], c);
                  const Real a = ((pb - pa) * m1 - (pb)*m2) * idenom;
                  const Real b = (pb)*q1 - a * q2;
                  f(i, j, k, c) = a * xsquared + b * x + pa;
                }
        }
    }
  }

  // resetLambda -- VariableCoeffPoissonOperator.cpp:220-249
  void resetLambda() {
    if (!lambdaNeedsResetting) return;
#pragma omp parallel for schedule(static)
    for (int n = 0; n < lambda.size(); n++) {
      FAB &l = lambda.fab[n]; const FAB &a = aCoef->fab[n];
      Real plus = 2.0 * 3 * beta / (dx * dx);                          // :241
      for (size_t q = 0; q < l.d.size(); q++) {
        Real v = a.d[q];                                               // copy  :234
        v = v * alpha;                                                 // mult  :235
        v = v + plus;                                                  // plus  :241
        l.d[q] = 1.0 / v;                                              // invert(1.0) :244
      }
    }
    lambdaNeedsResetting = false;
  }
  void computeLambda() { lambda.define(&lay, aCoef->nc, 0); resetLambda(); }  // :252-260

  void applyBC(LevelData &x, bool homog) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < x.size(); n++) ParseBC(x.fab[n], lay.boxes[n], lay.domain, lay.periodic, bc, dx, homog);
  }

  // [Chombo] QuadCFInterp::coarseFineInterp(phiFine, phiCoarse), ratio 2 (restated from the published algorithm, SURVEY
  // App. B.10; Chombo is not vendored): every face ghost cell on a coarse-fine interface gets the value at x = 2h of the
  // parabola through the second interior cell (0), the first interior cell (h) and phiStar at the coarse cell centre
  // h(nref+3)/2 (QuadCFInterpF.ChF QUADINTERP).  phiStar = the coarse field taken to the ghost cell's tangential position
  // by a second-order Taylor expansion: centred first and second differences (one-sided second-order ones next to a
  // domain face) in each tangential direction, plus the mixed term from the four diagonal coarse cells (dropped where one
  // of them is outside the domain).  `crse` holds the coarse level on `cdom` (its valid cells; copyTo of Chombo).
  // As with homogeneousCFInterp every face ghost cell inside the domain is filled; the exchange of the caller
  // (applyOpI / residualI :131 / :48) overwrites those that a neighbouring box covers.
  void quadCFInterp(LevelData &phi, const FAB &crse, const Box &cdom) {
    const int nref = 2;
    const Real h = dx, H = dxCrse;
#pragma omp parallel for schedule(static)
    for (int n = 0; n < phi.size(); n++) {
      FAB &f = phi.fab[n];
      for (int dir = 0; dir < 3; dir++)
        for (int side = -1; side <= 1; side += 2) {
          const Box region = adjCellBox(lay.boxes[n], dir, side, 1) & lay.domain & f.b;
          if (region.empty()) continue;
          int ii[3] = {0, 0, 0}; ii[dir] = side;
          const int t1 = (dir + 1) % 3, t2 = (dir + 2) % 3;
          const int ta = std::min(t1, t2), tb = std::max(t1, t2);   // tangential directions in ascending order
          for (int k = region.lo[2]; k <= region.hi[2]; k++)
            for (int j = region.lo[1]; j <= region.hi[1]; j++)
              for (int i = region.lo[0]; i <= region.hi[0]; i++) {
                const int ivf[3] = {i, j, k};
                int ivc[3]; for (int d = 0; d < 3; d++) ivc[d] = coarsen1(ivf[d], nref);
                auto C = [&](int o_a, int o_b) -> Real {
                  int q[3] = {ivc[0], ivc[1], ivc[2]}; q[ta] += o_a; q[tb] += o_b;
                  return crse(q[0], q[1], q[2]);
                };
                const Real c0 = C(0, 0);
                Real phistar = c0;
                Real xt[2];
                for (int w = 0; w < 2; w++) {
                  const int t = w ? tb : ta;
                  const Real x = (ivf[t] + 0.5) * h - (ivc[t] + 0.5) * H;
                  xt[w] = x;
                  const bool hasLo = ivc[t] - 1 >= cdom.lo[t], hasHi = ivc[t] + 1 <= cdom.hi[t];
                  auto Ct = [&](int o) { return w ? C(0, o) : C(o, 0); };
                  Real d1, d2;
                  if (hasLo && hasHi) {
                    d1 = (Ct(1) - Ct(-1)) / (2.0 * H);
                    d2 = ((Ct(1) - 2.0 * c0) + Ct(-1)) / (H * H);
                  } else if (hasHi) {
                    d1 = ((4.0 * Ct(1) - 3.0 * c0) - Ct(2)) / (2.0 * H);
                    d2 = ((c0 - 2.0 * Ct(1)) + Ct(2)) / (H * H);
                  } else {
                    d1 = ((3.0 * c0 - 4.0 * Ct(-1)) + Ct(-2)) / (2.0 * H);
                    d2 = ((c0 - 2.0 * Ct(-1)) + Ct(-2)) / (H * H);
                  }
                  phistar = phistar + (d1 * x + 0.5 * d2 * x * x);
                }
                const bool corners = ivc[ta] - 1 >= cdom.lo[ta] && ivc[ta] + 1 <= cdom.hi[ta] && ivc[tb] - 1 >= cdom.lo[tb] &&
                                     ivc[tb] + 1 <= cdom.hi[tb];
                if (corners) {
                  const Real mixed = (((C(1, 1) - C(1, -1)) - C(-1, 1)) + C(-1, -1)) / (4.0 * H * H);
                  phistar = phistar + mixed * xt[0] * xt[1];
                }
                // QUADINTERP
                const Real x = 2.0 * h;
                const Real pa = f(i - 2 * ii[0], j - 2 * ii[1], k - 2 * ii[2]);
                const Real pb = f(i - ii[0], j - ii[1], k - ii[2]);
                const Real a = (2.0 / h / h) * ((2.0 * phistar + pa * (nref + 1.0)) - pb * (nref + 3.0)) / (nref * nref + 4.0 * nref + 3.0);
                const Real b = (pb - pa) / h - a * h;
                f(i, j, k) = (pa + b * x) + a * x * x;
              }
        }
    }
  }
  // [Chombo] AMRPoissonOp::AMROperatorNF / AMRResidualNF on the finest level: coarse-fine interpolation from the coarser
  // level's field, then the level operator (applyOpI / residualI of the reference)
  void amrOperatorNF(LevelData &lhs, LevelData &phi, const FAB &crse, const Box &cdom, bool homogPhysBC) {
    quadCFInterp(phi, crse, cdom);
    applyOp(lhs, phi, homogPhysBC);
  }
  void amrResidualNF(LevelData &lhs, LevelData &phi, const FAB &crse, const Box &cdom, const LevelData &rhs, bool homogPhysBC) {
    quadCFInterp(phi, crse, cdom);
    residual(lhs, phi, rhs, homogPhysBC);
  }

  // levelGSRB -- VariableCoeffPoissonOperator.cpp:273-332
  void gsrbColor(LevelData &dpsi, const LevelData &rhs, int whichPass) {
    homogeneousCFInterp(dpsi);                                         // :296 (no-op without a coarser AMR level)
    exchange(dpsi, lay.exFace1);                                       // :301
    applyBC(dpsi, true);                                               // :307-310
#pragma omp parallel for schedule(static)
    for (int n = 0; n < dpsi.size(); n++)                              // :313-330
      k_gsrb(fabview(dpsi.fab[n]), fabview(rhs.fab[n]), lay.boxes[n], dx, alpha, fabview(aCoef->fab[n]), beta,
             fabview(bCoef->fab[n]), fabview(lambda.fab[n]), whichPass, dpsi.nc);
  }
  void levelGSRB(LevelData &dpsi, const LevelData &rhs) {
    resetLambda();                                                     // :283
    for (int whichPass = 0; whichPass <= 1; whichPass++) gsrbColor(dpsi, rhs, whichPass);  // :290
  }
  // [Chombo] AMRPoissonOp::relax, s_relaxMode == 1
  void relax(LevelData &e, const LevelData &r, int iterations) {
    for (int i = 0; i < iterations; i++) levelGSRB(e, r);
  }
  // residualI -- VariableCoeffPoissonOperator.cpp:30-67
  void residual(LevelData &lhs, LevelData &dpsi, const LevelData &rhs, bool homog) {
    applyBC(dpsi, homog);                                              // :44-46
    exchange(dpsi, lay.exFace1);                                       // :48
#pragma omp parallel for schedule(static)
    for (int n = 0; n < dpsi.size(); n++)                              // :50-66
      k_res(fabview(lhs.fab[n]), fabview(dpsi.fab[n]), fabview(rhs.fab[n]), alpha, fabview(aCoef->fab[n]), beta,
            fabview(bCoef->fab[n]), lay.boxes[n], dx, dpsi.nc);
  }
  // applyOpI / applyOpNoBoundary -- VariableCoeffPoissonOperator.cpp:106-149
  void applyOp(LevelData &lhs, LevelData &dpsi, bool homog) {
    applyBC(dpsi, homog);                                              // :116-118
    exchange(dpsi, lay.exFace1);                                       // :131
#pragma omp parallel for schedule(static)
    for (int n = 0; n < dpsi.size(); n++)                              // :133-148
      k_op(fabview(lhs.fab[n]), fabview(dpsi.fab[n]), alpha, fabview(aCoef->fab[n]), beta, fabview(bCoef->fab[n]),
           lay.boxes[n], dx, dpsi.nc);
  }
  // restrictResidual -- VariableCoeffPoissonOperator.cpp:151-194
  void restrictResidual(LevelData &resCoarse, LevelData &dpsiFine, const LevelData &rhsFine) {
    homogeneousCFInterp(dpsiFine);                                     // :156
    applyBC(dpsiFine, true);                                           // :158-161
    exchange(dpsiFine, lay.exFace1);                                   // :163
#pragma omp parallel for schedule(static)
    for (int n = 0; n < dpsiFine.size(); n++) {
      const Box &region = lay.boxes[n];
      int iv[3] = {region.lo[0], region.lo[1], region.lo[2]};          // :173
      int civ[3] = {coarsen1(iv[0], 2), coarsen1(iv[1], 2), coarsen1(iv[2], 2)};  // :174
      resCoarse.fab[n].setVal(0.0);                                    // :177
      Box sregion = region;
      for (int d = 0; d < 3; d++) { sregion.lo[d] -= iv[d]; sregion.hi[d] -= iv[d]; }
      k_restrict(fabview(resCoarse.fab[n], civ), fabview(dpsiFine.fab[n], iv), fabview(rhsFine.fab[n], iv), alpha,
                 fabview(aCoef->fab[n], iv), beta, fabview(bCoef->fab[n], iv), sregion, dx, dpsiFine.nc);  // :188-192
    }
  }
  // [Chombo] AMRPoissonOp::prolongIncrement (mgref = 2)
  void prolongIncrement(LevelData &phi, const LevelData &coarse) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < phi.size(); n++) {
      const Box &region = lay.boxes[n];
      int iv[3] = {region.lo[0], region.lo[1], region.lo[2]};
      int civ[3] = {coarsen1(iv[0], 2), coarsen1(iv[1], 2), coarsen1(iv[2], 2)};
      Box sregion = region;
      for (int d = 0; d < 3; d++) { sregion.lo[d] -= iv[d]; sregion.hi[d] -= iv[d]; }
      k_prolong(fabview(phi.fab[n], iv), fabview(coarse.fab[n], civ), sregion, 2, phi.nc);
    }
  }
  // preCond -- VariableCoeffPoissonOperator.cpp:72-104
  void preCond(LevelData &dpsi, const LevelData &rhs) {
    resetLambda();                                                     // :90
#pragma omp parallel for schedule(static)
    for (int n = 0; n < dpsi.size(); n++) {                            // :94-101
      const Box &g = rhs.fab[n].b & dpsi.fab[n].b;
      FAB &p = dpsi.fab[n]; const FAB &r = rhs.fab[n]; const FAB &l = lambda.fab[n];
      for (int k = g.lo[2]; k <= g.hi[2]; k++)
        for (int j = g.lo[1]; j <= g.hi[1]; j++)
          for (int i = g.lo[0]; i <= g.hi[0]; i++) p(i, j, k) = r(i, j, k) * l(i, j, k);
    }
    relax(dpsi, rhs, 2);                                               // :103
  }

  // ---- [Chombo] AMRPoissonOp vector ops over valid cells -----------------
  void setToZero(LevelData &x) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < x.size(); n++) x.fab[n].setVal(0.0);
  }
  // assignLocal / assign: copy over the intersection of the two fab boxes
  void assign(LevelData &y, const LevelData &x) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < x.size(); n++) {
      Box g = x.fab[n].b & y.fab[n].b;
      for (int k = g.lo[2]; k <= g.hi[2]; k++)
        for (int j = g.lo[1]; j <= g.hi[1]; j++)
          for (int i = g.lo[0]; i <= g.hi[0]; i++) y.fab[n](i, j, k) = x.fab[n](i, j, k);
    }
  }
  // incr: y += s*x  (FArrayBox::plus(x, scale) over the intersection)
  void incr(LevelData &y, const LevelData &x, Real s) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < x.size(); n++) {
      Box g = x.fab[n].b & y.fab[n].b;
      for (int k = g.lo[2]; k <= g.hi[2]; k++)
        for (int j = g.lo[1]; j <= g.hi[1]; j++)
          for (int i = g.lo[0]; i <= g.hi[0]; i++) y.fab[n](i, j, k) = y.fab[n](i, j, k) + s * x.fab[n](i, j, k);
    }
  }
  void scale(LevelData &y, Real s) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < y.size(); n++) for (auto &v : y.fab[n].d) v = v * s;
  }
  // dotProduct: sum over boxes of FArrayBox::dotProduct over the valid box
  Real dot(const LevelData &a, const LevelData &b) {
    std::vector<Real> part(a.size(), 0.0);
#pragma omp parallel for schedule(static)
    for (int n = 0; n < a.size(); n++) {
      const Box &g = lay.boxes[n]; Real s = 0.0;
      for (int k = g.lo[2]; k <= g.hi[2]; k++)
        for (int j = g.lo[1]; j <= g.hi[1]; j++)
          for (int i = g.lo[0]; i <= g.hi[0]; i++) s += a.fab[n](i, j, k) * b.fab[n](i, j, k);
      part[n] = s;
    }
    Real val = 0.0; for (Real s : part) val += s; return val;
  }
  // norm(LevelData, interval, p): p=0 max|x|; p=2 sqrt(sum_boxes (fabnorm2)^2)
  Real norm(const LevelData &a, int ord) {
    std::vector<Real> part(a.size(), 0.0);
#pragma omp parallel for schedule(static)
    for (int n = 0; n < a.size(); n++) {
      const Box &g = lay.boxes[n]; Real s = 0.0;
      for (int k = g.lo[2]; k <= g.hi[2]; k++)
        for (int j = g.lo[1]; j <= g.hi[1]; j++)
          for (int i = g.lo[0]; i <= g.hi[0]; i++) {
            Real v = a.fab[n](i, j, k);
            if (ord == 0) s = std::max(s, std::fabs(v));
            else if (ord == 1) s += std::fabs(v);
            else s += v * v;
          }
      part[n] = s;
    }
    Real val = 0.0;
    if (ord == 0) { for (Real s : part) val = std::max(val, s); return val; }
    if (ord == 1) { for (Real s : part) val += s; return val; }
    for (Real s : part) { Real fn = std::sqrt(s); val += fn * fn; }     // FArrayBox::norm(p=2) then squared
    return std::sqrt(val);
  }
};

// ---------------------------------------------------------------------------
// [Chombo] BiCGStabSolver<T>::solve restated (defaults: imax 80, eps 1e-6,
// reps 1e-12, hang 1e-8, small 1e-30, numRestarts 5, normType 2)
// ---------------------------------------------------------------------------
struct BiCGParams {
  int imax = 80; Real eps = 1.0e-6, reps = 1.0e-12, hang = 1.0e-8, small = 1.0e-30;
  int numRestarts = 5, normType = 2; bool homogeneous = false; int verbosity = 0;
};

// LinOp concept: create(T&, const T& like), setToZero, assignLocal, incr, scale, dot, norm,
// residual(lhs, phi, rhs, homog), applyOp(lhs, phi, homog), preCond(cor, res)
template <class T, class LinOp>
static int bicgstab(LinOp &op, T &a_phi, const T &a_rhs, const BiCGParams &P, int *exitStatus,
                    std::vector<Real> *hist) {
  T r, r_tilde, e, p, p_tilde, s_tilde, t, v;
  op.create(r, a_rhs); op.create(r_tilde, a_rhs); op.create(e, a_phi); op.create(p, a_rhs);
  op.create(p_tilde, a_phi); op.create(s_tilde, a_phi); op.create(t, a_rhs); op.create(v, a_rhs);
  int recount = 0;
  op.setToZero(r);
  op.residual(r, a_phi, a_rhs, P.homogeneous);
  op.assignLocal(r_tilde, r);
  op.setToZero(e); op.setToZero(p_tilde); op.setToZero(s_tilde);
  int i = 0;
  Real rho[4] = {0, 0, 0, 0};
  Real norm[2];
  norm[0] = op.norm(r, P.normType);
  Real initial_norm = norm[0], initial_rnorm = norm[0];
  norm[1] = norm[0];
  Real alpha[2] = {0, 0}, beta[2] = {0, 0}, omega[2] = {0, 0};
  bool init = true; int restarts = 0;
  if (exitStatus) *exitStatus = -1;
  if (hist) hist->push_back(norm[0]);
  if (P.verbosity >= 5) printf("      BiCGStab:: initial Residual norm = %.15e\n", initial_norm);

  while ((i < P.imax && norm[0] > P.eps * norm[1]) && (norm[1] > 0)) {
    i++;
    norm[1] = norm[0]; alpha[1] = alpha[0]; beta[1] = beta[0]; omega[1] = omega[0];
    rho[3] = rho[2]; rho[2] = rho[1];
    rho[1] = op.dot(r_tilde, r);
    if (rho[1] == 0.0) {
      op.incr(a_phi, e, 1.0);
      if (exitStatus) *exitStatus = 2;
      return i;
    }
    if (init) { op.assignLocal(p, r); init = false; }
    else {
      beta[1] = (rho[1] / rho[2]) * (alpha[1] / omega[1]);
      op.scale(p, beta[1]);
      op.incr(p, v, -beta[1] * omega[1]);
      op.incr(p, r, 1.0);
    }
    op.preCond(p_tilde, p);
    op.setToZero(v);
    op.applyOp(v, p_tilde, true);
    Real m = op.dot(r_tilde, v);
    alpha[0] = rho[1] / m;
    if (std::fabs(m) > P.small * std::fabs(rho[1])) {
      op.incr(r, v, -alpha[0]);
      norm[0] = op.norm(r, P.normType);
      op.incr(e, p_tilde, alpha[0]);
    } else {
      op.setToZero(r);
      norm[0] = 0.0;
    }
    if (norm[0] > P.eps * initial_norm && norm[0] > P.reps * initial_rnorm) {
      op.preCond(s_tilde, r);
      op.setToZero(t);
      op.applyOp(t, s_tilde, true);
      omega[0] = op.dot(t, r) / op.dot(t, t);
      op.incr(e, s_tilde, omega[0]);
      op.incr(r, t, -omega[0]);
      norm[0] = op.norm(r, P.normType);
    }
    if (hist) hist->push_back(norm[0]);
    if (P.verbosity >= 4)
      printf("      BiCGStab::     iteration = %d, error norm = %.15e, rate = %g\n", i, norm[0], norm[1] / norm[0]);
    if (norm[0] <= P.eps * initial_norm || norm[0] <= P.reps * initial_rnorm) {
      if (exitStatus) *exitStatus = 1;
      break;
    }
    if (omega[0] == 0.0 || norm[0] > (1 - P.hang) * norm[1]) {
      if (recount == 0) recount = 1;
      else {
        recount = 0;
        op.incr(a_phi, e, 1.0);
        if (restarts == P.numRestarts) {
          if (exitStatus) *exitStatus = 3;
          return i;
        }
        op.residual(r, a_phi, a_rhs, P.homogeneous);
        norm[0] = op.norm(r, P.normType);
        rho[1] = 0.0; rho[1] = 0.0; rho[2] = 0.0; rho[3] = 0.0;
        alpha[0] = 0; beta[0] = 0; omega[0] = 0;
        op.assignLocal(r_tilde, r);
        op.setToZero(e);
        restarts++;
        init = true;
      }
    }
  }
  op.incr(a_phi, e, 1.0);
  return i;
}

// LinOp adaptor for one level op (bottom solver: BiCGStabSolver<LevelData<FArrayBox>>)
struct LevelLinOp {
  Op *op;
  void create(LevelData &x, const LevelData &like) { x.define(like.lay, like.nc, like.ng); }
  void setToZero(LevelData &x) { op->setToZero(x); }
  void assignLocal(LevelData &y, const LevelData &x) { op->assign(y, x); }
  void incr(LevelData &y, const LevelData &x, Real s) { op->incr(y, x, s); }
  void scale(LevelData &y, Real s) { op->scale(y, s); }
  Real dot(const LevelData &a, const LevelData &b) { return op->dot(a, b); }
  Real norm(const LevelData &a, int ord) { return op->norm(a, ord); }
  void residual(LevelData &l, LevelData &phi, const LevelData &rhs, bool h) { op->residual(l, phi, rhs, h); }
  void applyOp(LevelData &l, LevelData &phi, bool h) { op->applyOp(l, phi, h); }
  void preCond(LevelData &c, const LevelData &r) { op->preCond(c, r); }
};

// [Chombo] CoarseAverage::averageToCoarse / averageToCoarseHarmonic (AverageF.ChF AVERAGE / AVERAGEHARMONIC):
// refScale = 1/nRef^3; sum over the nRef^3 fine cells (ii fastest); arithmetic: sum*refScale;
// harmonic: 1/(sum(1/fine)*refScale)
static void coarseAverage(LevelData &coarse, const LevelData &fine, int nRef, int type) {
  Real refScale = 1.0 / (Real)(nRef * nRef * nRef);
#pragma omp parallel for schedule(static)
  for (int n = 0; n < coarse.size(); n++) {
    const Box &cb = coarse.lay->boxes[n]; FAB &c = coarse.fab[n]; const FAB &f = fine.fab[n];
    for (int kc = cb.lo[2]; kc <= cb.hi[2]; kc++)
      for (int jc = cb.lo[1]; jc <= cb.hi[1]; jc++)
        for (int ic = cb.lo[0]; ic <= cb.hi[0]; ic++) {
          Real coarseSum = 0.0;
          for (int kk = 0; kk < nRef; kk++)
            for (int jj = 0; jj < nRef; jj++)
              for (int ii = 0; ii < nRef; ii++) {
                Real fv = f(ic * nRef + ii, jc * nRef + jj, kc * nRef + kk);
                coarseSum = coarseSum + (type == 1 ? 1.0 / fv : fv);
              }
          c(ic, jc, kc) = (type == 1) ? 1.0 / (coarseSum * refScale) : coarseSum * refScale;
        }
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// The problem: Main_PoissonSolver.cpp poissonSolve() state, single AMR level
// ---------------------------------------------------------------------------
struct orc_problem {
  orc_params P;
  Layout grids;                 // set_grids, max_level = 0: domainSplit base level (SetGrids.cpp:54-62)
  Real dx0;
  LevelData mgvars, dpsi, rhs, aCoef, bCoef;   // Main_PoissonSolver.cpp:79-88
  // MG hierarchy built by MGnewOp(depth = 0,1,...) until NULL (Factory.cpp:139-234)
  std::vector<Op *> ops;
  std::vector<LevelData> e, r, tmp;            // MultiGrid m_correction / m_residual + scratch
  int lastBottomIters = 0;
  Real constant_K = 0.0;
};

namespace {

// [Chombo] domainSplit: uniform lattice of boxes of at most max_grid_size
static void domainSplit(const Box &dom, int maxSize, std::vector<Box> &out) {
  int nb[3], base[3];
  for (int d = 0; d < 3; d++) { nb[d] = (dom.size(d) + maxSize - 1) / maxSize; base[d] = dom.size(d) / nb[d]; }
  for (int bk = 0; bk < nb[2]; bk++)
    for (int bj = 0; bj < nb[1]; bj++)
      for (int bi = 0; bi < nb[0]; bi++) {
        int b3[3] = {bi, bj, bk}; Box b;
        for (int d = 0; d < 3; d++) {
          // spread remainder over the first boxes (N divisible by max_grid_size in all configs)
          int rem = dom.size(d) - base[d] * nb[d];
          int lo = dom.lo[d] + b3[d] * base[d] + std::min(b3[d], rem);
          int sz = base[d] + (b3[d] < rem ? 1 : 0);
          b.lo[d] = lo; b.hi[d] = lo + sz - 1;
        }
        out.push_back(b);
      }
}

// cell-centre location, Source/SetLevelData.cpp:58-60: loc = (iv + 0.5)*dx - L/2
static inline void cellLoc(const orc_params &P, Real dx, int i, int j, int k, Real loc[3]) {
  int iv[3] = {i, j, k};
  for (int d = 0; d < 3; d++) {
    Real l = iv[d] + 0.5 * 1.0;
    l *= dx;
    l -= ((P.L / P.N[0]) * P.N[d]) / 2.0;   // domainLength[d] = coarsestDx * nCells[d]  (PoissonParameters.cpp:82-85); dx = the level's
    loc[d] = l;
  }
}

// get_bh_radius -- Source/SetBinaryBH.H:15-20
static inline Real get_bh_radius(Real loc_bh[3], Real off) {
  loc_bh[0] -= off;
  return std::sqrt(loc_bh[0] * loc_bh[0] + loc_bh[1] * loc_bh[1] + loc_bh[2] * loc_bh[2]);
}

// get_Aij -- Source/SetBinaryBH.H:24-52
static Real get_Aij(int i, int j, Real rbh1, Real rbh2, const Real n1[3], const Real n2[3], const Real J1[3],
                    const Real J2[3], const Real P1[3], const Real P2[3]) {
  Real epsilon[3][3][3] = {{{0.}}};
  epsilon[0][1][2] = 1.0; epsilon[1][2][0] = 1.0; epsilon[2][0][1] = 1.0;
  epsilon[0][2][1] = -1.0; epsilon[2][1][0] = -1.0; epsilon[1][0][2] = -1.0;
  Real Aij = 1.5 / rbh1 / rbh1 * (n1[i] * P1[j] + n1[j] * P1[i]) + 1.5 / rbh2 / rbh2 * (n2[i] * P2[j] + n2[j] * P2[i]);
  for (int k = 0; k < 3; k++) {
    Aij += 1.5 / rbh1 / rbh1 * (n1[i] * n1[j] - Real(i == j)) * P1[k] * n1[k] +
           1.5 / rbh2 / rbh2 * (n2[i] * n2[j] - Real(i == j)) * P2[k] * n2[k];
    for (int l = 0; l < 3; l++) {
      Aij += -3.0 / rbh1 / rbh1 / rbh1 * (epsilon[i][l][k] * n1[j] + epsilon[j][l][k] * n1[i]) * n1[l] * J1[k] -
             3.0 / rbh2 / rbh2 / rbh2 * (epsilon[i][l][k] * n2[j] + epsilon[j][l][k] * n2[i]) * n2[l] * J2[k];
    }
  }
  return Aij;
}

// set_binary_bh_psi -- Source/SetBinaryBH.H:85-99
static inline Real set_binary_bh_psi(const Real loc[3], const orc_params &P) {
  Real l1[3] = {loc[0], loc[1], loc[2]}; Real rbh1 = get_bh_radius(l1, P.bh1_offset);
  Real l2[3] = {loc[0], loc[1], loc[2]}; Real rbh2 = get_bh_radius(l2, P.bh2_offset);
  return P.bh1_bare_mass / rbh1 + P.bh2_bare_mass / rbh2;
}

// my_phi_function -- MyPhiFunction.H:11-16
static inline Real my_phi_function(const Real loc[3], Real amplitude, Real wavelength) {
  Real r2 = loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2];
  return amplitude * std::exp(-r2 / wavelength);
}

// set_m_value -- Source/SetLevelData.cpp:266-278
static inline Real m_value(const orc_params &P, Real constant_K) {
  Real Pi_field = 0.0, V_of_phi = 0.0;
  Real rho = 0.5 * Pi_field * Pi_field + V_of_phi;
  return (2.0 / 3.0) * (constant_K * constant_K) - 16.0 * M_PI * P.G_Newton * rho;
}

// A2 expression shared by set_rhs / set_a_coef (SetLevelData.cpp:110-116, 305-311)
static inline Real A2_of(const FAB &mv, int i, int j, int k) {
  return std::pow(mv(i, j, k, 1), 2.0) + std::pow(mv(i, j, k, 4), 2.0) + std::pow(mv(i, j, k, 6), 2.0) +
         2 * std::pow(mv(i, j, k, 2), 2.0) + 2 * std::pow(mv(i, j, k, 3), 2.0) + 2 * std::pow(mv(i, j, k, 5), 2.0);
}

// LinOp adaptor for the single-level MultilevelLinearOp (outer solver); defined after vcycle
struct OuterLinOp;

static void mg_cycle(orc_problem *pb, int depth, LevelData &correction, const LevelData &residual);

static int bottomSolve(orc_problem *pb, LevelData &e, const LevelData &r) {
  LevelLinOp lop{pb->ops.back()};
  BiCGParams bp; bp.homogeneous = true;  // MultiGrid::define: m_bottomSolver->define(op, true)
  int status = 0;
  return bicgstab<LevelData, LevelLinOp>(lop, e, r, bp, &status, nullptr);
}

// [Chombo] MultiGrid::cycle (m_cycle = 1, V-cycle)
static void mg_cycle(orc_problem *pb, int depth, LevelData &correction, const LevelData &residual) {
  int nd = (int)pb->ops.size();
  int S = pb->P.numMGsmooth;  // pre = post = bottom (Main_PoissonSolver.cpp:111-113)
  Op *op = pb->ops[depth];
  if (depth == nd - 1) {
    long bottomCells = op->lay.domain.numPts();
    if (bottomCells == 1) op->relax(correction, residual, 1);
    else {
      op->relax(correction, residual, S);
      pb->lastBottomIters = bottomSolve(pb, correction, residual);
    }
  } else {
    op->relax(correction, residual, S);
    op->restrictResidual(pb->r[depth + 1], correction, residual);
    pb->ops[depth + 1]->setToZero(pb->e[depth + 1]);
    mg_cycle(pb, depth + 1, pb->e[depth + 1], pb->r[depth + 1]);
    op->prolongIncrement(correction, pb->e[depth + 1]);
    op->relax(correction, residual, S);
  }
}

// [Chombo] MultilevelLinearOp<FArrayBox> restricted to one AMR level
struct OuterLinOp {
  orc_problem *pb;
  Op *op() { return pb->ops[0]; }
  void create(LevelData &x, const LevelData &like) { x.define(like.lay, like.nc, like.ng); }
  void setToZero(LevelData &x) { op()->setToZero(x); }
  void assignLocal(LevelData &y, const LevelData &x) { op()->assign(y, x); }
  void incr(LevelData &y, const LevelData &x, Real s) { op()->incr(y, x, s); }
  void scale(LevelData &y, Real s) { op()->scale(y, s); }
  Real dot(const LevelData &a, const LevelData &b) { return op()->dot(a, b); }
  Real norm(const LevelData &a, int ord) { return op()->norm(a, ord); }
  void residual(LevelData &l, LevelData &phi, const LevelData &rhs, bool h) { op()->residual(l, phi, rhs, h); }
  void applyOp(LevelData &l, LevelData &phi, bool h) { op()->applyOp(l, phi, h); }
  // MultilevelLinearOp::preCond: zero cor, then m_num_mg_iterations x AMRVCycle (single level: oneCycle -> cycle(0))
  void preCond(LevelData &cor, const LevelData &res) {
    op()->setToZero(cor);
    for (int it = 0; it < pb->P.numMGIterations; it++) mg_cycle(pb, 0, cor, res);
  }
};

static LevelData *fieldOf(orc_problem *pb, int depth, int field) {
  switch (field) {
    case ORC_F_E: return &pb->e[depth];
    case ORC_F_R: return &pb->r[depth];
    case ORC_F_A: return (depth == 0 || pb->ops.empty()) ? &pb->aCoef : pb->ops[depth]->aCoef;
    case ORC_F_B: return (depth == 0 || pb->ops.empty()) ? &pb->bCoef : pb->ops[depth]->bCoef;
    case ORC_F_LAMBDA: return &pb->ops[depth]->lambda;
    case ORC_F_TMP: return &pb->tmp[depth];
    case ORC_F_DPSI: return &pb->dpsi;
    case ORC_F_RHS: return &pb->rhs;
    default: return nullptr;
  }
}

}  // namespace

// ===========================================================================
// extern "C" API
// ===========================================================================
extern "C" {

#define BX(b) Box(*b##lo0, *b##lo1, *b##lo2, *b##hi0, *b##hi1, *b##hi2)

void orc_gsrbhelmholtzvc3d(ORC_FRA(dpsi), ORC_CFRA(rhs), ORC_BOX(region), const double *dx, const double *alpha,
                           ORC_CFRA(aCoef), const double *beta, ORC_CFRA(bCoef), ORC_CFRA(lambda),
                           const int *redBlack) {
  int ncomp = *dpsinc;
  if (ncomp != *rhsnc || ncomp != *bCoefnc) { fprintf(stderr, "MAYDAYERROR\n"); abort(); }  // :77-87
  (void)aCoefnc; (void)lambdanc;
  k_gsrb(VIEW(dpsi), VIEW(rhs), BX(region), *dx, *alpha, VIEW(aCoef), *beta, VIEW(bCoef), VIEW(lambda), *redBlack,
         ncomp);
}
void orc_vccomputeop3d(ORC_FRA(lofdpsi), ORC_CFRA(dpsi), const double *alpha, ORC_CFRA(aCoef), const double *beta,
                       ORC_CFRA(bCoef), ORC_BOX(region), const double *dx) {
  int ncomp = *dpsinc;
  if (ncomp != *lofdpsinc || ncomp != *bCoefnc) { fprintf(stderr, "MAYDAYERROR\n"); abort(); }
  (void)aCoefnc;
  k_op(VIEW(lofdpsi), VIEW(dpsi), *alpha, VIEW(aCoef), *beta, VIEW(bCoef), BX(region), *dx, ncomp);
}
void orc_vccomputeres3d(ORC_FRA(res), ORC_CFRA(dpsi), ORC_CFRA(rhs), const double *alpha, ORC_CFRA(aCoef),
                        const double *beta, ORC_CFRA(bCoef), ORC_BOX(region), const double *dx) {
  int ncomp = *dpsinc;
  if (ncomp != *resnc || ncomp != *bCoefnc) { fprintf(stderr, "MAYDAYERROR\n"); abort(); }
  (void)aCoefnc; (void)rhsnc;
  k_res(VIEW(res), VIEW(dpsi), VIEW(rhs), *alpha, VIEW(aCoef), *beta, VIEW(bCoef), BX(region), *dx, ncomp);
}
void orc_restrictresvc3d(ORC_FRA(res), ORC_CFRA(dpsi), ORC_CFRA(rhs), const double *alpha, ORC_CFRA(aCoef),
                         const double *beta, ORC_CFRA(bCoef), ORC_BOX(region), const double *dx) {
  int ncomp = *dpsinc;
  (void)aCoefnc; (void)rhsnc; (void)resnc; (void)bCoefnc;
  k_restrict(VIEW(res), VIEW(dpsi), VIEW(rhs), *alpha, VIEW(aCoef), *beta, VIEW(bCoef), BX(region), *dx, ncomp);
}
void orc_getlaplacianpsif(ORC_FRA1(lap), ORC_CFRA1(psi), const double *dx, ORC_BOX(box)) {
  k_lap(VIEW(lap), VIEW(psi), *dx, BX(box));
}
void orc_getrhogradphif(ORC_FRA1(rho), ORC_CFRA1(phi), const double *dx, ORC_BOX(box)) {
  k_rho(VIEW(rho), VIEW(phi), *dx, BX(box));
}
void orc_prolong(ORC_FRA(phi), ORC_CFRA(coarse), ORC_BOX(region), const int *m) {
  (void)coarsenc;
  k_prolong(VIEW(phi), VIEW(coarse), BX(region), *m, *phinc);
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// bench.py's reference arm: use n host threads whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

orc_problem *orc_create(const orc_params *p) {
  orc_problem *pb = new orc_problem;
  pb->P = *p;
  if (p->max_level != 0) { fprintf(stderr, "oracle: single AMR level only (max_level = 0)\n"); delete pb; return nullptr; }
  pb->grids.domain = Box(0, 0, 0, p->N[0] - 1, p->N[1] - 1, p->N[2] - 1);   // PoissonParameters.cpp:110-114
  pb->grids.periodic = p->is_periodic != 0;
  domainSplit(pb->grids.domain, p->max_grid_size, pb->grids.boxes);       // SetGrids.cpp:54-58
  pb->grids.buildCopier(pb->grids.exFace1, 1, true);
  pb->grids.buildCopier(pb->grids.exFull3, 3, false);
  pb->dx0 = p->L / p->N[0];                                                // PoissonParameters.cpp:82
  pb->mgvars.define(&pb->grids, 8, 3);                                     // Main_PoissonSolver.cpp:79-88
  pb->dpsi.define(&pb->grids, 1, 3);
  pb->rhs.define(&pb->grids, 1, 0);
  pb->aCoef.define(&pb->grids, 1, 0);
  pb->bCoef.define(&pb->grids, 1, 0);
  return pb;
}

void orc_destroy(orc_problem *pb) {
  if (!pb) return;
  for (Op *o : pb->ops) delete o;
  delete pb;
}

// set_initial_conditions -- Source/SetLevelData.cpp:32-71 (+ set_binary_bh_Aij SetBinaryBH.H:54-83)
// (one AMR level: dx is the LEVEL's spacing, P.L / P.N the coarsest level's -- Main_PoissonSolver.cpp:93 calls it per level)
static void init_conditions_level(const orc_params &P, Real dx, LevelData &mgvars, LevelData &dpsi) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < mgvars.size(); n++) {
    FAB &mv = mgvars.fab[n]; FAB &dp = dpsi.fab[n];
    const Box &b = mv.b;                                                  // ghosted box :42
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++) {
          mv(i, j, k, 0) = 1.0;                                           // :54
          if (dp.b.contains(i, j, k)) dp(i, j, k, 0) = 0.0;               // :55
          Real loc[3]; cellLoc(P, dx, i, j, k, loc);                      // :58-60
          mv(i, j, k, 7) = my_phi_function(loc, P.phi_amplitude, P.phi_wavelength);  // :63-65
          // set_binary_bh_Aij
          Real l1[3] = {loc[0], loc[1], loc[2]}; Real rbh1 = get_bh_radius(l1, P.bh1_offset);
          Real l2[3] = {loc[0], loc[1], loc[2]}; Real rbh2 = get_bh_radius(l2, P.bh2_offset);
          Real n1[3] = {l1[0] / rbh1, l1[1] / rbh1, l1[2] / rbh1};
          Real n2[3] = {l2[0] / rbh2, l2[1] / rbh2, l2[2] / rbh2};
          Real J1[3] = {0.0, 0.0, P.bh1_spin}, J2[3] = {0.0, 0.0, P.bh2_spin};
          Real P1[3] = {0.0, P.bh1_momentum, 0.0}, P2[3] = {0.0, P.bh2_momentum, 0.0};
          mv(i, j, k, 1) = get_Aij(0, 0, rbh1, rbh2, n1, n2, J1, J2, P1, P2);   // c_A11_0
          mv(i, j, k, 4) = get_Aij(1, 1, rbh1, rbh2, n1, n2, J1, J2, P1, P2);   // c_A22_0
          mv(i, j, k, 6) = get_Aij(2, 2, rbh1, rbh2, n1, n2, J1, J2, P1, P2);   // c_A33_0
          mv(i, j, k, 2) = get_Aij(0, 1, rbh1, rbh2, n1, n2, J1, J2, P1, P2);   // c_A12_0
          mv(i, j, k, 3) = get_Aij(0, 2, rbh1, rbh2, n1, n2, J1, J2, P1, P2);   // c_A13_0
          mv(i, j, k, 5) = get_Aij(1, 2, rbh1, rbh2, n1, n2, J1, J2, P1, P2);   // c_A23_0
        }
  }
}

void orc_set_initial_conditions(orc_problem *pb) { init_conditions_level(pb->P, pb->dx0, pb->mgvars, pb->dpsi); }

// set_a_coef (SetLevelData.cpp:281-325), set_b_coef (:330-340), set_rhs (:73-127) on one AMR level
static void coefs_and_rhs_level(const orc_params &P, Real dx, double constant_K, LevelData &mgvars, LevelData &rhsLD, LevelData &aLD,
                                LevelData &bLD) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < rhsLD.size(); n++) {
    FAB &mv = mgvars.fab[n]; FAB &rhs = rhsLD.fab[n]; FAB &aC = aLD.fab[n]; FAB &bC = bLD.fab[n];
    const Box &tb = rhs.b;  // no ghost cells
    FAB lap, rho; lap.define(tb, 1); rho.define(tb, 1);
    View psiV = fabview(mv); View phiV = fabview(mv); phiV.p += mv.sc * 7;   // CHF_CONST_FRA1(mv, c_psi / c_phi_0)
    k_rho(fabview(rho), phiV, dx, tb);                                      // :296-298
    for (int k = tb.lo[2]; k <= tb.hi[2]; k++)
      for (int j = tb.lo[1]; j <= tb.hi[1]; j++)
        for (int i = tb.lo[0]; i <= tb.hi[0]; i++) {
          Real loc[3]; cellLoc(P, dx, i, j, k, loc);
          Real m = m_value(P, constant_K);
          Real A2 = A2_of(mv, i, j, k);
          Real psi_bh = set_binary_bh_psi(loc, P);
          Real psi_0 = mv(i, j, k, 0) + psi_bh;
          aC(i, j, k) = -0.625 * m * std::pow(psi_0, 4.0) - A2 * std::pow(psi_0, -8.0) +
                        2.0 * M_PI * P.G_Newton * rho(i, j, k);             // :321-322
        }
    bC.setVal(1.0);                                                         // :338
    rhs.setVal(0.0);                                                        // :83
    k_lap(fabview(lap), psiV, dx, tb);                                      // :88-90
    k_rho(fabview(rho), phiV, dx, tb);                                      // :94-96
    for (int k = tb.lo[2]; k <= tb.hi[2]; k++)
      for (int j = tb.lo[1]; j <= tb.hi[1]; j++)
        for (int i = tb.lo[0]; i <= tb.hi[0]; i++) {
          Real loc[3]; cellLoc(P, dx, i, j, k, loc);
          Real m = m_value(P, constant_K);                                  // :105-106
          Real A2 = A2_of(mv, i, j, k);                                     // :110-116
          Real psi_bh = set_binary_bh_psi(loc, P);                          // :118
          Real psi_0 = mv(i, j, k, 0) + psi_bh;                             // :119
          rhs(i, j, k) = 0.125 * m * std::pow(psi_0, 5.0) - 0.125 * A2 * std::pow(psi_0, -7.0) -
                         2.0 * M_PI * P.G_Newton * rho(i, j, k) * psi_0 - lap(i, j, k);   // :121-124
        }
  }
}
void orc_set_coefs_and_rhs(orc_problem *pb, double constant_K) {
  pb->constant_K = constant_K;
  coefs_and_rhs_level(pb->P, pb->dx0, constant_K, pb->mgvars, pb->rhs, pb->aCoef, pb->bCoef);
}


// defineOperatorFactory + MultiGrid::define: MGnewOp(domain, depth) until NULL (Factory.cpp:139-234)
int orc_define_solver(orc_problem *pb) {
  const orc_params &P = pb->P;
  for (Op *o : pb->ops) delete o;
  pb->ops.clear(); pb->e.clear(); pb->r.clear(); pb->tmp.clear();
  const int s_maxCoarse = 2;   // [Chombo] AMRPoissonOp::s_maxCoarse
  for (int depth = 0;; depth++) {
    // [Chombo 3.2] MultiGrid::define: push an operator, m_depth++, ask for the next one only while
    // (m_depth < a_maxDepth || a_maxDepth < 0)  =>  maxDepth D >= 1 gives D operators, D = 0 gives one
    if (P.preCondSolverDepth >= 0 && depth >= (P.preCondSolverDepth > 1 ? P.preCondSolverDepth : 1)) break;
    int coarsening = 1;
    Box domain = pb->grids.domain;
    for (int i = 0; i < depth; i++) { coarsening *= 2; domain = domain.coarsened(2); }   // :161-166
    if (coarsening > 1) {                                                   // :168-172
      bool ok = true;
      for (auto &b : pb->grids.boxes) if (!b.coarsenable(coarsening * s_maxCoarse)) { ok = false; break; }
      if (!ok) break;   // MGnewOp returns NULL
    }
    Op *op = new Op;
    op->dx = pb->dx0 * coarsening;                                          // :174
    op->lay.domain = domain; op->lay.periodic = pb->grids.periodic;
    for (auto &b : pb->grids.boxes) op->lay.boxes.push_back(b.coarsened(coarsening));   // coarsen_dbl :177
    op->lay.buildCopier(op->lay.exFace1, 1, true);                          // ex.coarsen :179-185
    op->alpha = P.alpha; op->beta = P.beta;                                 // :191-192
    for (int d = 0; d < 3; d++) { op->bc.lo[d] = P.bc_lo[d]; op->bc.hi[d] = P.bc_hi[d]; }
    op->bc.value = P.bc_value;
    if (depth == 0) { op->aCoef = &pb->aCoef; op->bCoef = &pb->bCoef; }     // :194-197
    else {
      op->aOwn.define(&op->lay, 1, 0); op->bOwn.define(&op->lay, 1, 0);     // :203-204
      int type = P.coefficient_average_type >= 0 ? P.coefficient_average_type : 0;   // Factory.cpp:44-46, 321
      if (type != 0 && type != 1) { fprintf(stderr, "MGNewOp -- bad averagetype\n"); abort(); }  // :222-224
      coarseAverage(op->aOwn, pb->aCoef, coarsening, type);                 // :208-220 (directly from AMR level)
      coarseAverage(op->bOwn, pb->bCoef, coarsening, type);
      op->aCoef = &op->aOwn; op->bCoef = &op->bOwn;
    }
    op->computeLambda();                                                    // :229
    pb->ops.push_back(op);
  }
  int nd = (int)pb->ops.size();
  pb->e.resize(nd); pb->r.resize(nd); pb->tmp.resize(nd);
  for (int d = 0; d < nd; d++) {
    pb->e[d].define(&pb->ops[d]->lay, 1, d == 0 ? 3 : 1);   // createCoarser(ghosted)
    pb->r[d].define(&pb->ops[d]->lay, 1, 0);
    pb->tmp[d].define(&pb->ops[d]->lay, 1, 0);
  }
  return nd;
}

int orc_mg_depths(const orc_problem *pb) { return (int)pb->ops.size(); }

void orc_level_dims(const orc_problem *pb, int depth, int n[3], double *dx) {
  if (depth == 0 || pb->ops.empty()) {   // level 0 geometry exists before the solver is defined
    for (int d = 0; d < 3; d++) n[d] = pb->grids.domain.size(d);
    *dx = pb->dx0;
    return;
  }
  const Op *op = pb->ops[depth];
  for (int d = 0; d < 3; d++) n[d] = op->lay.domain.size(d);
  *dx = op->dx;
}

void orc_get_field(orc_problem *pb, int depth, int field, double *out) {
  const LevelData *ld; int comp = 0;
  if (field >= ORC_F_MGVAR0) { ld = &pb->mgvars; comp = field - ORC_F_MGVAR0; }
  else ld = fieldOf(pb, depth, field);
  const Box &dom = ld->lay->domain; long nx = dom.size(0), ny = dom.size(1);
#pragma omp parallel for schedule(static)
  for (int n = 0; n < ld->size(); n++) {
    const Box &b = ld->lay->boxes[n]; const FAB &f = ld->fab[n];
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++) out[i + nx * (j + ny * (long)k)] = f(i, j, k, comp);
  }
}

void orc_get_field_ghosted(orc_problem *pb, int depth, int field, int ng, double *out) {
  const LevelData *ld; int comp = 0;
  if (field >= ORC_F_MGVAR0) { ld = &pb->mgvars; comp = field - ORC_F_MGVAR0; }
  else ld = fieldOf(pb, depth, field);
  const Box &dom = ld->lay->domain; long nx = dom.size(0) + 2 * ng, ny = dom.size(1) + 2 * ng;
  // valid cells first, then ghost cells outside the domain (from the box that owns the adjacent cells)
  for (int n = 0; n < ld->size(); n++) {
    const FAB &f = ld->fab[n]; Box g = ld->lay->boxes[n].grown(std::min(ng, ld->ng));
    const Box &v = ld->lay->boxes[n];
    for (int k = g.lo[2]; k <= g.hi[2]; k++)
      for (int j = g.lo[1]; j <= g.hi[1]; j++)
        for (int i = g.lo[0]; i <= g.hi[0]; i++) {
          // a cell outside the domain is reported by the box that owns the nearest domain cell
          int ci = std::min(std::max(i, dom.lo[0]), dom.hi[0]), cj = std::min(std::max(j, dom.lo[1]), dom.hi[1]),
              ck = std::min(std::max(k, dom.lo[2]), dom.hi[2]);
          bool owner = (ci >= v.lo[0] && ci <= v.hi[0] && cj >= v.lo[1] && cj <= v.hi[1] && ck >= v.lo[2] && ck <= v.hi[2]);
          if (owner) out[(i + ng) + nx * ((j + ng) + ny * (long)(k + ng))] = f(i, j, k, comp);
        }
  }
}

void orc_set_field(orc_problem *pb, int depth, int field, const double *in) {
  LevelData *ld; int comp = 0;
  if (field >= ORC_F_MGVAR0) { ld = &pb->mgvars; comp = field - ORC_F_MGVAR0; }
  else ld = fieldOf(pb, depth, field);
  const Box &dom = ld->lay->domain; long nx = dom.size(0), ny = dom.size(1);
#pragma omp parallel for schedule(static)
  for (int n = 0; n < ld->size(); n++) {
    const Box &b = ld->lay->boxes[n]; FAB &f = ld->fab[n];
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++) f(i, j, k, comp) = in[i + nx * (j + ny * (long)k)];
  }
  if ((field == ORC_F_A || field == ORC_F_B) && depth < (int)pb->ops.size()) pb->ops[depth]->lambdaNeedsResetting = true;
}

void orc_op_relax(orc_problem *pb, int d, int iterations) { pb->ops[d]->relax(pb->e[d], pb->r[d], iterations); }
void orc_op_gsrb_color(orc_problem *pb, int d, int whichPass) {
  pb->ops[d]->resetLambda(); pb->ops[d]->gsrbColor(pb->e[d], pb->r[d], whichPass);
}
void orc_op_residual(orc_problem *pb, int d, int homog) { pb->ops[d]->residual(pb->tmp[d], pb->e[d], pb->r[d], homog != 0); }
void orc_op_apply(orc_problem *pb, int d, int homog) { pb->ops[d]->applyOp(pb->tmp[d], pb->e[d], homog != 0); }
void orc_op_restrict(orc_problem *pb, int d) { pb->ops[d]->restrictResidual(pb->r[d + 1], pb->e[d], pb->r[d]); }
void orc_op_prolong(orc_problem *pb, int d) { pb->ops[d]->prolongIncrement(pb->e[d], pb->e[d + 1]); }
void orc_op_precond(orc_problem *pb, int d) { pb->ops[d]->preCond(pb->e[d], pb->r[d]); }
double orc_op_norm(orc_problem *pb, int d, int field, int ord) { return pb->ops[d]->norm(*fieldOf(pb, d, field), ord); }
double orc_op_dot(orc_problem *pb, int d, int f1, int f2) { return pb->ops[d]->dot(*fieldOf(pb, d, f1), *fieldOf(pb, d, f2)); }

// ---- one AMR level > 0: a box of the refined domain (split into max_grid_size boxes) with its coarser level's spacing.
// What the reference's operator class itself does on such a level: levelGSRB and restrictResidual with
// homogeneousCFInterp (VariableCoeffPoissonOperator.cpp:156,296).  Fields are patch-shaped arrays, x fastest.
struct orc_patch {
  Op op;
  LevelData e, r, a, b, rc, lof;   // correction (1 ghost), residual, coefficients, MG-coarsened residual, operator output
  Layout clay;
  Box box;
  FAB crse; Box cdom;              // the coarser AMR level's field on its whole domain (QuadCFInterp input)
  // the level's problem data (Main_PoissonSolver.cpp:79-88 on level ilev > 0): multigrid_vars, dpsi, rhs
  orc_params P; bool hasP = false;
  LevelData mgvars, dpsi, rhs;
};

orc_patch *orc_patch_create(const int n_domain[3], const int lo[3], const int hi[3], int max_grid_size, double dx, double dx_crse,
                            double alpha, double beta, const int bc_lo[3], const int bc_hi[3], double bc_value) {
  const Box pbox(lo, hi);
  const Box dom(0, 0, 0, n_domain[0] - 1, n_domain[1] - 1, n_domain[2] - 1);
  // (AMR boxes are aligned to block_factor; restrictResidual needs every box coarsenable by 2)
  if (pbox.empty() || !dom.contains(pbox) || !pbox.coarsenable(2) || max_grid_size % 2) return nullptr;
  orc_patch *pp = new orc_patch;
  pp->box = pbox;
  Op &op = pp->op;
  op.lay.domain = Box(0, 0, 0, n_domain[0] - 1, n_domain[1] - 1, n_domain[2] - 1);
  domainSplit(pp->box, max_grid_size, op.lay.boxes);
  op.lay.buildCopier(op.lay.exFace1, 1, true);
  op.dx = dx; op.alpha = alpha; op.beta = beta;
  op.hasCoarser = true; op.dxCrse = dx_crse;
  for (int d = 0; d < 3; d++) { op.bc.lo[d] = bc_lo[d]; op.bc.hi[d] = bc_hi[d]; }
  op.bc.value = bc_value;
  pp->e.define(&op.lay, 1, 1); pp->r.define(&op.lay, 1, 0); pp->a.define(&op.lay, 1, 0); pp->b.define(&op.lay, 1, 0);
  op.aCoef = &pp->a; op.bCoef = &pp->b;
  op.lambda.define(&op.lay, 1, 0);
  pp->clay.domain = op.lay.domain.coarsened(2);
  for (auto &bx : op.lay.boxes) pp->clay.boxes.push_back(bx.coarsened(2));
  pp->rc.define(&pp->clay, 1, 0);
  pp->lof.define(&op.lay, 1, 0);
  pp->cdom = op.lay.domain.coarsened(2);
  pp->crse.define(pp->cdom, 1);
  return pp;
}
// The same level given as a LIST of boxes (BRMeshRefine's output: boxes that touch, union not a rectangle), cut further
// into max_grid_size boxes.  Fields are exchanged with the caller as arrays over the union's bounding box.
orc_patch *orc_patch_create_boxes(const int n_domain[3], int nboxes, const int *boxes, int max_grid_size, double dx, double dx_crse,
                                  double alpha, double beta, const int bc_lo[3], const int bc_hi[3], double bc_value) {
  const Box dom(0, 0, 0, n_domain[0] - 1, n_domain[1] - 1, n_domain[2] - 1);
  if (nboxes < 1 || max_grid_size % 2) return nullptr;
  orc_patch *pp = new orc_patch;
  Op &op = pp->op;
  op.lay.domain = dom;
  int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-1, -1, -1};
  for (int q = 0; q < nboxes; q++) {
    const Box b(boxes + 6 * q, boxes + 6 * q + 3);
    if (b.empty() || !dom.contains(b) || !b.coarsenable(2)) { delete pp; return nullptr; }
    for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], b.lo[d]); hi[d] = std::max(hi[d], b.hi[d]); }
    std::vector<Box> cut;
    domainSplit(b, max_grid_size, cut);
    for (auto &c : cut) op.lay.boxes.push_back(c);
  }
  pp->box = Box(lo, hi);
  op.lay.buildCopier(op.lay.exFace1, 1, true);
  op.dx = dx; op.alpha = alpha; op.beta = beta;
  op.hasCoarser = true; op.dxCrse = dx_crse;
  for (int d = 0; d < 3; d++) { op.bc.lo[d] = bc_lo[d]; op.bc.hi[d] = bc_hi[d]; }
  op.bc.value = bc_value;
  pp->e.define(&op.lay, 1, 1); pp->r.define(&op.lay, 1, 0); pp->a.define(&op.lay, 1, 0); pp->b.define(&op.lay, 1, 0);
  op.aCoef = &pp->a; op.bCoef = &pp->b;
  op.lambda.define(&op.lay, 1, 0);
  pp->clay.domain = op.lay.domain.coarsened(2);
  for (auto &bx : op.lay.boxes) pp->clay.boxes.push_back(bx.coarsened(2));
  pp->rc.define(&pp->clay, 1, 0);
  pp->lof.define(&op.lay, 1, 0);
  pp->cdom = op.lay.domain.coarsened(2);
  pp->crse.define(pp->cdom, 1);
  return pp;
}
void orc_patch_destroy(orc_patch *pp) { delete pp; }

static LevelData *patchField(orc_patch *pp, int field) {
  switch (field) {
    case ORC_F_E: return &pp->e;
    case ORC_F_R: return &pp->r;
    case ORC_F_A: return &pp->a;
    case ORC_F_B: return &pp->b;
    case ORC_F_TMP: return &pp->rc;
    case ORC_F_LAMBDA: return &pp->op.lambda;
    case ORC_F_RHS: return &pp->lof;
  }
  fprintf(stderr, "orc_patch: bad field %d\n", field); abort();
}
// patch-shaped array (coarsened patch for ORC_F_TMP = the restricted residual)
void orc_patch_set(orc_patch *pp, int field, const double *in) {
  LevelData *ld = patchField(pp, field);
  const Box pb = (field == ORC_F_TMP) ? pp->box.coarsened(2) : pp->box;
  const long nx = pb.size(0), ny = pb.size(1);
  for (int n = 0; n < ld->size(); n++) {
    const Box &b = ld->lay->boxes[n]; FAB &f = ld->fab[n];
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++) f(i, j, k) = in[(i - pb.lo[0]) + nx * ((j - pb.lo[1]) + ny * (long)(k - pb.lo[2]))];
  }
  if (field == ORC_F_A || field == ORC_F_B) pp->op.lambdaNeedsResetting = true;
}
void orc_patch_get(orc_patch *pp, int field, double *out) {
  if (field == ORC_F_LAMBDA) pp->op.resetLambda();
  const LevelData *ld = patchField(pp, field);
  const Box pb = (field == ORC_F_TMP) ? pp->box.coarsened(2) : pp->box;
  const long nx = pb.size(0), ny = pb.size(1);
  for (int n = 0; n < ld->size(); n++) {
    const Box &b = ld->lay->boxes[n]; const FAB &f = ld->fab[n];
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++) out[(i - pb.lo[0]) + nx * ((j - pb.lo[1]) + ny * (long)(k - pb.lo[2]))] = f(i, j, k);
  }
}
int orc_patch_num_boxes(const orc_patch *pp) { return (int)pp->op.lay.boxes.size(); }

// ---- the nonlinear loop's per-level steps on an AMR level > 0 (Main_PoissonSolver.cpp:93, 154-160, 189-205) ----
// set_initial_conditions on the level's ghosted boxes (psi = 1, dpsi = 0, phi and A_ij analytic, ghost cells included)
void orc_patch_set_initial_conditions(orc_patch *pp, const orc_params *P) {
  pp->P = *P; pp->hasP = true;
  pp->mgvars.define(&pp->op.lay, 8, 1);    // the reference allocates three ghost layers; its stencils read one
  pp->dpsi.define(&pp->op.lay, 1, 1);
  pp->rhs.define(&pp->op.lay, 1, 0);
  init_conditions_level(pp->P, pp->op.dx, pp->mgvars, pp->dpsi);
}
// set_a_coef / set_b_coef / set_rhs: into the patch's coefficient fields (ORC_F_A, ORC_F_B) and its rhs (ORC_F_DPSI + 1 = ORC_F_RHS slot below)
void orc_patch_set_coefs_and_rhs(orc_patch *pp, double constant_K) {
  coefs_and_rhs_level(pp->P, pp->op.dx, constant_K, pp->mgvars, pp->rhs, pp->a, pp->b);
  pp->op.lambdaNeedsResetting = true;
}
// Main_PoissonSolver.cpp:189-205: QuadCFInterp of dpsi from the coarser level's dpsi, exchange, psi += dpsi over the GHOSTED
// boxes.  dpsi arrives as a bounding-box array (valid cells); its physical-boundary ghost is what [Chombo] BiCGStab leaves
// there, a*near + b (inhomogeneous ParseBC fill, see orc_update_psi0's single-level analysis); coarse = the coarser level's
// dpsi on its whole domain (orc_patch_set_coarse).
void orc_patch_update_psi(orc_patch *pp, const double *dpsi_in) {
  const Box pb = pp->box;
  const long nx = pb.size(0), ny = pb.size(1);
  for (int n = 0; n < pp->dpsi.size(); n++) {
    const Box &b = pp->op.lay.boxes[n]; FAB &f = pp->dpsi.fab[n];
    f.setVal(0.0);
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++) f(i, j, k) = dpsi_in[(i - pb.lo[0]) + nx * ((j - pb.lo[1]) + ny * (long)(k - pb.lo[2]))];
  }
  pp->op.quadCFInterp(pp->dpsi, pp->crse, pp->cdom);    // :193-195
  pp->op.applyBC(pp->dpsi, false);
  exchange(pp->dpsi, pp->op.lay.exFace1);               // :200-201 + SetLevelData.cpp:249 (face ghosts are the ones read)
  for (int n = 0; n < pp->mgvars.size(); n++) {
    FAB &mv = pp->mgvars.fab[n]; const FAB &dp = pp->dpsi.fab[n];
    const Box &b = mv.b;                                // ghosted :256
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++) mv(i, j, k, 0) += dp(i, j, k, 0);   // :260
  }
}
// component `comp` of multigrid_vars (comp 0..7) or the rhs (comp 8) over the bounding box (valid cells)
void orc_patch_get_var(orc_patch *pp, int comp, double *out) {
  const Box pb = pp->box;
  const long nx = pb.size(0), ny = pb.size(1);
  const LevelData &ld = comp == 8 ? pp->rhs : pp->mgvars;
  for (int n = 0; n < ld.size(); n++) {
    const Box &b = pp->op.lay.boxes[n]; const FAB &f = ld.fab[n];
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++)
          out[(i - pb.lo[0]) + nx * ((j - pb.lo[1]) + ny * (long)(k - pb.lo[2]))] = f(i, j, k, comp == 8 ? 0 : comp);
  }
}
void orc_patch_relax(orc_patch *pp, int iterations) { pp->op.relax(pp->e, pp->r, iterations); }
void orc_patch_gsrb_color(orc_patch *pp, int whichPass) { pp->op.resetLambda(); pp->op.gsrbColor(pp->e, pp->r, whichPass); }
void orc_patch_restrict(orc_patch *pp) { pp->op.restrictResidual(pp->rc, pp->e, pp->r); }
void orc_patch_precond(orc_patch *pp) { pp->op.preCond(pp->e, pp->r); }
// the coarser AMR level's field: an array over its whole domain (n_domain / 2), x fastest
void orc_patch_set_coarse(orc_patch *pp, const double *in) { std::copy(in, in + pp->crse.d.size(), pp->crse.d.begin()); }
// AMROperatorNF / AMRResidualNF: lof (read back as ORC_F_RHS) = L(e) resp. r - L(e), coarse-fine ghosts by QuadCFInterp
void orc_patch_amr_operator_nf(orc_patch *pp, int homog) { pp->op.amrOperatorNF(pp->lof, pp->e, pp->crse, pp->cdom, homog != 0); }
void orc_patch_amr_residual_nf(orc_patch *pp, int homog) { pp->op.amrResidualNF(pp->lof, pp->e, pp->crse, pp->cdom, pp->r, homog != 0); }
// the ghost value homogeneousCFInterp produces from the two interior cells (far, near) -- the coefficients' pin
double orc_interp_homo(double dx, double dx_crse, double far_value, double near_value) {
  Op op; op.dx = dx; op.dxCrse = dx_crse; op.hasCoarser = true;
  op.lay.domain = Box(0, 0, 0, 7, 0, 0);
  op.lay.boxes.push_back(Box(0, 0, 0, 3, 0, 0));
  LevelData x; x.define(&op.lay, 1, 1);
  x.fab[0].setVal(0.0);
  x.fab[0](2, 0, 0) = far_value; x.fab[0](3, 0, 0) = near_value;
  op.homogeneousCFInterp(x);
  return x.fab[0](4, 0, 0);
}

int orc_vcycle(orc_problem *pb) {
  pb->lastBottomIters = 0;
  mg_cycle(pb, 0, pb->e[0], pb->r[0]);
  return pb->lastBottomIters;
}

int orc_bottom_solve(orc_problem *pb) { return bottomSolve(pb, pb->e.back(), pb->r.back()); }

void orc_load_rhs_zero_e(orc_problem *pb) {
  pb->ops[0]->assign(pb->r[0], pb->rhs);
  pb->ops[0]->setToZero(pb->e[0]);
}

// solver.solve(dpsi, rhs) -- Main_PoissonSolver.cpp:173-184
int orc_outer_solve(orc_problem *pb, int *exit_status, double *final_norm, double *norms, int max_norms) {
  OuterLinOp lop{pb};
  BiCGParams bp;
  bp.homogeneous = false;            // :172-173
  bp.normType = 0;                   // :176
  bp.eps = pb->P.tolerance;          // :177
  bp.imax = pb->P.max_iterations;    // :178
  bp.verbosity = pb->P.verbosity >= 5 ? 4 : 0;
  std::vector<Real> hist; int status = 0;
  int it = bicgstab<LevelData, OuterLinOp>(lop, pb->dpsi, pb->rhs, bp, &status, &hist);
  if (exit_status) *exit_status = status;
  if (final_norm) *final_norm = hist.empty() ? 0.0 : hist.back();
  for (int q = 0; q < max_norms && q < (int)hist.size(); q++) norms[q] = hist[q];
  return it;
}

// Main_PoissonSolver.cpp:189-205 (set_update_psi0, SetLevelData.cpp:243-263) + computeNorm :208
double orc_update_psi0(orc_problem *pb) {
  exchange(pb->dpsi, pb->grids.exFull3);                                  // SetLevelData.cpp:249
#pragma omp parallel for schedule(static)
  for (int n = 0; n < pb->mgvars.size(); n++) {
    FAB &mv = pb->mgvars.fab[n]; const FAB &dp = pb->dpsi.fab[n];
    const Box &b = mv.b;                                                  // ghosted :256
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++) mv(i, j, k, 0) += dp(i, j, k, 0);   // :260
  }
  // [Chombo] computeNorm(Vector<LD*>, refRatio, dxCrse, interval, p = 2): (sum |x|^2 dx^3)^(1/2), single level
  std::vector<Real> part(pb->dpsi.size(), 0.0);
#pragma omp parallel for schedule(static)
  for (int n = 0; n < pb->dpsi.size(); n++) {
    const Box &g = pb->grids.boxes[n]; Real s = 0.0;
    for (int k = g.lo[2]; k <= g.hi[2]; k++)
      for (int j = g.lo[1]; j <= g.hi[1]; j++)
        for (int i = g.lo[0]; i <= g.hi[0]; i++) { Real v = pb->dpsi.fab[n](i, j, k); s += v * v; }
    part[n] = s;
  }
  Real sum = 0.0; for (Real s : part) sum += s;
  Real dV = pb->dx0 * pb->dx0 * pb->dx0;
  return std::sqrt(sum * dV);
}

// set_regrid_condition (Source/SetLevelData.cpp:188-240; mode 0) / set_constant_K_integrand (:128-186; mode 1) on freshly
// initialised data (set_grids: Source/SetGrids.cpp:86-95) over the index box [lo, hi] of a level with spacing dx
void orc_condition_box(const orc_params *Pp, double dx, const int lo[3], const int hi[3], int mode, double *out) {
  const orc_params &P = *Pp;
  Layout lay;
  lay.domain = Box(lo, hi).grown(4);
  lay.boxes.push_back(Box(lo, hi));
  LevelData mgvars, dpsi;
  mgvars.define(&lay, 8, 1); dpsi.define(&lay, 1, 1);
  init_conditions_level(P, dx, mgvars, dpsi);
  const FAB &mv = mgvars.fab[0];
  const Box tb(lo, hi);
  FAB rho, lap; rho.define(tb, 1); lap.define(tb, 1);
  View psiV = fabview(mv); View phiV = fabview(mv); phiV.p += mv.sc * 7;
  k_rho(fabview(rho), phiV, dx, tb);                                        // :205-208 / :151-154
  k_lap(fabview(lap), psiV, dx, tb);                                        // :145-148 (mode 1)
  const long nx = tb.size(0), ny = tb.size(1);
  for (int k = tb.lo[2]; k <= tb.hi[2]; k++)
    for (int j = tb.lo[1]; j <= tb.hi[1]; j++)
      for (int i = tb.lo[0]; i <= tb.hi[0]; i++) {
        Real loc[3]; cellLoc(P, dx, i, j, k, loc);
        Real m = m_value(P, 0.0);                                           // set_m_value(m, phi, params, 0.0)
        Real A2 = A2_of(mv, i, j, k);
        Real psi_bh = set_binary_bh_psi(loc, P);
        Real psi_0 = mv(i, j, k, 0) + psi_bh;
        Real v;
        if (mode == 0)
          v = 1.5 * std::abs(m) + 1.5 * A2 * std::pow(psi_0, -7.0) + 24.0 * M_PI * P.G_Newton * std::abs(rho(i, j, k)) * std::pow(psi_0, 1.0) +
              std::log(psi_0);                                              // :233-236
        else
          v = -1.5 * m + 1.5 * A2 * std::pow(psi_0, -12.0) + 24.0 * M_PI * P.G_Newton * rho(i, j, k) * std::pow(psi_0, -4.0) +
              12.0 * lap(i, j, k) * std::pow(psi_0, -5.0);                  // :180-183
        out[(i - tb.lo[0]) + nx * ((j - tb.lo[1]) + ny * (long)(k - tb.lo[2]))] = v;
      }
}

// set_output_data (Source/SetLevelData.cpp:343-396): the 32 GRChombo variables (GRChomboUserVariables.hpp) over the index
// box [lo, hi] (the checkpoint's boxes carry three ghost layers, WriteOutput.H:180) of a level with spacing dx, from the
// given psi (multigrid_vars component 0 over the same box); phi and A_ij are the analytic initial data, which nothing
// modifies after set_initial_conditions.  out = 32 x box, component slowest.
void orc_output_box(const orc_params *Pp, double dx, const int lo[3], const int hi[3], const double *psi, double constant_K, double *out) {
  const orc_params &P = *Pp;
  enum { c_chi = 0, c_h11 = 1, c_h22 = 4, c_h33 = 6, c_K = 7, c_A11 = 8, c_A12 = 9, c_A13 = 10, c_A22 = 11, c_A23 = 12, c_A33 = 13,
         c_lapse = 18, c_phi = 25, NV = 32 };
  const Box tb(lo, hi);
  const long nx = tb.size(0), ny = tb.size(1), np = tb.numPts();
  Layout lay; lay.domain = tb; lay.boxes.push_back(tb);
  LevelData mgvars, dpsi;
  mgvars.define(&lay, 8, 0); dpsi.define(&lay, 1, 0);
  init_conditions_level(P, dx, mgvars, dpsi);
  const FAB &mv = mgvars.fab[0];
  for (long q = 0; q < NV * np; q++) out[q] = 0.0;                           // :357-359
  for (long q = 0; q < np; q++) {
    out[c_h11 * np + q] = 1.0; out[c_h22 * np + q] = 1.0; out[c_h33 * np + q] = 1.0; out[c_lapse * np + q] = 1.0;   // :363-366
    out[c_K * np + q] = constant_K;                                          // :369
  }
  for (int k = tb.lo[2]; k <= tb.hi[2]; k++)
    for (int j = tb.lo[1]; j <= tb.hi[1]; j++)
      for (int i = tb.lo[0]; i <= tb.hi[0]; i++) {
        const long q = (i - tb.lo[0]) + nx * ((j - tb.lo[1]) + ny * (long)(k - tb.lo[2]));
        Real loc[3]; cellLoc(P, dx, i, j, k, loc);                           // :376-378
        Real psi_bh = set_binary_bh_psi(loc, P);
        Real chi = std::pow(psi[q] + psi_bh, -4.0);                          // :381-383
        out[c_chi * np + q] = chi;
        Real factor = std::pow(chi, 1.5);                                    // :384
        out[c_phi * np + q] = mv(i, j, k, 7);                                // :387
        out[c_A11 * np + q] = mv(i, j, k, 1) * factor;                       // :388-393
        out[c_A12 * np + q] = mv(i, j, k, 2) * factor;
        out[c_A13 * np + q] = mv(i, j, k, 3) * factor;
        out[c_A22 * np + q] = mv(i, j, k, 4) * factor;
        out[c_A23 * np + q] = mv(i, j, k, 5) * factor;
        out[c_A33 * np + q] = mv(i, j, k, 6) * factor;
      }
}

// dpsi := the given valid cells, ghost layer 1 = the inhomogeneous ParseBC fill (what [Chombo] BiCGStab leaves there: the
// initial residual's inhomogeneous fill + the homogeneous ghosts of the accumulated correction).  For callers whose solver
// ran outside this library (the hierarchy twin of tests/amr_twin.py) before orc_update_psi0.
void orc_set_dpsi_with_bc(orc_problem *pb, const double *in) {
  const Box &dom = pb->grids.domain;
  const long nx = dom.size(0), ny = dom.size(1);
  for (int n = 0; n < pb->dpsi.size(); n++) {
    const Box &b = pb->grids.boxes[n]; FAB &f = pb->dpsi.fab[n];
    f.setVal(0.0);
    for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
        for (int i = b.lo[0]; i <= b.hi[0]; i++) f(i, j, k) = in[i + nx * (j + ny * (long)k)];
  }
  pb->ops[0]->applyBC(pb->dpsi, false);
}

// NL loop -- Main_PoissonSolver.cpp:131-216 (non-periodic: constant_K = 0)
int orc_nl_solve(orc_problem *pb, double *dpsi_norms, int max_out) {
  int it = 0;
  for (int NL_iter = 0; NL_iter < pb->P.max_NL_iterations; NL_iter++) {
    orc_set_coefs_and_rhs(pb, 0.0);                                        // :154-160
    orc_define_solver(pb);                                                 // :163-170
    int status; double fn;
    orc_outer_solve(pb, &status, &fn, nullptr, 0);                         // :184
    double dpsi_norm = orc_update_psi0(pb);                                // :189-208
    if (NL_iter < max_out) dpsi_norms[NL_iter] = dpsi_norm;
    it = NL_iter + 1;
    if (dpsi_norm < pb->P.tolerance || dpsi_norm > 1e5) break;             // :212
  }
  return it;
}

}  // extern "C"
