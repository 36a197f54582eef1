// grids.cu -- set_grids (Source/SetGrids.cpp:31-207): the AMR hierarchy the reference builds before it solves.
//
//   * base level: domainSplit lattice of max_grid_size boxes (:54-58);
//   * repeat while a new level appears (:70-136): evaluate set_regrid_condition on freshly initialised data on every
//     existing level (:79-103), tag the cells with |condition| >= refine_threshold * max|condition| of their level, grow
//     the tags by 2 cells and clip them to the domain (set_tag_cells, :172-207), hand them to BRMeshRefine::regrid with
//     nesting radius 2 (:64-68, :113-114).
//
// The regrid condition is a device kernel (source.cu k_condition: with psi = 1 it is a function of position alone, so it
// is evaluated analytically over the bounding box of each connected part of a level).  Everything else here is index
// bookkeeping on the host, as it is in the reference.
//
// [Chombo 3.2] BRMeshRefine / MeshRefine::regrid is NOT vendored with the reference (GNUmakefile:12,27) -- restated from the
// published algorithm (Berger & Rigoutsos 1991, as Chombo's design document describes its use):
//   top-down over the levels; tags of level l (+ the coarsened new boxes of level l+2 grown by the nesting radius, so the
//   new level l+1 holds them) are clipped to the proper nesting domain of level l (cells at least `nesting_radius` cells
//   away from any cell of the domain that is not in level l), coarsened by block_factor / ref_ratio, clustered into boxes
//   by signatures: a box is accepted when its tagged fraction >= fill_ratio and it lies in the nesting domain, else it is
//   cut at a hole of a signature (the one nearest the middle), else at the strongest inflection of a signature's second
//   difference, else bisected along its longest edge; accepted boxes are broken up to max_grid_size and refined to level
//   l+1's index space (so they are multiples of block_factor).
// Which cells end up refined depends on those choices: UNPINNED upstream.  What the solver computes depends
// only on the UNION of a level's boxes, not on how it is cut (the library holds each connected part in one masked array).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <numeric>

#include "mgic_internal.h"

namespace {

struct IBox {
  int lo[3], hi[3];
  long long vol() const { return (long long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1); }
};
struct Pt { int x[3]; };

// dense bit set over an index box; reads outside the box give `outside`
struct Bits {
  IBox b;
  int n[3] = {0, 0, 0};
  std::vector<unsigned char> v;
  void define(const IBox &box) {
    b = box;
    for (int d = 0; d < 3; d++) n[d] = std::max(0, box.hi[d] - box.lo[d] + 1);
    v.assign((size_t)n[0] * n[1] * n[2], 0);
  }
  bool inside(int i, int j, int k) const {
    return i >= b.lo[0] && i <= b.hi[0] && j >= b.lo[1] && j <= b.hi[1] && k >= b.lo[2] && k <= b.hi[2];
  }
  size_t at(int i, int j, int k) const { return (size_t)(i - b.lo[0]) + (size_t)n[0] * ((size_t)(j - b.lo[1]) + (size_t)n[1] * (size_t)(k - b.lo[2])); }
  bool get(int i, int j, int k) const { return inside(i, j, k) && v[at(i, j, k)]; }
  void set(int i, int j, int k) { if (inside(i, j, k)) v[at(i, j, k)] = 1; }
  void fill(const IBox &q) {
    for (int k = std::max(q.lo[2], b.lo[2]); k <= std::min(q.hi[2], b.hi[2]); k++)
      for (int j = std::max(q.lo[1], b.lo[1]); j <= std::min(q.hi[1], b.hi[1]); j++)
        for (int i = std::max(q.lo[0], b.lo[0]); i <= std::min(q.hi[0], b.hi[0]); i++) v[at(i, j, k)] = 1;
  }
};

inline int floordiv(int a, int r) { return a >= 0 ? a / r : -((-a + r - 1) / r); }

// One axis of a box filter over a Bits array, in place, O(cells): with grow = true a cell becomes set when a set cell lies
// within `radius` along the axis (dilation; nothing outside the array is set); with grow = false a cell stays set only
// if every position within `radius` along the axis is set, where positions outside the array count as set when they are
// outside the domain [0, ndom) and as clear otherwise (erosion with "beyond the domain boundary is inside").  A cube filter
// is the three axes one after the other.
void filter_axis(Bits &u, int axis, int radius, bool grow, const int *ndom) {
  const size_t st[3] = {1, (size_t)u.n[0], (size_t)u.n[0] * u.n[1]};
  const int len = u.n[axis], a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
  if (len == 0 || radius <= 0) return;
  std::vector<int> dist(len);
  const int FAR = 1 << 29;
  for (int q2 = 0; q2 < u.n[a2]; q2++)
    for (int q1 = 0; q1 < u.n[a1]; q1++) {
      unsigned char *row = u.v.data() + q1 * st[a1] + q2 * st[a2];
      // distance along the row to the nearest "hit": a set cell (grow) / a clear position (shrink)
      int last = -FAR;
      if (!grow && u.b.lo[axis] - 1 >= 0) last = -1;                         // the position just before the array is a clear one
      for (int x = 0; x < len; x++) {
        const bool hit = grow ? row[x * st[axis]] != 0 : row[x * st[axis]] == 0;
        if (hit) last = x;
        dist[x] = x - last;
      }
      last = FAR;
      if (!grow && u.b.hi[axis] + 1 <= ndom[axis] - 1) last = len;            // ... and the one just after it
      for (int x = len - 1; x >= 0; x--) {
        const bool hit = grow ? row[x * st[axis]] != 0 : row[x * st[axis]] == 0;
        if (hit) last = x;
        dist[x] = std::min(dist[x], last - x);
      }
      for (int x = 0; x < len; x++) {
        if (grow) { if (dist[x] <= radius) row[x * st[axis]] = 1; }
        else if (dist[x] <= radius) row[x * st[axis]] = 0;
      }
    }
}

IBox bounding(const std::vector<IBox> &boxes) {
  IBox bb = {{1 << 30, 1 << 30, 1 << 30}, {-(1 << 30), -(1 << 30), -(1 << 30)}};
  for (const IBox &q : boxes)
    for (int d = 0; d < 3; d++) { bb.lo[d] = std::min(bb.lo[d], q.lo[d]); bb.hi[d] = std::max(bb.hi[d], q.hi[d]); }
  return bb;
}

// [Chombo] domainSplit: lattice of boxes of at most maxSize cells per edge
void domain_split(const int n[3], int maxSize, std::vector<IBox> &out) {
  for (int k = 0; k < n[2]; k += maxSize)
    for (int j = 0; j < n[1]; j += maxSize)
      for (int i = 0; i < n[0]; i += maxSize)
        out.push_back({{i, j, k}, {std::min(i + maxSize, n[0]) - 1, std::min(j + maxSize, n[1]) - 1, std::min(k + maxSize, n[2]) - 1}});
}

// proper nesting domain of a level: its cells whose whole (2r+1)^3 neighbourhood lies in the level or outside the domain
Bits nesting_domain(const std::vector<IBox> &boxes, const int ndom[3], int radius) {
  Bits u;
  if (boxes.empty()) { u.define({{0, 0, 0}, {-1, -1, -1}}); return u; }
  u.define(bounding(boxes));
  for (const IBox &q : boxes) u.fill(q);
  // erosion by the (2 radius + 1)^3 cube, beyond the domain boundary counting as inside (= `radius` erosions by 3^3)
  for (int axis = 0; axis < 3; axis++) filter_axis(u, axis, radius, false, ndom);
  return u;
}

// ---- Berger-Rigoutsos clustering of tagged cells (in whatever index space the points live) --------------------------
struct Cluster {
  const Bits *ok;       // cells a box may contain (the nesting domain in this index space)
  double fill;
  int maxSize;
  std::vector<IBox> out;

  // every cell of q is an allowed one: a box query on the summed-area table of `ok`
  std::vector<int> sat;
  bool nested(const IBox &q) {
    const int n0 = ok->n[0], n1 = ok->n[1], n2 = ok->n[2];
    for (int d = 0; d < 3; d++)
      if (q.lo[d] < ok->b.lo[d] || q.hi[d] > ok->b.hi[d]) return false;   // cells outside the array are not allowed ones
    const size_t s1 = (size_t)n0 + 1, s2 = s1 * ((size_t)n1 + 1);
    if (sat.empty()) {
      sat.assign(s2 * ((size_t)n2 + 1), 0);
      for (int k = 0; k < n2; k++)
        for (int j = 0; j < n1; j++) {
          int row = 0;
          for (int i = 0; i < n0; i++) {
            row += ok->v[(size_t)i + (size_t)n0 * ((size_t)j + (size_t)n1 * k)] ? 1 : 0;
            sat[(i + 1) + s1 * (j + 1) + s2 * (k + 1)] = row + sat[(i + 1) + s1 * j + s2 * (k + 1)] + sat[(i + 1) + s1 * (j + 1) + s2 * k] -
                                                         sat[(i + 1) + s1 * j + s2 * k];
          }
        }
    }
    const int a0 = q.lo[0] - ok->b.lo[0], a1 = q.lo[1] - ok->b.lo[1], a2 = q.lo[2] - ok->b.lo[2];
    const int b0 = q.hi[0] - ok->b.lo[0] + 1, b1 = q.hi[1] - ok->b.lo[1] + 1, b2 = q.hi[2] - ok->b.lo[2] + 1;
    auto S = [&](int i, int j, int k) { return (long long)sat[i + s1 * j + s2 * k]; };
    const long long cnt = S(b0, b1, b2) - S(a0, b1, b2) - S(b0, a1, b2) - S(b0, b1, a2) + S(a0, a1, b2) + S(a0, b1, a2) + S(b0, a1, a2) - S(a0, a1, a2);
    return cnt == q.vol();
  }
  void accept(const IBox &q) {   // break up to maxSize, pieces as equal as possible
    int np[3];
    for (int d = 0; d < 3; d++) np[d] = (q.hi[d] - q.lo[d] + 1 + maxSize - 1) / maxSize;
    for (int c = 0; c < np[2]; c++)
      for (int b = 0; b < np[1]; b++)
        for (int a = 0; a < np[0]; a++) {
          const int idx[3] = {a, b, c};
          IBox p;
          for (int d = 0; d < 3; d++) {
            const int len = q.hi[d] - q.lo[d] + 1, base = len / np[d], rem = len % np[d];
            p.lo[d] = q.lo[d] + idx[d] * base + std::min(idx[d], rem);
            p.hi[d] = p.lo[d] + base + (idx[d] < rem ? 1 : 0) - 1;
          }
          out.push_back(p);
        }
  }
  void run(std::vector<Pt> &pts) {
    if (pts.empty()) return;
    IBox bb = {{pts[0].x[0], pts[0].x[1], pts[0].x[2]}, {pts[0].x[0], pts[0].x[1], pts[0].x[2]}};
    for (const Pt &p : pts)
      for (int d = 0; d < 3; d++) { bb.lo[d] = std::min(bb.lo[d], p.x[d]); bb.hi[d] = std::max(bb.hi[d], p.x[d]); }
    const bool isNested = nested(bb);
    if ((double)pts.size() >= fill * (double)bb.vol() && isNested) { accept(bb); return; }
    // signatures
    std::vector<long long> sig[3];
    for (int d = 0; d < 3; d++) sig[d].assign(bb.hi[d] - bb.lo[d] + 1, 0);
    for (const Pt &p : pts)
      for (int d = 0; d < 3; d++) sig[d][p.x[d] - bb.lo[d]]++;
    int cutDir = -1, cutAt = 0;   // the right part starts at index cutAt (bb-relative) of direction cutDir
    // (a) a hole: the zero of a signature nearest the middle; the longer direction wins a tie
    {
      double best = 1e300;
      for (int d = 0; d < 3; d++) {
        const int len = (int)sig[d].size();
        for (int x = 1; x < len - 1; x++)
          if (sig[d][x] == 0) {
            const double dist = std::fabs(x - 0.5 * (len - 1)) / len;
            if (dist < best - 1e-12) { best = dist; cutDir = d; cutAt = x; }
          }
      }
      if (cutDir >= 0) {   // step past the hole: the right part starts at the next non-empty plane
        while (cutAt < (int)sig[cutDir].size() && sig[cutDir][cutAt] == 0) cutAt++;
      }
    }
    // (b) the strongest inflection of the signature's second difference
    if (cutDir < 0) {
      long long bestJump = 0;
      double bestDist = 1e300;
      for (int d = 0; d < 3; d++) {
        const int len = (int)sig[d].size();
        if (len < 4) continue;
        std::vector<long long> lap(len, 0);
        for (int x = 1; x < len - 1; x++) lap[x] = sig[d][x - 1] - 2 * sig[d][x] + sig[d][x + 1];
        for (int x = 1; x < len - 2; x++) {
          if ((lap[x] < 0) == (lap[x + 1] < 0) || lap[x] == 0 || lap[x + 1] == 0) continue;   // a sign change between x and x+1
          const long long jump = std::llabs(lap[x + 1] - lap[x]);
          const double dist = std::fabs(x + 0.5 - 0.5 * (len - 1)) / len;
          if (jump > bestJump || (jump == bestJump && dist < bestDist - 1e-12)) { bestJump = jump; bestDist = dist; cutDir = d; cutAt = x + 1; }
        }
      }
    }
    // (c) bisect the longest edge
    if (cutDir < 0) {
      int len = 0;
      for (int d = 0; d < 3; d++)
        if ((int)sig[d].size() > len) { len = (int)sig[d].size(); cutDir = d; }
      if (len < 2) {   // one cell that is tagged but may not be refined (outside the nesting domain): drop it
        if (isNested) accept(bb);
        return;
      }
      cutAt = len / 2;
    }
    std::vector<Pt> left, right;
    const int cut = bb.lo[cutDir] + cutAt;
    for (const Pt &p : pts) (p.x[cutDir] < cut ? left : right).push_back(p);
    if (left.empty() || right.empty()) {   // cannot happen for the cuts above; guard against endless recursion
      accept(bb);
      return;
    }
    std::vector<Pt>().swap(pts);
    run(left);
    run(right);
  }
};

}  // namespace

struct mgic_grids {
  mgic_params P;
  double refineThresh = 0.1, fillRatio = 0.5;
  int nestingRadius = 2, tagsGrow = 2;
  std::vector<std::vector<IBox>> lev;   // boxes per level, level 0 = the domainSplit lattice
  std::vector<double> maxCondition;     // max |regrid condition| per level at the last pass
  std::vector<long long> tagged;        // tagged cells per level (after growing) at the last pass
};

namespace {

void level_domain(const mgic_params &P, int level, int n[3]) {
  for (int d = 0; d < 3; d++) n[d] = P.N[d] << level;   // refRatio == 2 everywhere (PoissonParameters.cpp:75-79)
}

// [Chombo] MeshRefine::regrid restated (see the header of this file): tags[l] = grown, clipped tag points of level l for
// l = 0 .. top; `cur` = the present boxes of levels 0 .. top.  Returns the new boxes of levels 1 .. top+1 in out[1 ..].
void regrid(const mgic_grids &G, const std::vector<std::vector<Pt>> &tags, int top, int maxLevel, std::vector<std::vector<IBox>> &out) {
  const int bf = G.P.block_factor, cf = std::max(1, bf / 2);   // level-l cells per block of level l+1's block factor
  const int maxC = std::max(1, G.P.max_grid_size / bf);
  out.assign(top + 2, std::vector<IBox>());
  for (int l = std::min(top, maxLevel - 1); l >= 0; l--) {
    int ndom[3];
    level_domain(G.P, l, ndom);
    const Bits pnd = nesting_domain(G.lev[l], ndom, G.nestingRadius);
    // blocks of cf^3 cells that lie wholly in the nesting domain
    IBox cb;
    for (int d = 0; d < 3; d++) { cb.lo[d] = floordiv(pnd.b.lo[d], cf); cb.hi[d] = floordiv(pnd.b.hi[d], cf); }
    Bits okC;
    okC.define(cb);
    for (int K = cb.lo[2]; K <= cb.hi[2]; K++)
      for (int J = cb.lo[1]; J <= cb.hi[1]; J++)
        for (int I = cb.lo[0]; I <= cb.hi[0]; I++) {
          bool all = true;
          for (int k = 0; k < cf && all; k++)
            for (int j = 0; j < cf && all; j++)
              for (int i = 0; i < cf && all; i++) all = pnd.get(I * cf + i, J * cf + j, K * cf + k);
          if (all) okC.v[okC.at(I, J, K)] = 1;
        }
    // tagged blocks: this level's tags + what the new level l+2 needs below it
    Bits tagC;
    tagC.define(cb);
    for (const Pt &p : tags[l]) {
      const int I = floordiv(p.x[0], cf), J = floordiv(p.x[1], cf), K = floordiv(p.x[2], cf);
      if (okC.get(I, J, K)) tagC.set(I, J, K);
    }
    if (l + 2 < (int)out.size())
      for (const IBox &q : out[l + 2]) {   // level l+2 box -> level l+1 cells, grown by the nesting radius -> level l blocks
        IBox g;
        for (int d = 0; d < 3; d++) {
          g.lo[d] = floordiv(floordiv(floordiv(q.lo[d], 2) - G.nestingRadius, 2), cf);
          g.hi[d] = floordiv(floordiv(floordiv(q.hi[d], 2) + G.nestingRadius, 2), cf);
        }
        for (int K = g.lo[2]; K <= g.hi[2]; K++)
          for (int J = g.lo[1]; J <= g.hi[1]; J++)
            for (int I = g.lo[0]; I <= g.hi[0]; I++)
              if (okC.get(I, J, K)) tagC.set(I, J, K);
      }
    std::vector<Pt> pts;
    for (int K = cb.lo[2]; K <= cb.hi[2]; K++)
      for (int J = cb.lo[1]; J <= cb.hi[1]; J++)
        for (int I = cb.lo[0]; I <= cb.hi[0]; I++)
          if (tagC.v[tagC.at(I, J, K)]) pts.push_back({{I, J, K}});
    Cluster C;
    C.ok = &okC; C.fill = G.fillRatio; C.maxSize = maxC;
    C.run(pts);
    const int rf = cf * 2;   // block -> level l+1 cells
    for (const IBox &q : C.out) {
      IBox f;
      for (int d = 0; d < 3; d++) { f.lo[d] = q.lo[d] * rf; f.hi[d] = (q.hi[d] + 1) * rf - 1; }
      out[l + 1].push_back(f);
    }
    std::sort(out[l + 1].begin(), out[l + 1].end(), [](const IBox &a, const IBox &b) {   // z, then y, then x: a stable, documented order
      if (a.lo[2] != b.lo[2]) return a.lo[2] < b.lo[2];
      if (a.lo[1] != b.lo[1]) return a.lo[1] < b.lo[1];
      return a.lo[0] < b.lo[0];
    });
  }
}

// connected parts of a level: boxes whose closures meet (faces, edges or corners) belong together
void components(const std::vector<IBox> &boxes, std::vector<int> &comp, int *ncomp) {
  const int n = (int)boxes.size();
  std::vector<int> parent(n);
  std::iota(parent.begin(), parent.end(), 0);
  std::function<int(int)> find = [&](int a) { return parent[a] == a ? a : parent[a] = find(parent[a]); };
  for (int a = 0; a < n; a++)
    for (int b = a + 1; b < n; b++) {
      bool meet = true;
      for (int d = 0; d < 3; d++) meet = meet && boxes[a].lo[d] <= boxes[b].hi[d] + 1 && boxes[b].lo[d] <= boxes[a].hi[d] + 1;
      if (meet) parent[find(a)] = find(b);
    }
  comp.assign(n, -1);
  int nc = 0;
  std::vector<int> id(n, -1);
  for (int a = 0; a < n; a++) {
    const int r = find(a);
    if (id[r] < 0) id[r] = nc++;
    comp[a] = id[r];
  }
  *ncomp = nc;
}

// set_regrid_condition + set_tag_cells on one level: tag points (grown by tagsGrow, clipped to the domain)
int tag_level(mgic_ctx *c, mgic_grids *G, int level, std::vector<Pt> &pts) {
  const std::vector<IBox> &boxes = G->lev[level];
  int ndom[3];
  level_domain(G->P, level, ndom);
  const double dx = (G->P.L / G->P.N[0]) / (double)(1 << level);
  std::vector<int> comp;
  int nc = 0;
  components(boxes, comp, &nc);
  // pass 1: the condition on every connected part (device), kept on the host with the part's cell mask
  struct Part { IBox bb; Bits cells; std::vector<double> val; };
  std::vector<Part> parts(nc);
  double vmax = 0.0;
  for (int q = 0; q < nc; q++) {
    std::vector<IBox> mine;
    for (size_t b = 0; b < boxes.size(); b++)
      if (comp[b] == q) mine.push_back(boxes[b]);
    Part &pt = parts[q];
    pt.bb = bounding(mine);
    pt.cells.define(pt.bb);
    for (const IBox &b : mine) pt.cells.fill(b);
    const int n[3] = {pt.cells.n[0], pt.cells.n[1], pt.cells.n[2]};
    const size_t cells = pt.cells.v.size();
    double *d = nullptr;
    MGIC_CUDA(mgic_dev_malloc(&d, cells * sizeof(double)));
    int rc = mgk::condition_box(c, G->P, dx, pt.bb.lo, n, 0, d);
    pt.val.resize(cells);
    if (rc == MGIC_OK && cudaMemcpyAsync(pt.val.data(), d, cells * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = MGIC_ERR_CUDA;
    if (rc == MGIC_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = MGIC_ERR_CUDA;
    mgic_dev_free(d);
    if (rc != MGIC_OK) { mgic_set_error("regrid condition on level %d failed", level); return rc; }
    for (size_t i = 0; i < cells; i++)
      if (pt.cells.v[i]) vmax = std::max(vmax, std::fabs(pt.val[i]));   // norm(levelRhs, interval, 0), SetGrids.cpp:184
  }
  G->maxCondition[level] = vmax;
  const double tagVal = vmax * G->refineThresh;                          // :186
  // pass 2: tag, grow, clip (:189-202)
  IBox all = bounding(boxes);
  for (int d = 0; d < 3; d++) { all.lo[d] = std::max(0, all.lo[d] - G->tagsGrow); all.hi[d] = std::min(ndom[d] - 1, all.hi[d] + G->tagsGrow); }
  Bits tg;
  tg.define(all);
  const int g = G->tagsGrow;
  for (const Part &pt : parts)
    for (int k = pt.bb.lo[2]; k <= pt.bb.hi[2]; k++)
      for (int j = pt.bb.lo[1]; j <= pt.bb.hi[1]; j++)
        for (int i = pt.bb.lo[0]; i <= pt.bb.hi[0]; i++) {
          const size_t a = pt.cells.at(i, j, k);
          if (!pt.cells.v[a] || !(std::fabs(pt.val[a]) >= tagVal)) continue;
          tg.v[tg.at(i, j, k)] = 1;
        }
  for (int axis = 0; axis < 3; axis++) filter_axis(tg, axis, g, true, ndom);   // tags.grow(tagsGrow), clipped to the domain by `all`
  pts.clear();
  for (int k = all.lo[2]; k <= all.hi[2]; k++)
    for (int j = all.lo[1]; j <= all.hi[1]; j++)
      for (int i = all.lo[0]; i <= all.hi[0]; i++)
        if (tg.v[tg.at(i, j, k)]) pts.push_back({{i, j, k}});
  G->tagged[level] = (long long)pts.size();
  return MGIC_OK;
}

int check_params(const mgic_params *P) {
  MGIC_REQUIRE(P->block_factor >= 2 && (P->block_factor & (P->block_factor - 1)) == 0, "block_factor must be a power of two >= 2");
  MGIC_REQUIRE(P->max_grid_size >= P->block_factor && P->max_grid_size % P->block_factor == 0, "max_grid_size must be a multiple of block_factor");
  for (int d = 0; d < 3; d++) MGIC_REQUIRE(P->N[d] >= 1 && P->N[d] % P->block_factor == 0, "N must be a multiple of block_factor");
  MGIC_REQUIRE(P->max_level >= 0 && P->max_level <= 12, "max_level out of range");
  return MGIC_OK;
}

}  // namespace

extern "C" int mgic_grids_destroy(mgic_grids *G) {
  delete G;
  return MGIC_OK;
}

// BRMeshRefine::regrid alone, on tags the caller supplies (host only: no device is touched).  levels 0 .. top exist:
// nboxes[l] boxes each, flattened in `boxes` (6 ints); ntags[l] tag cells each, flattened in `tags` (3 ints, level l's
// index space, already grown / clipped as the caller wishes).  The new boxes of levels 1 .. top+1 replace the old ones.
extern "C" int mgic_grids_regrid(const mgic_params *P, double fill_ratio, int top, const int *nboxes, const int *boxes, const int *ntags,
                                 const int *tags, mgic_grids **out) {
  MGIC_REQUIRE(P && out && top >= 0 && nboxes && boxes && ntags, "bad argument");
  MGIC_TRY(check_params(P));
  MGIC_REQUIRE(fill_ratio > 0.0 && fill_ratio <= 1.0, "fill_ratio must be in (0, 1]");
  mgic_grids *G = new mgic_grids;
  G->P = *P; G->fillRatio = fill_ratio;
  G->lev.resize(top + 1);
  std::vector<std::vector<Pt>> tg(top + 1);
  size_t qb = 0, qt = 0;
  for (int l = 0; l <= top; l++) {
    for (int b = 0; b < nboxes[l]; b++, qb++) {
      IBox q;
      for (int d = 0; d < 3; d++) { q.lo[d] = boxes[6 * qb + d]; q.hi[d] = boxes[6 * qb + 3 + d]; }
      G->lev[l].push_back(q);
    }
    for (int t = 0; t < ntags[l]; t++, qt++) tg[l].push_back({{tags[3 * qt], tags[3 * qt + 1], tags[3 * qt + 2]}});
  }
  std::vector<std::vector<IBox>> nw;
  regrid(*G, tg, top, top + 1, nw);
  G->lev.resize(top + 2);
  for (int l = 1; l <= top + 1; l++) G->lev[l] = nw[l];
  while (G->lev.size() > 1 && G->lev.back().empty()) G->lev.pop_back();
  G->maxCondition.assign(G->lev.size(), 0.0);
  G->tagged.assign(G->lev.size(), 0);
  *out = G;
  return MGIC_OK;
}

// set_grids (Source/SetGrids.cpp:31-148): needs a device (the regrid condition is evaluated there)
extern "C" int mgic_grids_generate(mgic_ctx *c, const mgic_params *P, double refine_threshold, double fill_ratio, mgic_grids **out) {
  MGIC_REQUIRE(c && P && out, "NULL argument");
  MGIC_TRY(check_params(P));
  MGIC_REQUIRE(fill_ratio > 0.0 && fill_ratio <= 1.0, "fill_ratio must be in (0, 1]");
  MGIC_REQUIRE(refine_threshold >= 0.0, "refine_threshold must not be negative");
  MGIC_CUDA(cudaSetDevice(c->device));
  mgic_grids *G = new mgic_grids;
  G->P = *P; G->refineThresh = refine_threshold; G->fillRatio = fill_ratio;
  const int maxLevel = P->max_level;
  G->lev.resize(1);
  domain_split(P->N, P->max_grid_size, G->lev[0]);                         // :54-58
  int topLevel = 0;
  bool moreLevels = maxLevel > 0;                                          // :61-62
  while (moreLevels) {                                                     // :70
    moreLevels = false;
    const int oldTop = topLevel;
    G->maxCondition.assign(topLevel + 1, 0.0);
    G->tagged.assign(topLevel + 1, 0);
    std::vector<std::vector<Pt>> tags(topLevel + 1);
    for (int l = 0; l <= topLevel; l++) {                                  // :79-110
      const int rc = tag_level(c, G, l, tags[l]);
      if (rc != MGIC_OK) { delete G; return rc; }
    }
    std::vector<std::vector<IBox>> nw;
    regrid(*G, tags, topLevel, maxLevel, nw);                              // :113-114
    int newFinest = 0;
    for (int l = 1; l < (int)nw.size(); l++)
      if (!nw[l].empty()) newFinest = l;
    if (newFinest > topLevel) topLevel++;                                  // :116-118
    G->lev.resize(topLevel + 1);
    for (int l = 1; l <= topLevel; l++) G->lev[l] = nw[l];                 // :120-132
    while (G->lev.size() > 1 && G->lev.back().empty()) { G->lev.pop_back(); topLevel--; }
    if (topLevel < maxLevel && topLevel > oldTop) moreLevels = true;       // :135-136
  }
  G->maxCondition.resize(G->lev.size(), 0.0);
  G->tagged.resize(G->lev.size(), 0);
  *out = G;
  return MGIC_OK;
}

extern "C" int mgic_grids_levels(const mgic_grids *G) { return G ? (int)G->lev.size() : 0; }
extern "C" int mgic_grids_num_boxes(const mgic_grids *G, int level) {
  return (G && level >= 0 && level < (int)G->lev.size()) ? (int)G->lev[level].size() : 0;
}
// boxes of one level (6 ints each) and, optionally, the connected part each belongs to (0 .. *nparts-1)
extern "C" int mgic_grids_get_boxes(const mgic_grids *G, int level, int *boxes, int *part_of_box, int *nparts) {
  MGIC_REQUIRE(G && level >= 0 && level < (int)G->lev.size(), "bad argument");
  const std::vector<IBox> &b = G->lev[level];
  if (boxes)
    for (size_t q = 0; q < b.size(); q++)
      for (int d = 0; d < 3; d++) { boxes[6 * q + d] = b[q].lo[d]; boxes[6 * q + 3 + d] = b[q].hi[d]; }
  std::vector<int> comp;
  int nc = 0;
  components(b, comp, &nc);
  if (part_of_box) std::copy(comp.begin(), comp.end(), part_of_box);
  if (nparts) *nparts = nc;
  return MGIC_OK;
}
extern "C" int mgic_grids_level_stats(const mgic_grids *G, int level, double *max_condition, long long *tagged_cells, long long *cells) {
  MGIC_REQUIRE(G && level >= 0 && level < (int)G->lev.size(), "bad argument");
  if (max_condition) *max_condition = G->maxCondition[level];
  if (tagged_cells) *tagged_cells = G->tagged[level];
  if (cells) {
    long long n = 0;
    for (const IBox &q : G->lev[level]) n += q.vol();
    *cells = n;
  }
  return MGIC_OK;
}

// the hierarchy of grids as the problem object: every level > 0 becomes its connected parts, each one masked array
extern "C" int mgic_hier_create_from_grids(mgic_ctx *c, const mgic_params *P, const mgic_grids *G, mgic_hier **out) {
  MGIC_REQUIRE(c && P && G && out, "NULL argument");
  const int nfiner = (int)G->lev.size() - 1;
  std::vector<int> nnodes, nboxes, flat;
  for (int l = 1; l <= nfiner; l++) {
    std::vector<int> comp;
    int nc = 0;
    components(G->lev[l], comp, &nc);
    nnodes.push_back(nc);
    for (int q = 0; q < nc; q++) {
      int nb = 0;
      for (size_t b = 0; b < G->lev[l].size(); b++)
        if (comp[b] == q) {
          nb++;
          for (int d = 0; d < 3; d++) flat.push_back(G->lev[l][b].lo[d]);
          for (int d = 0; d < 3; d++) flat.push_back(G->lev[l][b].hi[d]);
        }
      nboxes.push_back(nb);
    }
  }
  if (nnodes.empty()) nnodes.push_back(0);
  if (nboxes.empty()) nboxes.push_back(0);
  if (flat.empty()) flat.push_back(0);
  return mgic_hier_create(c, P, nfiner, nnodes.data(), nboxes.data(), flat.data(), out);
}
