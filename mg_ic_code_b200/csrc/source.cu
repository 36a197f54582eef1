// source.cu -- device evaluation of the problem data that feeds the operator:
// set_initial_conditions, set_rhs, set_a_coef, set_update_psi0 of Source/SetLevelData.cpp with the two
// stencil kernels of Source/SetLevelDataF.ChF and the Bowen-York formulas of Source/SetBinaryBH.H fused in.
//
// multigrid_vars lives in HBM as 8 components (MultigridUserVariables.hpp:10-23), each padded by ONE ghost
// layer on every side (the reference allocates three, Main_PoissonSolver.cpp:79, but its stencils read one).
// Arithmetic keeps the reference's evaluation order; exp() and the psi_0 powers are the only operations
// that are not bit-reproducible against libm (tested to 1e-13 relative).
#include "mgic_internal.h"

namespace {

enum { c_psi = 0, c_A11 = 1, c_A12 = 2, c_A13 = 3, c_A22 = 4, c_A23 = 5, c_A33 = 6, c_phi = 7 };

struct SrcP {
  double G_Newton, phi_amplitude, phi_wavelength;
  double m1, m2, spin1, spin2, mom1, mom2, off1, off2;
  double dx, half[3];  // dx = this level's spacing; half[d] = (coarsest dx * N[d]) / 2 = domainLength / 2
  int nx, ny, nz, k0;
  long long sy, sz, sc;
  // AMR level > 0 (mgic_vars_create_patch): the array covers the level's bounding box whose lower corner is (i0, j0, k0) in
  // the level's index space; mask = the cells of the level's boxes (null: all); psiG[f][cell] = psi in the coarse-fine
  // ghost cell beyond face f of `cell` (the fields carry no room for it: one ghost position can belong to two faces)
  int i0, j0, patch;
  int ndom[3];
  const unsigned char *mask;
  double *psiG[6];
};

__device__ __forceinline__ double eps3(int a, int b, int c) {
  // Levi-Civita entries set at Source/SetBinaryBH.H:30-36
  if ((a == 0 && b == 1 && c == 2) || (a == 1 && b == 2 && c == 0) || (a == 2 && b == 0 && c == 1)) return 1.0;
  if ((a == 0 && b == 2 && c == 1) || (a == 2 && b == 1 && c == 0) || (a == 1 && b == 0 && c == 2)) return -1.0;
  return 0.0;
}

// get_Aij -- Source/SetBinaryBH.H:24-52 (Alcubierre 3.4.22), accumulation order as written there
__device__ double get_Aij(int i, int j, double r1, double r2, const double *n1, const double *n2, const double *J1,
                          const double *J2, const double *P1, const double *P2) {
  double Aij = 1.5 / r1 / r1 * (n1[i] * P1[j] + n1[j] * P1[i]) + 1.5 / r2 / r2 * (n2[i] * P2[j] + n2[j] * P2[i]);
  const double dij = (i == j) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    Aij += 1.5 / r1 / r1 * (n1[i] * n1[j] - dij) * P1[k] * n1[k] + 1.5 / r2 / r2 * (n2[i] * n2[j] - dij) * P2[k] * n2[k];
#pragma unroll
    for (int l = 0; l < 3; l++) {
      Aij += -3.0 / r1 / r1 / r1 * (eps3(i, l, k) * n1[j] + eps3(j, l, k) * n1[i]) * n1[l] * J1[k] -
             3.0 / r2 / r2 / r2 * (eps3(i, l, k) * n2[j] + eps3(j, l, k) * n2[i]) * n2[l] * J2[k];
    }
  }
  return Aij;
}

// cell centre: Source/SetLevelData.cpp:58-60   loc = (iv + 0.5) * dx - domainLength / 2
__device__ __forceinline__ void cell_loc(const SrcP &P, int i, int j, int kg, double loc[3]) {
  const int iv[3] = {i, j, kg};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double l = iv[d] + 0.5 * 1.0;
    l *= P.dx;
    l -= P.half[d];
    loc[d] = l;
  }
}

// set_initial_conditions over the ghosted box -- Source/SetLevelData.cpp:32-71, SetBinaryBH.H:54-83, MyPhiFunction.H:11-16
__global__ void __launch_bounds__(128) k_init(SrcP P, double *__restrict__ mv) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - 1;
  const int j = (int)(blockIdx.y * blockDim.y + threadIdx.y) - 1;
  const int k = (int)blockIdx.z - 1;
  if (i > P.nx || j > P.ny) return;
  const long long q = (i + 1) + (j + 1) * P.sy + (k + 1) * P.sz;
  double loc[3];
  cell_loc(P, i + P.i0, j + P.j0, k + P.k0, loc);
  mv[q + c_psi * P.sc] = 1.0;
  const double r2 = loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2];
  mv[q + c_phi * P.sc] = P.phi_amplitude * exp(-r2 / P.phi_wavelength);
  double l1[3] = {loc[0] - P.off1, loc[1], loc[2]};
  double l2[3] = {loc[0] - P.off2, loc[1], loc[2]};
  const double r1 = sqrt(l1[0] * l1[0] + l1[1] * l1[1] + l1[2] * l1[2]);
  const double rr2 = sqrt(l2[0] * l2[0] + l2[1] * l2[1] + l2[2] * l2[2]);
  const double n1[3] = {l1[0] / r1, l1[1] / r1, l1[2] / r1};
  const double n2[3] = {l2[0] / rr2, l2[1] / rr2, l2[2] / rr2};
  const double J1[3] = {0.0, 0.0, P.spin1}, J2[3] = {0.0, 0.0, P.spin2};
  const double P1[3] = {0.0, P.mom1, 0.0}, P2[3] = {0.0, P.mom2, 0.0};
  mv[q + c_A11 * P.sc] = get_Aij(0, 0, r1, rr2, n1, n2, J1, J2, P1, P2);
  mv[q + c_A22 * P.sc] = get_Aij(1, 1, r1, rr2, n1, n2, J1, J2, P1, P2);
  mv[q + c_A33 * P.sc] = get_Aij(2, 2, r1, rr2, n1, n2, J1, J2, P1, P2);
  mv[q + c_A12 * P.sc] = get_Aij(0, 1, r1, rr2, n1, n2, J1, J2, P1, P2);
  mv[q + c_A13 * P.sc] = get_Aij(0, 2, r1, rr2, n1, n2, J1, J2, P1, P2);
  mv[q + c_A23 * P.sc] = get_Aij(1, 2, r1, rr2, n1, n2, J1, J2, P1, P2);
}

// set_rhs + set_a_coef fused -- Source/SetLevelData.cpp:73-127, 281-325; SetLevelDataF.ChF:15-58, 65-103
__global__ void __launch_bounds__(128) k_rhs_acoef(SrcP P, const double *__restrict__ mv, double *__restrict__ rhs,
                                                   double *__restrict__ aC, double constant_K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= P.nx || j >= P.ny) return;
  const long long q = (i + 1) + (j + 1) * P.sy + (k + 1) * P.sz;
  const long long o = i + (long long)j * P.nx + (long long)k * P.nx * P.ny;
  if (P.mask && !P.mask[o]) return;   // not a cell of the level's boxes
  const long long st[3] = {1, P.sy, P.sz};
  const double *psi = mv + c_psi * P.sc, *phi = mv + c_phi * P.sc;
  // GETRHOGRADPHIF: rho = sum_d 0.5 * (0.5/dx * (phi+ - phi-))^2
  double rho = 0.0;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double g = 0.5 / P.dx * (phi[q + st[d]] - phi[q - st[d]]);
    rho = rho + 0.5 * g * g;
  }
  // GETLAPLACIANPSIF: lap = sum_d 1/dx/dx * (psi- - 2 psi + psi+)
  double lap = 0.0;
  if (!P.patch) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const double dd = 1.0 / P.dx / P.dx * (+1.0 * psi[q - st[d]] - 2.0 * psi[q] + 1.0 * psi[q + st[d]]);
      lap = lap + dd;
    }
  } else {
    // psi's neighbour: a cell of the level or a physical ghost -> the padded array (what exchange / the solver left there);
    // a coarse-fine ghost -> psiG (QuadCFInterp'd dpsi accumulated onto the initial 1, Main_PoissonSolver.cpp:193-205)
    const int iv[3] = {i, j, k}, n[3] = {P.nx, P.ny, P.nz}, lo[3] = {P.i0, P.j0, P.k0};
    const long long so[3] = {1, (long long)P.nx, (long long)P.nx * P.ny};
#pragma unroll
    for (int d = 0; d < 3; d++) {
      double pm, pp;
      {
        const int w = iv[d] - 1;
        const bool cell = w >= 0 && (!P.mask || P.mask[o - so[d]]);
        pm = (cell || lo[d] + w < 0) ? psi[q - st[d]] : P.psiG[2 * d][o];
      }
      {
        const int w = iv[d] + 1;
        const bool cell = w < n[d] && (!P.mask || P.mask[o + so[d]]);
        pp = (cell || lo[d] + w >= P.ndom[d]) ? psi[q + st[d]] : P.psiG[2 * d + 1][o];
      }
      const double dd = 1.0 / P.dx / P.dx * (+1.0 * pm - 2.0 * psi[q] + 1.0 * pp);
      lap = lap + dd;
    }
  }
  double loc[3];
  cell_loc(P, i + P.i0, j + P.j0, k + P.k0, loc);
  // set_m_value (:266-278): rho_matter = 0
  const double rho_m = 0.5 * 0.0 * 0.0 + 0.0;
  const double m = (2.0 / 3.0) * (constant_K * constant_K) - 16.0 * M_PI * P.G_Newton * rho_m;
  const double a11 = mv[q + c_A11 * P.sc], a22 = mv[q + c_A22 * P.sc], a33 = mv[q + c_A33 * P.sc];
  const double a12 = mv[q + c_A12 * P.sc], a13 = mv[q + c_A13 * P.sc], a23 = mv[q + c_A23 * P.sc];
  const double A2 = a11 * a11 + a22 * a22 + a33 * a33 + 2 * (a12 * a12) + 2 * (a13 * a13) + 2 * (a23 * a23);  // :110-116
  // set_binary_bh_psi (SetBinaryBH.H:85-99)
  const double x1 = loc[0] - P.off1, x2 = loc[0] - P.off2;
  const double r1 = sqrt(x1 * x1 + loc[1] * loc[1] + loc[2] * loc[2]);
  const double r2 = sqrt(x2 * x2 + loc[1] * loc[1] + loc[2] * loc[2]);
  const double psi_0 = psi[q] + (P.m1 / r1 + P.m2 / r2);  // :118-119
  const double p2 = psi_0 * psi_0, p4 = p2 * p2, p5 = p4 * psi_0, p7 = p4 * p2 * psi_0, p8 = p4 * p4;
  if (aC) aC[o] = -0.625 * m * p4 - A2 * (1.0 / p8) + 2.0 * M_PI * P.G_Newton * rho;  // :321-322
  if (rhs) rhs[o] = 0.125 * m * p5 - 0.125 * A2 * (1.0 / p7) - 2.0 * M_PI * P.G_Newton * rho * psi_0 - lap;  // :121-124
}

// set_regrid_condition (Source/SetLevelData.cpp:188-240) on the data set_grids evaluates it on (Source/SetGrids.cpp:86-95:
// fresh set_initial_conditions -- psi = 1, phi and A_ij analytic, ghost cells included): a function of the cell position
// alone, so it is evaluated analytically over any index box [lo, lo + n) of a level with spacing dx -- no multigrid_vars
// array, no dependence on how the level is cut into boxes.  mode 1: set_constant_K_integrand (:128-186) on the same data.
__device__ __forceinline__ double phi_at(const SrcP &P, int i, int j, int k) {
  double loc[3];
  cell_loc(P, i, j, k, loc);
  const double r2 = loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2];
  return P.phi_amplitude * exp(-r2 / P.phi_wavelength);   // MyPhiFunction.H:11-16
}
__global__ void __launch_bounds__(128) k_condition(SrcP P, int mode, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= P.nx || j >= P.ny) return;
  const int gi = i + P.i0, gj = j + P.j0, gk = k + P.k0;
  // GETRHOGRADPHIF on the analytic phi (SetLevelDataF.ChF:65-103)
  double rho = 0.0;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double g = 0.5 / P.dx * (phi_at(P, gi + (d == 0), gj + (d == 1), gk + (d == 2)) - phi_at(P, gi - (d == 0), gj - (d == 1), gk - (d == 2)));
    rho = rho + 0.5 * g * g;
  }
  double loc[3];
  cell_loc(P, gi, gj, gk, loc);
  const double m = (2.0 / 3.0) * (0.0 * 0.0) - 16.0 * M_PI * P.G_Newton * (0.5 * 0.0 * 0.0 + 0.0);   // set_m_value(m, phi, params, 0.0)
  double l1[3] = {loc[0] - P.off1, loc[1], loc[2]};
  double l2[3] = {loc[0] - P.off2, loc[1], loc[2]};
  const double r1 = sqrt(l1[0] * l1[0] + l1[1] * l1[1] + l1[2] * l1[2]);
  const double rr2 = sqrt(l2[0] * l2[0] + l2[1] * l2[1] + l2[2] * l2[2]);
  const double n1[3] = {l1[0] / r1, l1[1] / r1, l1[2] / r1};
  const double n2[3] = {l2[0] / rr2, l2[1] / rr2, l2[2] / rr2};
  const double J1[3] = {0.0, 0.0, P.spin1}, J2[3] = {0.0, 0.0, P.spin2};
  const double P1[3] = {0.0, P.mom1, 0.0}, P2[3] = {0.0, P.mom2, 0.0};
  const double a11 = get_Aij(0, 0, r1, rr2, n1, n2, J1, J2, P1, P2), a22 = get_Aij(1, 1, r1, rr2, n1, n2, J1, J2, P1, P2);
  const double a33 = get_Aij(2, 2, r1, rr2, n1, n2, J1, J2, P1, P2), a12 = get_Aij(0, 1, r1, rr2, n1, n2, J1, J2, P1, P2);
  const double a13 = get_Aij(0, 2, r1, rr2, n1, n2, J1, J2, P1, P2), a23 = get_Aij(1, 2, r1, rr2, n1, n2, J1, J2, P1, P2);
  const double A2 = a11 * a11 + a22 * a22 + a33 * a33 + 2 * (a12 * a12) + 2 * (a13 * a13) + 2 * (a23 * a23);
  const double psi_0 = 1.0 + (P.m1 / r1 + P.m2 / rr2);
  const double p2 = psi_0 * psi_0, p4 = p2 * p2;
  double v;
  if (mode == 0)   // :233-236
    v = 1.5 * fabs(m) + 1.5 * A2 * (1.0 / (p4 * p2 * psi_0)) + 24.0 * M_PI * P.G_Newton * fabs(rho) * psi_0 + log(psi_0);
  else             // :180-183 with laplacian(psi = 1) = 0
    v = -1.5 * m + 1.5 * A2 * (1.0 / (p4 * p4 * p4)) + 24.0 * M_PI * P.G_Newton * rho * (1.0 / p4) + 12.0 * 0.0 * (1.0 / (p4 * psi_0));
  out[i + (long long)j * P.nx + (long long)k * P.nx * P.ny] = v;
}

// set_output_data (Source/SetLevelData.cpp:343-396) for ONE box of a level with the checkpoint's ghost layers
// (Source/WriteOutput.H:180: three): the 32 GRChombo variables (GRChomboUserVariables.hpp), component slowest, over the box
// grown by `ng`.  psi of a ghost cell is what the reference's multigrid_vars holds there after the nonlinear loop: a cell of
// the level -> that cell (the 3-ghost exchange of set_update_psi0); the first layer beyond a face of the box -> the physical
// ghost of the padded array or the accumulated QuadCFInterp ghost psiG; anything else was never written: the initial 1.
// phi and A_ij are the analytic initial data everywhere (set_initial_conditions fills ghost cells, nothing modifies them).
struct OutBox { int lo[3], n[3], ng; };   // the box (level index space), its size, ghost width
__global__ void __launch_bounds__(128) k_output_box(SrcP P, OutBox B, const double *__restrict__ mv, double constant_K, double *__restrict__ out) {
  const int gnx = B.n[0] + 2 * B.ng, gny = B.n[1] + 2 * B.ng, gnz = B.n[2] + 2 * B.ng;
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y, c = blockIdx.z;
  if (a >= gnx || b >= gny) return;
  const int g[3] = {B.lo[0] - B.ng + a, B.lo[1] - B.ng + b, B.lo[2] - B.ng + c};      // the cell in the level's index space
  const int lo[3] = {P.i0, P.j0, P.k0}, n[3] = {P.nx, P.ny, P.nz};
  const int l[3] = {g[0] - lo[0], g[1] - lo[1], g[2] - lo[2]};                         // in the node's array
  const long long so[3] = {1, (long long)P.nx, (long long)P.nx * P.ny}, st[3] = {1, P.sy, P.sz};
  auto level_cell = [&](const int *q) {
    if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= n[0] || q[1] >= n[1] || q[2] >= n[2]) return false;
    return !P.mask || P.mask[q[0] + q[1] * so[1] + q[2] * so[2]] != 0;
  };
  const double *psi = mv + c_psi * P.sc;
  double psi_here = 1.0;
  if (level_cell(l)) psi_here = psi[(l[0] + 1) + (l[1] + 1) * st[1] + (l[2] + 1) * st[2]];
  else {
    // one step beyond exactly one face of the BOX?
    int out_dirs = 0, f = -1;
    for (int d = 0; d < 3; d++) {
      const int r = g[d] - B.lo[d];
      if (r == -1) { out_dirs++; f = 2 * d; }
      else if (r == B.n[d]) { out_dirs++; f = 2 * d + 1; }
      else if (r < -1 || r > B.n[d]) out_dirs += 2;
    }
    if (out_dirs == 1) {
      const int d = f >> 1;
      if (g[d] < 0 || g[d] >= P.ndom[d]) psi_here = psi[(l[0] + 1) + (l[1] + 1) * st[1] + (l[2] + 1) * st[2]];   // physical ghost (padded array)
      else if (P.patch) {
        int q[3] = {l[0], l[1], l[2]};
        q[d] += (f & 1) ? -1 : 1;                                                      // the box's cell inside the face
        psi_here = P.psiG[f][q[0] + q[1] * so[1] + q[2] * so[2]];
      }
    }
  }
  double loc[3];
  cell_loc(P, g[0], g[1], g[2], loc);
  const double r2 = loc[0] * loc[0] + loc[1] * loc[1] + loc[2] * loc[2];
  const double phi = P.phi_amplitude * exp(-r2 / P.phi_wavelength);
  double l1[3] = {loc[0] - P.off1, loc[1], loc[2]};
  double l2[3] = {loc[0] - P.off2, loc[1], loc[2]};
  const double r1 = sqrt(l1[0] * l1[0] + l1[1] * l1[1] + l1[2] * l1[2]);
  const double rr2 = sqrt(l2[0] * l2[0] + l2[1] * l2[1] + l2[2] * l2[2]);
  const double n1[3] = {l1[0] / r1, l1[1] / r1, l1[2] / r1};
  const double n2[3] = {l2[0] / rr2, l2[1] / rr2, l2[2] / rr2};
  const double J1[3] = {0.0, 0.0, P.spin1}, J2[3] = {0.0, 0.0, P.spin2};
  const double P1[3] = {0.0, P.mom1, 0.0}, P2[3] = {0.0, P.mom2, 0.0};
  const double psi_0 = psi_here + (P.m1 / r1 + P.m2 / rr2);
  const double p2 = psi_0 * psi_0;
  const double chi = 1.0 / (p2 * p2);                    // pow(psi_0, -4)   :381-383
  const double factor = chi * sqrt(chi);                 // pow(chi, 1.5)    :384
  const long long np = (long long)gnx * gny * gnz, q = a + (long long)gnx * (b + (long long)gny * c);
  enum { o_chi = 0, o_h11 = 1, o_h22 = 4, o_h33 = 6, o_K = 7, o_A11 = 8, o_A12 = 9, o_A13 = 10, o_A22 = 11, o_A23 = 12, o_A33 = 13, o_lapse = 18,
         o_phi = 25, NV = 32 };
  for (int v = 0; v < NV; v++) out[v * np + q] = 0.0;    // :357-359
  out[o_h11 * np + q] = 1.0; out[o_h22 * np + q] = 1.0; out[o_h33 * np + q] = 1.0; out[o_lapse * np + q] = 1.0;   // :363-366
  out[o_K * np + q] = constant_K;                        // :369
  out[o_chi * np + q] = chi;
  out[o_phi * np + q] = phi;                             // :387
  out[o_A11 * np + q] = get_Aij(0, 0, r1, rr2, n1, n2, J1, J2, P1, P2) * factor;   // :388-393
  out[o_A12 * np + q] = get_Aij(0, 1, r1, rr2, n1, n2, J1, J2, P1, P2) * factor;
  out[o_A13 * np + q] = get_Aij(0, 2, r1, rr2, n1, n2, J1, J2, P1, P2) * factor;
  out[o_A22 * np + q] = get_Aij(1, 1, r1, rr2, n1, n2, J1, J2, P1, P2) * factor;
  out[o_A23 * np + q] = get_Aij(1, 2, r1, rr2, n1, n2, J1, J2, P1, P2) * factor;
  out[o_A33 * np + q] = get_Aij(2, 2, r1, rr2, n1, n2, J1, J2, P1, P2) * factor;
}

// set_update_psi0 -- Source/SetLevelData.cpp:243-263: psi += dpsi over the GHOSTED box.  dpsi's domain-face
// ghost is what the solver's last homogeneous BC fill left there: a*near (+0).
__global__ void __launch_bounds__(128) k_update_psi(SrcP P, Geom g, BCk bc, double *__restrict__ mv,
                                                    const double *__restrict__ dpsi) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - 1;
  const int j = (int)(blockIdx.y * blockDim.y + threadIdx.y) - 1;
  const int k = (int)blockIdx.z - 1;
  if (i > P.nx || j > P.ny) return;
  const int oi = (i < 0) + (i >= P.nx), oj = (j < 0) + (j >= P.ny), ok = (k < 0) + (k >= P.nz);
  if (oi + oj + ok > 1) return;  // edges / corners: never read by the 7-point stencils
  const int ci = min(max(i, 0), P.nx - 1), cj = min(max(j, 0), P.ny - 1), ck = min(max(k, 0), P.nz - 1);
  const long long cidx = ci + cj * g.sy + ck * g.sz;
  double v = dpsi[cidx];
  if (oi + oj + ok == 1) {
    const int f = oi ? (i < 0 ? 0 : 1) : (oj ? (j < 0 ? 2 : 3) : (k < 0 ? 4 : 5));
    if (bc.type[f] == MGIC_FACE_INTERIOR) v = dpsi[cidx + (k < 0 ? -g.sz : g.sz)];
    else if (bc.type[f] == MGIC_BC_PERIODIC) {
      const long long w = (f == 0) ? (g.nx - 1) : (f == 1) ? -(long long)(g.nx - 1)
                        : (f == 2) ? (long long)(g.ny - 1) * g.sy : (f == 3) ? -(long long)(g.ny - 1) * g.sy
                        : (f == 4) ? (long long)(g.nz - 1) * g.sz : -(long long)(g.nz - 1) * g.sz;
      v = dpsi[cidx + w];
    } else v = bc.a[f] * v + bc.b[f];
  }
  const long long q = (i + 1) + (j + 1) * P.sy + (k + 1) * P.sz;
  mv[q + c_psi * P.sc] += v;
}

// set_update_psi0 on an AMR level > 0: psi += dpsi on the level's cells; beyond each face of a cell either nothing (the
// neighbour is a cell of the level: the exchange's copy), the physical ghost of the padded array (a*near + b), or the
// coarse-fine ghost psiG += QuadCFInterp(dpsi) (cf[f][cell], filled just before by the operator).
__global__ void __launch_bounds__(128) k_update_psi_patch(SrcP P, BCk bc, double *__restrict__ mv, const double *__restrict__ dpsi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= P.nx || j >= P.ny) return;
  const long long o = i + (long long)j * P.nx + (long long)k * P.nx * P.ny;
  if (P.mask && !P.mask[o]) return;
  const long long q = (i + 1) + (j + 1) * P.sy + (k + 1) * P.sz;
  const long long st[3] = {1, P.sy, P.sz}, so[3] = {1, (long long)P.nx, (long long)P.nx * P.ny};
  const int iv[3] = {i, j, k}, n[3] = {P.nx, P.ny, P.nz}, lo[3] = {P.i0, P.j0, P.k0};
  const double v = dpsi[o];
  double *psi = mv + c_psi * P.sc;
  psi[q] += v;
  for (int f = 0; f < 6; f++) {
    const int d = f >> 1, side = (f & 1) ? 1 : -1;
    const int w = iv[d] + side;
    if (w >= 0 && w < n[d] && (!P.mask || P.mask[o + side * so[d]])) continue;
    if (lo[d] + w < 0 || lo[d] + w >= P.ndom[d]) psi[q + side * st[d]] += bc.a[f] * v + bc.b[f];
    else P.psiG[f][o] += bc.face[f][o];
  }
}

__global__ void k_fill(long long n, double *__restrict__ y, double v) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) y[q] = v;
}

SrcP make_srcp(const mgic_vars *v) {
  SrcP s;
  const mgic_params &P = v->P;
  s.G_Newton = P.G_Newton; s.phi_amplitude = P.phi_amplitude; s.phi_wavelength = P.phi_wavelength;
  s.m1 = P.bh1_bare_mass; s.m2 = P.bh2_bare_mass; s.spin1 = P.bh1_spin; s.spin2 = P.bh2_spin;
  s.mom1 = P.bh1_momentum; s.mom2 = P.bh2_momentum; s.off1 = P.bh1_offset; s.off2 = P.bh2_offset;
  s.dx = v->dx;
  for (int d = 0; d < 3; d++) s.half[d] = ((P.L / P.N[0]) * P.N[d]) / 2.0;  // PoissonParameters.cpp:82-85 (the coarsest level's dx)
  s.nx = v->n[0]; s.ny = v->n[1]; s.nz = v->nzl; s.k0 = v->k0;
  s.sy = v->sy; s.sz = v->sz; s.sc = v->sc;
  s.i0 = v->lo[0]; s.j0 = v->lo[1]; s.patch = v->isPatch ? 1 : 0;
  for (int d = 0; d < 3; d++) s.ndom[d] = v->isPatch ? v->ndom[d] : P.N[d];
  s.mask = v->mask;
  for (int f = 0; f < 6; f++) s.psiG[f] = v->psiG[f];
  return s;
}

int post(mgic_ctx *c, const char *what) {
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { mgic_set_error("kernel %s: %s", what, cudaGetErrorString(e)); return MGIC_ERR_CUDA; }
  return MGIC_OK;
}

}  // namespace

namespace mgk {

int init_conditions(mgic_vars *v) {
  SrcP s = make_srcp(v);
  dim3 blk(32, 4, 1), grd((s.nx + 2 + 31) / 32, (s.ny + 2 + 3) / 4, s.nz + 2);
  k_init<<<grd, blk, 0, v->ctx->stream>>>(s, v->d);
  return post(v->ctx, "init_conditions");
}

int set_rhs_acoef(mgic_vars *v, double *rhs, double *acoef, double constant_K) {
  SrcP s = make_srcp(v);
  dim3 blk(32, 4, 1), grd((s.nx + 31) / 32, (s.ny + 3) / 4, s.nz);
  k_rhs_acoef<<<grd, blk, 0, v->ctx->stream>>>(s, v->d, rhs, acoef, constant_K);
  return post(v->ctx, "set_rhs_acoef");
}

// the regrid condition / constant-K integrand on the index box [lo, lo + n) of a level with spacing dx (device array `out`)
int condition_box(mgic_ctx *c, const mgic_params &P, double dx, const int lo[3], const int n[3], int mode, double *out) {
  mgic_vars v;
  v.ctx = c; v.P = P;
  for (int d = 0; d < 3; d++) { v.n[d] = n[d]; v.lo[d] = lo[d]; }
  v.k0 = lo[2]; v.nzl = n[2]; v.dx = dx; v.isPatch = true;
  SrcP s = make_srcp(&v);
  dim3 blk(32, 4, 1), grd((s.nx + 31) / 32, (s.ny + 3) / 4, s.nz);
  k_condition<<<grd, blk, 0, c->stream>>>(s, mode, out);
  return post(c, "regrid_condition");
}

// the 32 GRChombo variables of one box [lo, lo + n) of the level `v` lives on, grown by ng ghost layers, into the device array out
int output_box(mgic_vars *v, const int lo[3], const int n[3], int ng, double constant_K, double *out) {
  SrcP s = make_srcp(v);
  OutBox B;
  for (int d = 0; d < 3; d++) { B.lo[d] = lo[d]; B.n[d] = n[d]; }
  B.ng = ng;
  dim3 blk(32, 4, 1), grd((n[0] + 2 * ng + 31) / 32, (n[1] + 2 * ng + 3) / 4, n[2] + 2 * ng);
  k_output_box<<<grd, blk, 0, v->ctx->stream>>>(s, B, v->d, constant_K, out);
  return post(v->ctx, "output_box");
}

int update_psi_patch(mgic_vars *v, const BCk &bc, const double *dpsi) {
  SrcP s = make_srcp(v);
  dim3 blk(32, 4, 1), grd((s.nx + 31) / 32, (s.ny + 3) / 4, s.nz);
  k_update_psi_patch<<<grd, blk, 0, v->ctx->stream>>>(s, bc, v->d, dpsi);
  return post(v->ctx, "update_psi_patch");
}

int fill(mgic_ctx *c, double *y, long long n, double value) {
  long long b = (n + 255) / 256;
  if (b > 4096) b = 4096;
  k_fill<<<(int)(b < 1 ? 1 : b), 256, 0, c->stream>>>(n, y, value);
  return post(c, "fill");
}

int update_psi(mgic_vars *v, const Geom &g, const BCk &bc, const double *dpsi) {
  SrcP s = make_srcp(v);
  dim3 blk(32, 4, 1), grd((s.nx + 2 + 31) / 32, (s.ny + 2 + 3) / 4, s.nz + 2);
  k_update_psi<<<grd, blk, 0, v->ctx->stream>>>(s, g, bc, v->d, dpsi);
  return post(v->ctx, "update_psi");
}

}  // namespace mgk
