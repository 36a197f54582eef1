// poisson_solver_b200.cpp -- the driver: reads a reference-format params.txt (plus key=value overrides), builds the
// base-level grids and runs the nonlinear loop of the reference (Main_PoissonSolver.cpp:45-256) against the B200
// operator through the reference's own interfaces: defineOperatorFactory -> MultilevelLinearOp::define ->
// BiCGStabSolver<Vector<LevelData<FArrayBox>*>>::solve -> set_update_psi0 / computeNorm.
//
//   poisson_solver_b200 params.txt [key=value ...] [--host-vcycle] [--dump-psi file] [--checkpoint file] [--json]
//
// max_level = 0: the host mirror classes, one device array per LevelData.  max_level > 0: set_grids (tagging + BRMeshRefine,
// mgic_grids_generate) and the nonlinear loop on the hierarchy through mgic_hier_* (each level's connected parts in masked
// arrays); --checkpoint writes the GRChombo checkpoint (MGICCHK1 container, tools/mgic2hdf5.py -> HDF5).
#include <chrono>
#include <cstring>

#include "MultilevelLinearOp.H"
#include "SetLevelDataDevice.H"

static int set_grids(Vector<DisjointBoxLayout> &a_grids, const PoissonParameters &a_params) {
  // Source/SetGrids.cpp:54-62: domainSplit base level; max_level = 0 stops there
  if (a_params.maxLevel != 0) MayDay::Error("B200 driver: the single-level path was asked for an AMR hierarchy");
  Vector<Box> boxes;
  domainSplit(a_params.coarsestDomain, boxes, a_params.maxGridSize, a_params.blockFactor);
  a_grids.assign(1, DisjointBoxLayout(boxes));
  return 0;
}

struct RunOptions { bool hostVcycle = false, json = false; std::string dumpPsi, checkpoint; };

// Main_PoissonSolver.cpp:259-293 + poissonSolve on a hierarchy (max_level > 0)
static int solveHierarchy(const PoissonParameters &a_params, const RunOptions &opt) {
  ParmParse pp;
  auto dev = std::make_shared<DeviceContext>(0);
  mgic_params P;
  fillDeviceParams(P, a_params, BCHolder::fromParmParse());
  P.max_level = a_params.maxLevel;
  int numMGIter = 1; pp.query("numMGIterations", numMGIter);
  int numMGSmooth = 4; pp.query("numMGsmooth", numMGSmooth);
  int preCondSolverDepth = -1; pp.query("preCondSolverDepth", preCondSolverDepth);
  Real tolerance = 1.0e-7; pp.query("tolerance", tolerance);
  int max_iter = 10; pp.query("max_iterations", max_iter);
  int max_NL_iter = 4; pp.query("max_NL_iterations", max_NL_iter);
  P.numMGsmooth = numMGSmooth; P.numMGIterations = numMGIter; P.preCondSolverDepth = preCondSolverDepth;
  P.tolerance = tolerance; P.max_iterations = max_iter; P.max_NL_iterations = max_NL_iter;
  if (a_params.periodic[0]) MayDay::Error("B200 driver: the periodic constant-K branch is out of scope (SURVEY.md 2.1 #6)");
  const auto t0 = std::chrono::steady_clock::now();
  mgic_grids *G = nullptr;
  mgic_hier *H = nullptr;
  MGIC_CALL(mgic_grids_generate(dev->ctx, &P, a_params.refineThresh, a_params.fillRatio, &G));      // set_grids
  for (int l = 0; l < mgic_grids_levels(G); l++) pout() << "set_grids: level " << l << ": " << mgic_grids_num_boxes(G, l) << " boxes" << std::endl;
  MGIC_CALL(mgic_hier_create_from_grids(dev->ctx, &P, G, &H));
  MGIC_CALL(mgic_hier_set_initial_conditions(H));                                                    // Main:93
  Real dpsi_norm = 0.0;
  std::vector<Real> norms;
  std::vector<int> iters;
  int status = 0;
  for (int NL_iter = 0; NL_iter < max_NL_iter; NL_iter++) {                                          // :131
    pout() << "Main Loop Iteration " << (NL_iter + 1) << " out of " << max_NL_iter << std::endl;
    MGIC_CALL(mgic_hier_set_sources(H, 0.0));                                                        // :154-160
    MGIC_CALL(mgic_hier_define_solver(H));                                                           // :163-170
    int it = 0;
    MGIC_CALL(mgic_hier_solve(H, &it, &status));                                                     // :184
    MGIC_CALL(mgic_hier_update_psi(H));                                                              // :189-205
    MGIC_CALL(mgic_hier_dpsi_norm(H, &dpsi_norm));                                                   // :208
    MGIC_CALL(mgic_hier_release_solver(H));
    iters.push_back(it); norms.push_back(dpsi_norm);
    pout() << "The norm of dpsi after step " << NL_iter + 1 << " is " << dpsi_norm << std::endl;
    if (dpsi_norm < tolerance || dpsi_norm > 1e5) break;                                             // :212
  }
  MGIC_CALL(mgic_ctx_sync(dev->ctx));
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  pout() << "The norm of dpsi at the final step was " << dpsi_norm << std::endl;
  if (dpsi_norm > 1e-1) MayDay::Error("NL iterations did not converge - may need a better initial guess");
  if (!opt.checkpoint.empty()) MGIC_CALL(mgic_hier_write_checkpoint(H, opt.checkpoint.c_str(), 0.0)); // output_final_data, :229
  if (!opt.dumpPsi.empty()) {
    int n[3];
    MGIC_CALL(mgic_hier_node_info(H, 0, nullptr, nullptr, n, nullptr));
    std::vector<Real> psi((size_t)n[0] * n[1] * n[2]);
    MGIC_CALL(mgic_hier_download(H, 0, 0, psi.data()));
    FILE *f = std::fopen(opt.dumpPsi.c_str(), "wb");
    if (!f) MayDay::Error("cannot open psi dump file");
    std::fwrite(psi.data(), sizeof(Real), psi.size(), f);
    std::fclose(f);
  }
  if (opt.json) {
    std::printf("{\"levels\": %d, \"nodes\": %d, \"nl_iterations\": %d, \"dpsi_norms\": [", mgic_grids_levels(G), mgic_hier_nodes(H), (int)norms.size());
    for (size_t i = 0; i < norms.size(); i++) std::printf("%s%.17g", i ? ", " : "", norms[i]);
    std::printf("], \"bicgstab_iterations\": [");
    for (size_t i = 0; i < iters.size(); i++) std::printf("%s%d", i ? ", " : "", iters[i]);
    std::printf("], \"exit_status\": %d, \"seconds\": %.6f, \"kernel_launches\": %lld}\n", status - 1, secs, mgic_ctx_launch_count(dev->ctx));
  }
  mgic_hier_destroy(H);
  mgic_grids_destroy(G);
  return status - 1;
}

static int poissonSolve(const Vector<DisjointBoxLayout> &a_grids, const PoissonParameters &a_params, const RunOptions &opt) {
  ParmParse pp;
  const int nlevels = a_params.numLevels;
  auto dev = std::make_shared<DeviceContext>(0);
  const BCHolder bc = BCHolder::fromParmParse();
  MultigridVarsDevice multigrid_vars(dev, a_params, bc);
  const IntVect ghosts = IntVect::Unit * 3;
  Vector<LevelData<FArrayBox> *> dpsi(nlevels, NULL), rhs(nlevels, NULL);
  Vector<RefCountedPtr<LevelData<FArrayBox>>> aCoef(nlevels), bCoef(nlevels);
  Vector<ProblemDomain> vectDomains(nlevels, a_params.coarsestDomain);
  Vector<RealVect> vectDx(nlevels, RealVect(a_params.coarsestDx, a_params.coarsestDx, a_params.coarsestDx));
  dpsi[0] = new LevelData<FArrayBox>(a_grids[0], 1, ghosts);
  rhs[0] = new LevelData<FArrayBox>(a_grids[0], 1, IntVect::Zero);
  aCoef[0] = RefCountedPtr<LevelData<FArrayBox>>(new LevelData<FArrayBox>(a_grids[0], 1, IntVect::Zero));
  bCoef[0] = RefCountedPtr<LevelData<FArrayBox>>(new LevelData<FArrayBox>(a_grids[0], 1, IntVect::Zero));
  set_initial_conditions(multigrid_vars, *dpsi[0], vectDx[0], a_params);

  const int lBase = 0;
  MultilevelLinearOp<FArrayBox> mlOp;
  BiCGStabSolver<Vector<LevelData<FArrayBox> *>> solver;
  // defaults and keys of Main_PoissonSolver.cpp:106-126
  int numMGIter = 1; pp.query("numMGIterations", numMGIter); mlOp.m_num_mg_iterations = numMGIter;
  int numMGSmooth = 4; pp.query("numMGsmooth", numMGSmooth); mlOp.m_num_mg_smooth = numMGSmooth;
  int preCondSolverDepth = -1; pp.query("preCondSolverDepth", preCondSolverDepth); mlOp.m_preCondSolverDepth = preCondSolverDepth;
  Real tolerance = 1.0e-7; pp.query("tolerance", tolerance);
  int max_iter = 10; pp.query("max_iterations", max_iter);
  int max_NL_iter = 4; pp.query("max_NL_iterations", max_NL_iter);
  mlOp.m_use_device_vcycle = !opt.hostVcycle;

  Real dpsi_norm = 0.0;
  const Real constant_K = 0.0;
  std::vector<Real> norms;
  std::vector<int> iters;
  const auto t0 = std::chrono::steady_clock::now();
  for (int NL_iter = 0; NL_iter < max_NL_iter; NL_iter++) {
    pout() << "Main Loop Iteration " << (NL_iter + 1) << " out of " << max_NL_iter << std::endl;
    if (a_params.periodic[0]) MayDay::Error("B200 driver: the periodic constant-K branch is out of scope (SURVEY.md 2.1 #6)");
    set_a_coef(*aCoef[0], multigrid_vars, a_params, vectDx[0], constant_K);
    set_b_coef(*bCoef[0], multigrid_vars, a_params, vectDx[0]);
    set_rhs(*rhs[0], multigrid_vars, vectDx[0], a_params, constant_K);
    // operators, MG hierarchy and coarse coefficients are rebuilt every nonlinear iteration, like the reference
    RefCountedPtr<AMRLevelOpFactory<LevelData<FArrayBox>>> opFactory(defineOperatorFactory(a_grids, vectDomains, aCoef, bCoef, a_params));
    static_cast<VariableCoeffPoissonOperatorFactory *>(opFactory.get())->setDevice(dev);
    mlOp.define(a_grids, a_params.refRatio, vectDomains, vectDx, opFactory, lBase);
    solver.define(&mlOp, false);
    solver.m_verbosity = a_params.verbosity;
    solver.m_normType = 0;
    solver.m_eps = tolerance;
    solver.m_imax = max_iter;
    solver.solve(dpsi, rhs);
    iters.push_back(solver.m_iterations);
    dpsi_norm = set_update_psi0(multigrid_vars, *dpsi[0]);  // psi += dpsi (ghosts included) and computeNorm(dpsi)
    norms.push_back(dpsi_norm);
    pout() << "The norm of dpsi after step " << NL_iter + 1 << " is " << dpsi_norm << std::endl;
    if (dpsi_norm < tolerance || dpsi_norm > 1e5) break;
  }
  MGIC_CALL(mgic_ctx_sync(dev->ctx));
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  pout() << "The norm of dpsi at the final step was " << dpsi_norm << std::endl;
  if (dpsi_norm > 1e-1) MayDay::Error("NL iterations did not converge - may need a better initial guess");
  if (!opt.dumpPsi.empty()) {
    std::vector<Real> psi;
    multigrid_vars.download(0, psi);
    FILE *f = std::fopen(opt.dumpPsi.c_str(), "wb");
    if (!f) MayDay::Error("cannot open psi dump file");
    std::fwrite(psi.data(), sizeof(Real), psi.size(), f);
    std::fclose(f);
  }
  if (opt.json) {
    std::printf("{\"nl_iterations\": %d, \"dpsi_norms\": [", (int)norms.size());
    for (size_t i = 0; i < norms.size(); i++) std::printf("%s%.17g", i ? ", " : "", norms[i]);
    std::printf("], \"bicgstab_iterations\": [");
    for (size_t i = 0; i < iters.size(); i++) std::printf("%s%d", i ? ", " : "", iters[i]);
    std::printf("], \"exit_status\": %d, \"seconds\": %.6f, \"kernel_launches\": %lld}\n", solver.m_exitStatus - 1, secs,
                mgic_ctx_launch_count(dev->ctx));
  }
  delete dpsi[0];
  delete rhs[0];
  return solver.m_exitStatus - 1;  // for AMRMultiGrid-style solvers success = 1 (Main_PoissonSolver.cpp:252-255)
}

int main(int argc, char *argv[]) {
  if (argc < 2) { std::cerr << " usage " << argv[0] << " <input_file_name> [key=value ...]" << std::endl; return 1; }
  RunOptions opt;
  std::vector<char *> overrides;
  for (int i = 2; i < argc; i++) {
    if (!std::strcmp(argv[i], "--host-vcycle")) opt.hostVcycle = true;
    else if (!std::strcmp(argv[i], "--json")) opt.json = true;
    else if (!std::strcmp(argv[i], "--dump-psi") && i + 1 < argc) opt.dumpPsi = argv[++i];
    else if (!std::strcmp(argv[i], "--checkpoint") && i + 1 < argc) opt.checkpoint = argv[++i];
    else overrides.push_back(argv[i]);
  }
  ParmParse pp((int)overrides.size(), overrides.data(), NULL, argv[1]);
  PoissonParameters params;
  Vector<DisjointBoxLayout> grids;
  getPoissonParameters(params);
  if (params.maxLevel > 0) return solveHierarchy(params, opt);
  set_grids(grids, params);
  return poissonSolve(grids, params, opt);
}
