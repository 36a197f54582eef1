// Test helper: the host mirror's getPoissonParameters (mg_ic_code_b200/host/PoissonParameters.H) on an input file +
// overrides, printed as JSON -- compared by tests/test_reference_pins.py with the reference's own getPoissonParameters.
#include <cstdio>

#include "PoissonParameters.H"

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  ParmParse pp(argc - 2, argv + 2, NULL, argv[1]);
  PoissonParameters p;
  getPoissonParameters(p);
  std::printf("\n{\"nCells\": [%d, %d, %d], \"maxGridSize\": %d, \"blockFactor\": %d, \"bufferSize\": %d, "
              "\"coefficient_average_type\": %d, \"verbosity\": %d, \"maxLevel\": %d, \"numLevels\": %d, \"refRatio0\": %d, "
              "\"refRatioLast\": %d, \"nRefRatio\": %d, \"periodic\": %d, \"fillRatio\": %.17g, \"refineThresh\": %.17g, "
              "\"coarsestDx\": %.17g, \"domainLength\": [%.17g, %.17g, %.17g], \"alpha\": %.17g, \"beta\": %.17g, "
              "\"G_Newton\": %.17g, \"phi_amplitude\": %.17g, \"phi_wavelength\": %.17g, \"bh1_bare_mass\": %.17g, "
              "\"bh2_bare_mass\": %.17g, \"bh1_spin\": %.17g, \"bh2_spin\": %.17g, \"bh1_momentum\": %.17g, \"bh2_momentum\": %.17g, "
              "\"bh1_offset\": %.17g, \"bh2_offset\": %.17g}\n",
              p.nCells[0], p.nCells[1], p.nCells[2], p.maxGridSize, p.blockFactor, p.bufferSize, p.coefficient_average_type,
              p.verbosity, p.maxLevel, p.numLevels, p.refRatio.front(), p.refRatio.back(), (int)p.refRatio.size(),
              (int)p.coarsestDomain.isPeriodic(0), p.fillRatio, p.refineThresh, p.coarsestDx, p.domainLength[0], p.domainLength[1],
              p.domainLength[2], p.alpha, p.beta, p.G_Newton, p.phi_amplitude, p.phi_wavelength, p.bh1_bare_mass, p.bh2_bare_mass,
              p.bh1_spin, p.bh2_spin, p.bh1_momentum, p.bh2_momentum, p.bh1_offset, p.bh2_offset);
  return 0;
}
