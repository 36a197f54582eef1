"""set_grids on the GPU path (Source/SetGrids.cpp:31-207): the regrid condition kernel against the oracle (which is pinned
bit for bit to the reference's own set_regrid_condition, tests/test_reference_pins.py), the tagging / regrid loop against an
independent loop written here over the oracle's condition, and the reference's whole run -- set_grids, then poissonSolve on
the hierarchy it produced (Main_PoissonSolver.cpp:278-284) -- at a size the checks finish in seconds."""
import numpy as np
import pytest

import mg_ic_code_b200 as m
from oracle import condition_box, default_params

pytestmark = pytest.mark.gpu


def lattice(N, size):
    return [((i, j, k), (min(i + size, N[0]) - 1, min(j + size, N[1]) - 1, min(k + size, N[2]) - 1))
            for k in range(0, N[2], size) for j in range(0, N[1], size) for i in range(0, N[0], size)]


def set_grids_twin(params, thresh, fill, grow=2):
    """Source/SetGrids.cpp:70-136 over the oracle's set_regrid_condition; the clustering itself is the library's (host only)"""
    N, L, max_level = params["N"], params["L"], params["max_level"]
    boxes = [lattice(N, params["max_grid_size"])]
    top, more = 0, max_level > 0
    while more:
        more, old = False, top
        tags = []
        for l in range(top + 1):
            n = [x << l for x in N]
            lo = [min(b[0][d] for b in boxes[l]) for d in range(3)]
            hi = [max(b[1][d] for b in boxes[l]) for d in range(3)]
            cond = np.abs(condition_box(params, L / N[0] / (1 << l), lo, hi, 0))
            union = np.zeros(cond.shape, dtype=bool)
            for a, b in boxes[l]:
                union[a[2] - lo[2]:b[2] - lo[2] + 1, a[1] - lo[1]:b[1] - lo[1] + 1, a[0] - lo[0]:b[0] - lo[0] + 1] = True
            t = (cond >= thresh * cond[union].max()) & union
            pts = set()
            for k, j, i in zip(*np.nonzero(t)):
                for dk in range(-grow, grow + 1):
                    for dj in range(-grow, grow + 1):
                        for di in range(-grow, grow + 1):
                            q = (int(i) + lo[0] + di, int(j) + lo[1] + dj, int(k) + lo[2] + dk)
                            if all(0 <= q[d] < n[d] for d in range(3)):
                                pts.add(q)
            tags.append(sorted(pts))
        g = m.Grids.regrid(dict(params, max_level=min(top + 1, max_level)), boxes, tags, fill)
        new = [g.boxes(l) for l in range(1, g.levels)]
        g.close()
        if len(new) > top and new[top]:
            top += 1
        boxes = [boxes[0]] + new[:top]
        more = top < max_level and top > old
    return boxes


CASE = dict(N=(32, 32, 32), L=100.0, max_grid_size=16, block_factor=8, max_level=2, numMGsmooth=2, max_NL_iterations=4)


def test_regrid_condition_kernel_against_the_oracle(ctx):
    """k_condition through mgic_grids_generate's level statistics is indirect; here the kernel's values are compared cell by
    cell on a refined level's box, via a one-level hierarchy whose tags are everything (threshold 0)"""
    P = default_params(**dict(CASE, max_level=1))
    g = m.Grids.generate(ctx, m.make_params(P), refine_threshold=0.0, fill_ratio=0.5)
    assert g.levels == 2
    st = g.stats(0)
    cond = np.abs(condition_box(P, P["L"] / 32, (0, 0, 0), (31, 31, 31), 0))
    assert abs(st["max_condition"] - cond.max()) <= 1e-13 * cond.max()
    assert st["tagged_cells"] == 32 ** 3 and st["cells"] == 32 ** 3
    # threshold 0 tags everything; the nesting domain of the base level is the whole domain: level 1 covers it
    assert g.stats(1)["cells"] == 64 ** 3
    g.close()


@pytest.mark.parametrize("over", [dict(), dict(refine=0.3), dict(N=(32, 32, 64), fill=0.7)])
def test_set_grids_loop_matches_the_twin(ctx, over):
    over = dict(over)
    thresh, fill = over.pop("refine", 0.1), over.pop("fill", 0.5)
    P = default_params(**dict(CASE, **over))
    g = m.Grids.generate(ctx, m.make_params(P), refine_threshold=thresh, fill_ratio=fill)
    want = set_grids_twin(P, thresh, fill)
    assert g.levels == len(want) >= 2
    for l in range(g.levels):
        assert g.boxes(l) == want[l], l
    # the hierarchy is properly nested and refines around the punctures (x = +-10, y = z = 0)
    n1 = [x * 2 for x in P["N"]]
    dx1 = P["L"] / P["N"][0] / 2
    for bh in (10.0, -10.0):
        c = [int((bh + P["L"] / 2) / dx1)] + [int((dx1 * n1[d]) / 2 / dx1) for d in (1, 2)]
        assert any(all(lo[d] <= c[d] <= hi[d] for d in range(3)) for lo, hi in g.boxes(1)), "the puncture is not refined"
    g.close()


def test_reference_run_set_grids_then_solve(ctx):
    """Main_PoissonSolver.cpp:278-284 with max_level = 2: set_grids, then the nonlinear solve on the generated hierarchy
    converges like the single-level run and psi agrees with the single-level solve where both exist"""
    P = default_params(**CASE)
    mp = m.make_params(P)
    g = m.Grids.generate(ctx, mp, 0.1, 0.5)
    H = m.Hierarchy.from_grids(ctx, mp, g)
    assert H.nodes == 1 + sum(len(g.nodes(l)) for l in range(1, g.levels))
    norms = H.nl_solve()
    assert len(norms) >= 3 and norms[0] > 1e-2 and norms[-1] < 1e-7 and all(b < 0.05 * a for a, b in zip(norms, norms[1:])), norms
    one, psi1 = m.nl_solve(ctx, m.make_params(dict(P, max_level=0)))
    psi = H.download(0, "psi")
    assert np.abs(psi - psi1).max() < 5e-3                                  # refinement changes the coarse solution at discretisation level only
    lvl, lo, n, cells = H.node_info(1)
    assert lvl == 1 and cells == int(H.mask(1).sum())
    fine = H.download(1, "psi")
    assert np.all(fine[H.mask(1) == 0] == 0) and np.abs(fine[H.mask(1) == 1] - 1).max() < 0.05
    H.close(); g.close()


@pytest.mark.parametrize("max_level", [2, 3])
def test_solve_on_the_generated_hierarchy_matches_the_twin(ctx, max_level):
    """set_grids' own hierarchy (BRMeshRefine's boxes, every connected part of a level one masked array) through three nonlinear
    iterations against the oracle-backed twin on the same boxes: BiCGStab counts, dpsi norms, psi on every node to 1e-10"""
    from amr_twin import hierarchy_nl_solve
    from test_amr_hierarchy import bare_hierarchy
    P = default_params(**dict(CASE, max_level=max_level))
    mp = m.make_params(dict(P, max_NL_iterations=3))
    g = m.Grids.generate(ctx, mp, 0.1, 0.5)
    assert g.levels == max_level + 1
    boxes = {l: [[(tuple(lo), tuple(hi)) for lo, hi in part] for part in g.nodes(l)] for l in range(1, g.levels)}
    assert max(len(part) for l in boxes for part in boxes[l]) > 1, "the case is meant to have parts made of several touching boxes"
    o, patches = bare_hierarchy(boxes, N=32, L=P["L"], box=16)
    log = []
    norms_o, psi_o = hierarchy_nl_solve(o, patches, max_nl=3, log=log)
    H = m.Hierarchy.from_grids(ctx, mp, g)
    H.set_initial_conditions()
    got = []
    for it in range(3):
        nrm, its, st = H.nl_iteration()
        got.append(nrm)
        assert (its, st) == (log[it][0], log[it][1]), (it, its, st, log[it])
    assert np.allclose(got, norms_o, rtol=1e-7), (got, norms_o)
    for q in range(H.nodes):
        d = np.abs(H.download(q, "psi") - psi_o[q]).max() / np.abs(psi_o[q]).max()
        assert d < 1e-10, (q, d)
    H.close(); g.close(); o.close()


def test_grchombo_checkpoint(ctx, tmp_path):
    """output_final_data (Source/WriteOutput.H:127-227): header and per-level attributes as the reference sets them, the level's
    boxes, and per box the 32 GRChombo variables with three ghost layers -- set_output_data, whose oracle restatement is
    pinned bit for bit to the reference's (tests/test_reference_pins.py) -- from the solved hierarchy."""
    from mg_ic_code_b200 import checkpoint
    from oracle import output_box
    P = default_params(**dict(CASE, max_NL_iterations=2))
    mp = m.make_params(P)
    g = m.Grids.generate(ctx, mp, 0.1, 0.5)
    H = m.Hierarchy.from_grids(ctx, mp, g)
    H.nl_solve()
    path = tmp_path / "vcPoissonFinal.3d.mgic"
    H.write_checkpoint(path, constant_K=-0.25)
    hdr, levels = checkpoint.read(path)
    root = hdr["root"]
    assert root["ints"]["max_level"] == g.levels - 1 and root["ints"]["num_levels"] == g.levels and root["ints"]["num_components"] == 32
    assert root["ints"]["iteration"] == 0 and root["reals"]["time"] == 0.0 and root["ints"]["regrid_interval_1"] == 1
    assert root["strings"]["component_0"] == "chi" and root["strings"]["component_25"] == "phi" and root["strings"]["component_30"] == "Mom3"
    dx0 = P["L"] / P["N"][0]
    for l, lv in enumerate(hdr["levels"]):
        assert lv["group"] == f"level_{l}" and lv["ints"]["ref_ratio"] == 2 and lv["ints"]["tag_buffer_size"] == 3
        assert lv["reals"]["dx"] == dx0 / 2 ** l and lv["reals"]["dt"] == 0.25 * dx0 / 2 ** l and lv["ints"]["is_periodic_2"] == 1
        assert lv["prob_domain"] == [0, 0, 0] + [(x << l) - 1 for x in P["N"]]
        assert sorted(map(tuple, lv["boxes"])) == sorted(lo + hi for lo, hi in g.boxes(l))
        assert len(lv["offsets"]) == len(lv["boxes"]) + 1
    # data: per box, psi of every cell reconstructed from the solver's own arrays -> the oracle's set_output_data
    psi_nodes = [(H.node_info(q), H.download(q, "psi"), H.mask(q)) for q in range(H.nodes)]
    checked = 0
    for l, (lv, data) in enumerate(zip(hdr["levels"], levels)):
        for b, arr in zip(lv["boxes"], data):
            lo, hi = b[:3], b[3:]
            node = next(q for q, ((lvl, nlo, nn, _), _, _) in enumerate(psi_nodes)
                        if lvl == l and all(nlo[d] <= lo[d] and hi[d] < nlo[d] + nn[d] for d in range(3)))
            (_, nlo, nn, _), psi, mask = psi_nodes[node]
            sl = tuple(slice(lo[d] - nlo[d], hi[d] - nlo[d] + 1) for d in (2, 1, 0))
            assert mask[sl].all()
            # valid cells: chi, A_ij, phi, K, the constants
            want = output_box(P, dx0 / 2 ** l, lo, hi, psi[sl], -0.25)
            got = arr[:, 3:-3, 3:-3, 3:-3]
            for v in range(32):
                s = np.abs(want[v]).max()
                assert np.abs(got[v] - want[v]).max() <= 1e-12 * max(s, 1e-300), (l, b, v)
            # ghost cells never written by the reference keep psi = 1: the outermost layer away from other boxes of the level
            ghosted_lo, ghosted_hi = [x - 3 for x in lo], [x + 3 for x in hi]
            ones = output_box(P, dx0 / 2 ** l, ghosted_lo, ghosted_hi, np.ones(arr.shape[1:]), -0.25)
            corner = (slice(None), 0, 0, 0)
            if not any(0 <= ghosted_lo[d] - nlo[d] < nn[d] for d in range(3)):     # the corner cell lies outside the node's array
                assert abs(arr[corner][0] - ones[corner][0]) <= 1e-12 * ones[corner][0]
            assert np.all(arr[1] == 1.0) and np.all(arr[18] == 1.0) and np.all(arr[7] == -0.25) and not arr[14].any()
            checked += 1
    assert checked == sum(len(g.boxes(l)) for l in range(g.levels))
    H.close(); g.close()
