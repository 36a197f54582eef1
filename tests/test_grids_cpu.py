"""set_grids' clustering step -- [Chombo] BRMeshRefine::regrid restated in csrc/grids.cu (Berger-Rigoutsos) -- on synthetic
tags, host only (Grids.regrid touches no device).  Chombo's BRMeshRefine is not vendored with the reference, so there is no
upstream output to hold it to; what is checked are the properties Source/SetGrids.cpp relies on (:64-68, :113-114):
every tag that may be refined is covered, boxes are disjoint, multiples of block_factor, at most max_grid_size, properly
nested with radius 2, filled to at least fill_ratio unless they cannot be cut further, and the result is deterministic."""
import itertools

import numpy as np
import pytest

import mg_ic_code_b200 as m

P = dict(m.DEFAULTS, N=(32, 32, 32), L=100.0, block_factor=8, max_grid_size=16, max_level=3)


def lattice(n, size):
    return [((i, j, k), (i + size - 1, j + size - 1, k + size - 1)) for k in range(0, n, size) for j in range(0, n, size) for i in range(0, n, size)]


def raster(boxes, n):
    a = np.zeros((n, n, n), dtype=np.int32)
    for lo, hi in boxes:
        a[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1] += 1
    return a


def ball(n, centre, radius):
    k, j, i = np.mgrid[0:n, 0:n, 0:n]
    return (i - centre[0]) ** 2 + (j - centre[1]) ** 2 + (k - centre[2]) ** 2 <= radius ** 2


def points(mask):
    k, j, i = np.nonzero(mask)
    return list(zip(i.tolist(), j.tolist(), k.tolist()))


def check_level(boxes, n_fine, block, max_size):
    r = raster(boxes, n_fine)
    assert r.max() <= 1, "boxes overlap"
    for lo, hi in boxes:
        for d in range(3):
            assert lo[d] % block == 0 and (hi[d] + 1) % block == 0 and hi[d] - lo[d] + 1 <= max_size
    return r.astype(bool)


def test_two_blobs_give_two_parts_that_cover_their_tags():
    tags = ball(32, (7, 16, 16), 2.5) | ball(32, (25, 16, 16), 2.5)
    # (with fill_ratio 0.5 the bounding box of both blobs, 16 of 24 blocks tagged, is accepted as it is: one part)
    g = m.Grids.regrid(P, [lattice(32, 16)], [points(tags)], fill_ratio=0.5)
    assert len(g.nodes(1)) == 1
    g.close()
    g = m.Grids.regrid(P, [lattice(32, 16)], [points(tags)], fill_ratio=0.75)
    assert g.levels == 2
    boxes, part, nparts = g.boxes(1, parts=True)
    fine = check_level(boxes, 64, 8, 16)
    assert nparts == 2 and len(g.nodes(1)) == 2
    covered = fine.reshape(32, 2, 32, 2, 32, 2).all(axis=(1, 3, 5))       # level-0 cells under level 1
    assert np.all(covered[tags]), "a tagged cell is not refined"
    assert covered.sum() < 0.2 * 32 ** 3                                   # and it did not refine everything
    # deterministic
    g2 = m.Grids.regrid(P, [lattice(32, 16)], [points(tags)], fill_ratio=0.75)
    assert g2.boxes(1) == boxes
    g.close(); g2.close()


@pytest.mark.parametrize("fill", [0.3, 0.5, 0.8, 1.0])
def test_fill_ratio_is_honoured(fill):
    """an L-shaped tag set: with fill_ratio 1 only tagged blocks are refined; lower ratios may fill the corner in"""
    tags = np.zeros((32, 32, 32), dtype=bool)
    tags[8:24, 8:12, 12:20] = True
    tags[8:12, 8:24, 12:20] = True
    g = m.Grids.regrid(P, [lattice(32, 16)], [points(tags)], fill_ratio=fill)
    fine = check_level(g.boxes(1), 64, 8, 16)
    covered = fine.reshape(32, 2, 32, 2, 32, 2).all(axis=(1, 3, 5))
    assert np.all(covered[tags])
    blocks_tagged = tags.reshape(8, 4, 8, 4, 8, 4).any(axis=(1, 3, 5))    # a level-1 block of 8 cells = 4 level-0 cells
    blocks_cov = covered.reshape(8, 4, 8, 4, 8, 4).all(axis=(1, 3, 5))
    assert np.all(blocks_cov[blocks_tagged])
    if fill == 1.0:
        assert np.array_equal(blocks_cov, blocks_tagged)
    assert blocks_tagged.sum() >= fill * blocks_cov.sum() - 1e-9 or fill < 1.0
    g.close()


def test_tags_next_to_the_level_boundary_are_clipped_to_the_nesting_domain():
    """level 1 = one 32^3 box in a 64^3 domain; tags on level 1 everywhere: level 2 stays two level-1 cells (rounded up to a
    block of 4) inside level 1, except where level 1 touches the domain boundary"""
    lvl1 = [((0, 16, 16), (31, 47, 47))]
    t1 = np.zeros((64, 64, 64), dtype=bool)
    t1[16:48, 16:48, 0:32] = True
    t0 = np.zeros((32, 32, 32), dtype=bool)
    t0[8:24, 8:24, 0:16] = True
    g = m.Grids.regrid(P, [lattice(32, 16), lvl1], [points(t0), points(t1)], fill_ratio=0.5)
    assert g.levels == 3
    l1 = check_level(g.boxes(1), 64, 8, 16)
    l2 = check_level(g.boxes(2), 128, 8, 16)
    assert np.array_equal(l1, t1)                                          # level 1 reproduced from level 0's tags
    under = l2.reshape(64, 2, 64, 2, 64, 2).any(axis=(1, 3, 5))           # level-1 cells under level 2
    assert under.any()
    k, j, i = np.nonzero(under)
    assert i.min() == 0                                                     # reaches the domain face x = 0 ...
    assert i.max() <= 31 - 4 and j.min() >= 16 + 4 and j.max() <= 47 - 4 and k.min() >= 16 + 4 and k.max() <= 47 - 4   # ... but keeps a block off level 1's other faces
    # proper nesting, radius 2: every level-1 cell within 2 of a cell under level 2 is a level-1 cell (or outside the domain)
    for di, dj, dk in itertools.product((-2, 0, 2), repeat=3):
        ii, jj, kk = i + di, j + dj, k + dk
        inside = (ii >= 0) & (jj >= 0) & (kk >= 0) & (ii < 64) & (jj < 64) & (kk < 64)
        assert np.all(l1[kk[inside], jj[inside], ii[inside]])
    g.close()


def test_nesting_domain_of_a_level_that_is_not_a_box():
    """level 1 = an L-shaped union of boxes that touches two domain faces, every level-1 cell tagged, fill ratio 1: level 2 must
    cover exactly the blocks of 4 level-1 cells that lie wholly in level 1 eroded by the 5^3 cube (nesting radius 2, beyond the
    domain boundary counting as inside) -- the erosion is done axis by axis in grids.cu, here by scipy in one go"""
    from scipy import ndimage
    lvl1 = [((0, 0, 16), (31, 15, 47)), ((0, 16, 16), (15, 47, 47)), ((16, 16, 32), (31, 31, 47)), ((32, 48, 0), (63, 63, 15))]
    l1 = raster(lvl1, 64).astype(bool)
    t0 = l1.reshape(32, 2, 32, 2, 32, 2).any(axis=(1, 3, 5))
    g = m.Grids.regrid(P, [lattice(32, 16), lvl1], [points(t0), points(l1)], fill_ratio=1.0)
    assert g.levels == 3
    assert np.array_equal(check_level(g.boxes(1), 64, 8, 16), l1)          # level 1 reproduced (its tags are its own cells)
    l2 = check_level(g.boxes(2), 128, 8, 16)
    under = l2.reshape(64, 2, 64, 2, 64, 2).any(axis=(1, 3, 5))
    eroded = ndimage.binary_erosion(l1, structure=np.ones((5, 5, 5), dtype=bool), border_value=1)
    blocks = eroded.reshape(16, 4, 16, 4, 16, 4).all(axis=(1, 3, 5))
    expect = np.repeat(np.repeat(np.repeat(blocks, 4, axis=0), 4, axis=1), 4, axis=2)
    assert expect.any() and np.array_equal(under, expect)
    g.close()


def test_finer_level_forces_its_coarser_level_to_hold_it():
    """tags only on level 1 (none on level 0 around them) still keep the new level 1 under the new level 2"""
    lvl1 = [((16, 16, 16), (47, 47, 47))]
    t0 = np.zeros((32, 32, 32), dtype=bool)
    t0[12:14, 12:14, 12:14] = True                                          # a corner of what level 2 will need
    t1 = np.zeros((64, 64, 64), dtype=bool)
    t1[36:40, 36:40, 36:40] = True
    g = m.Grids.regrid(P, [lattice(32, 16), lvl1], [points(t0), points(t1)], fill_ratio=0.5)
    assert g.levels == 3
    l1 = check_level(g.boxes(1), 64, 8, 16)
    l2 = check_level(g.boxes(2), 128, 8, 16)
    under = l2.reshape(64, 2, 64, 2, 64, 2).any(axis=(1, 3, 5))
    assert np.all(under[t1])
    k, j, i = np.nonzero(under)
    for di, dj, dk in itertools.product((-2, 0, 2), repeat=3):
        assert np.all(l1[k + dk, j + dj, i + di]), "level 2 is not nested in the new level 1"
    g.close()


def test_bad_parameters_are_refused():
    with pytest.raises(m.MgicError):
        m.Grids.regrid(dict(P, block_factor=6), [lattice(32, 16)], [[(1, 1, 1)]])
    with pytest.raises(m.MgicError):
        m.Grids.regrid(dict(P, max_grid_size=12), [lattice(32, 16)], [[(1, 1, 1)]])
    with pytest.raises(m.MgicError):
        m.Grids.regrid(P, [lattice(32, 16)], [[(1, 1, 1)]], fill_ratio=0.0)
