"""Grid / Coord / TagGrid helpers (coord.py:7-114, tag.py:36-78) as batched device integer
ops through ``pomdp_coord_op``, bit-exact against tables recorded from the unmodified
reference (tests/golden/coord.npz; includes the reference's own six asserts, coord.py:121-126).
"""
import numpy as np
import torch

from gym_pomdp_b200 import geometry as G

from backends import backend  # noqa: F401


def t(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev).to(torch.int32)


def test_reference_kats_and_moves(golden, backend):
    g = golden("coord")
    # Coord + Coord on the device == Coord + Moves for the four KATs that add a move (coord.py:123-126)
    moves = g["moves"]
    for m in range(5):
        base = t(g["kat_a"], backend)
        got = G.coord_add_move_batch(base, torch.full((len(base),), m, dtype=torch.int32, device=backend)).cpu().numpy()
        assert np.array_equal(got, g["kat_a"] + moves[m])
    kat_moves = {(0, 1): 0, (-1, 0): 3, (0, -1): 2, (1, 0): 1}
    for a, b, s in zip(g["kat_a"][2:], g["kat_b"][2:], g["kat_sum"][2:]):
        got = G.coord_add_move_batch(t([a], backend), t([kat_moves[tuple(b)]], backend)).cpu().numpy()[0]
        assert tuple(got) == tuple(s)
    assert [G.opposite(m) for m in range(4)] == g["opposite"].tolist()
    assert [tuple(G.Moves.get_coord(i)) for i in range(5)] == [tuple(m) for m in moves]
    # host mirror keeps the reference's Coord semantics
    assert G.Coord(3, 3) + G.Coord(2, 2) == G.Coord(5, 5) and G.Coord(5, 2) + G.Coord(2, 5) == G.Coord(7, 7)
    assert G.Coord(0, 0).is_valid() and not G.Coord(-1, 0).is_valid()


def test_grid_codec_and_bounds(golden, backend):
    g = golden("coord")
    for (xs, ys) in [(7, 7), (11, 11), (15, 15), (10, 10), (5, 5), (10, 5)]:
        coords = g[f"grid_{xs}x{ys}_coord"]
        idx = np.arange(len(coords))
        assert np.array_equal(G.grid_get_coord_batch(t(idx, backend), xs).cpu().numpy(), coords)
        assert np.array_equal(G.grid_get_index_batch(t(coords, backend), xs).cpu().numpy(), idx)
        probe = g[f"grid_{xs}x{ys}_probe"]
        assert np.array_equal(G.grid_is_inside_batch(t(probe, backend), xs, ys).cpu().numpy(), g[f"grid_{xs}x{ys}_inside"])
        grid = G.Grid(xs, ys)
        assert [tuple(grid.get_coord(i)) for i in idx] == [tuple(c) for c in coords]
        assert [grid.get_index(c) for c in coords] == idx.tolist()


def test_tag_grid(golden, backend):
    g = golden("coord")
    idx = np.arange(29)
    assert np.array_equal(G.tag_get_coord_batch(t(idx, backend)).cpu().numpy(), g["tag_coord"])
    assert np.array_equal(G.tag_get_index_batch(t(g["tag_coord"], backend)).cpu().numpy(), g["tag_index"])
    probe = g["tag_probe"]
    inside = G.tag_is_inside_batch(t(probe, backend)).cpu().numpy()
    assert np.array_equal(inside, g["tag_inside"])
    # off-board probes: -1 where the reference's asserts would fire
    got = G.tag_get_index_batch(t(probe, backend)).cpu().numpy()
    assert ((got >= 0) == g["tag_inside"]).all()
    assert (G.tag_get_coord_batch(t([29, 31, -1], backend)).cpu().numpy() == -1).all()
    tg = G.TagGrid()
    assert [tuple(tg.get_tag_coord(i)) for i in idx] == [tuple(c) for c in g["tag_coord"]]
    assert [bool(tg.is_corner(c)) for c in probe] == g["tag_corner"].tolist()


def test_l1_distance_is_what_the_reference_calls_euclidean(golden, backend):
    g = golden("coord")
    pts = g["dist_pts"]
    a = np.repeat(pts, len(pts), axis=0)
    b = np.tile(pts, (len(pts), 1))
    got = G.l1_distance_batch(t(a, backend), t(b, backend)).cpu().numpy().reshape(len(pts), len(pts))
    assert np.array_equal(got, g["dist_l1"].astype(np.int64))
    assert G.Grid.euclidean_distance((0, 0), (3, 4)) == 7.0 and G.Grid.manhattan_distance((0, 0), (3, 4)) == 5.0


def test_large_batch_roundtrip(backend):
    n = 100003
    rs = np.random.RandomState(0)
    idx = rs.randint(0, 225, n)
    c = G.grid_get_coord_batch(t(idx, backend), 15)
    assert np.array_equal(G.grid_get_index_batch(c, 15).cpu().numpy(), idx)
    assert G.grid_is_inside_batch(c, 15, 15).all()
