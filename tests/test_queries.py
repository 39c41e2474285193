"""Batched ``_compute_prob`` (particle reweighting) and ``_generate_legal`` masks as kernels
(SURVEY.md §8f ranks 2-3), against the oracle's restatements of rock.py:250-264 / 273-291,
tag.py:209-217, battleship.py:80-89 / 157-165, tiger.py:125-138, network.py:43-55 / 129-130.
(The reference-recorded fixtures pin the same functions in tests/test_parity_golden.py.)
"""
import numpy as np
import torch

import gym_pomdp_b200 as gp
from oracle import pomdp_oracle as O

from backends import backend  # noqa: F401


def test_rock_obs_prob_and_masks_vs_oracle(backend):
    for board, k in [(7, 8), (11, 11), (15, 15)]:
        N = 3000
        env = gp.make("Rock-v0", board_size=board, num_rocks=k, batch_size=N, device=backend, seed=1)
        cfg = O.RockCfg(board, k)
        rs = np.random.RandomState(board)
        x, y, status = rs.randint(0, board, N), rs.randint(0, board, N), rs.randint(-1, 2, (N, k))
        action, ob = rs.randint(0, 5 + k, N), rs.randint(0, 3, N)
        state = env.pack(x, y, status)
        p = env._compute_prob(torch.as_tensor(action), state, torch.as_tensor(ob)).cpu().numpy()
        exp = np.array([O.rock_compute_prob(cfg, int(action[i]), int(x[i]), int(y[i]), status[i].tolist(), int(ob[i])) for i in range(N)],
                       dtype=np.float64)
        assert np.array_equal(p, exp)                          # doubles, bit for bit (eff and 1 - eff)
        mask = env._generate_legal(state).cpu().numpy()
        for i in range(0, N, 3):
            if board == 15 and (x[i], y[i]) == (12, 2):
                continue
            assert np.nonzero(mask[i])[0].tolist() == sorted(set(O.rock_generate_legal(cfg, int(x[i]), int(y[i]), status[i].tolist())))
        # done states: the list is a function of (agent, rocks) only
        done_state = env.pack(x, y, status, done=np.ones(N, int))
        assert torch.equal(env._generate_legal(done_state).cpu(), torch.as_tensor(mask))


def test_network_masks_span_two_words_and_probs(backend):
    n = 19                                                       # 39 actions -> two mask words
    env = gp.make("Network-v0", n_machines=n, batch_size=500, device=backend, seed=1)
    rs = np.random.RandomState(2)
    state = torch.as_tensor(rs.randint(0, 1 << n, 500), device=backend).int()
    assert env.legal_mask_words(state).shape == (500, 2)
    assert env._generate_legal(state).all() and env._generate_legal(state).shape == (500, 39)
    action, ob = rs.randint(0, 39, 500), rs.randint(0, 3, 500)
    p = env._compute_prob(torch.as_tensor(action), state, torch.as_tensor(ob)).cpu().numpy()
    s = state.cpu().numpy()
    exp = [O.network_compute_prob(int(action[i]), [(int(s[i]) >> m) & 1 for m in range(n)], int(ob[i])) for i in range(500)]
    assert np.array_equal(p, np.array(exp, dtype=np.float64))
    assert set(np.unique(p).tolist()) <= {0.0, 1.0, 0.95, 1 - 0.95}


def test_tiger_tag_battleship_probs(backend):
    env = gp.make("Tiger-v0", batch_size=18, device=backend, seed=1)
    st = env.pack([s for s in (0, 1) for _ in range(9)])
    a = torch.as_tensor([a for _ in (0, 1) for a in range(3) for _ in range(3)])
    o = torch.as_tensor([o for _ in range(6) for o in range(3)])
    for cp in (.85, .7):
        p = env._compute_prob(a, st, o, cp).cpu().numpy()
        exp = [O.tiger_compute_prob(int(a[i]), int(i >= 9), int(o[i]), cp) for i in range(18)]
        assert np.array_equal(p, np.array(exp))
    assert env._generate_legal(st).all()
    tag = gp.make("Tag-v0", num_opponents=2, batch_size=900, device=backend, seed=1)
    rs = np.random.RandomState(3)
    agent, opp, ob = rs.randint(0, 29, 900), rs.randint(0, 29, (900, 2)), rs.randint(0, 30, 900)
    opp[:200, 1] = agent[:200]
    ob[:100] = 29
    p = tag._compute_prob(torch.zeros(900), tag.pack(agent, opp), torch.as_tensor(ob)).cpu().numpy()
    exp = [float(O.tag_compute_prob(*O.tag_get_coord(int(agent[i])), [O.tag_get_coord(int(c)) for c in opp[i]], int(ob[i])))
           for i in range(900)]
    assert np.array_equal(p, np.array(exp))
    ship = gp.make("Battleship-v0", board_size=(10, 10), batch_size=256, device=backend, seed=1)
    ship.reset()
    for a in (3, 17, 55):
        ship.step(torch.full((256,), a, dtype=torch.int32))
    legal = ship._generate_legal(ship.state)
    assert legal.shape == (256, 100) and (legal.sum(1) == 97).all() and not legal[:, [3, 17, 55]].any()
    occ, vis, _, _ = ship.unpack(ship.state)
    for a in (3, 4, 55, 99):
        for o in (0, 1):
            p = ship._compute_prob(torch.full((256,), a), ship.state, torch.full((256,), o)).cpu().numpy()
            x, y = a % 10, a // 10
            exp = ((o == 0) & vis[:, x, y].cpu().numpy()) | ((o == 1) & occ[:, x, y].cpu().numpy()) | (o == 0)
            assert np.array_equal(p, exp.astype(np.float64))
