"""CPU restatement (pure Python, scalar) of gym_pomdp's step()/reset() generative models.

TEST INFRASTRUCTURE ONLY.  Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg may import this file; it is the CHECKER, never
the product path (gym_pomdp_b200/ must fail loudly when its CUDA library is missing).

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the unmodified reference
(/root/reference, imported through oracle/ref_shim.py) with its numpy draws scripted from
Philox words and writes tests/golden/*.npz; tests/test_oracle_golden.py checks every
function below (and the C restatement oracle/pomdp_oracle.c) against those fixtures, and
tests/test_oracle_vs_reference.py re-checks against the live reference when
/root/reference is present.

Conventions
-----------
* States are plain ints / lists in the reference's own units (coords, statuses in
  {-1,0,+1}, 0/1 machine flags) -- NOT the packed device words; packing lives in
  gym_pomdp_b200/ and is tested against these through the codec.
* ``draw(slot)`` returns the uint32 Philox word of that draw slot for this (env, step)
  (oracle/philox.py).  Slots are FIXED per decision, not consumed sequentially, so the
  kernels have uniform control flow; oracle/gen_golden.py maps the reference's sequential
  consumption onto slots.  Coupling rules are in oracle/ref_shim.py.
* Every function cites the reference lines it restates (paths under
  /root/reference/gym_pomdp/envs/).
"""
import math

TWO32 = 1 << 32

# ---------------------------------------------------------------------------------------
# draw -> decision rules (oracle/ref_shim.py docstring)
# ---------------------------------------------------------------------------------------


def bern_threshold(p):
    """T such that  (r < T)  <=>  (r / 2**32 < p)  for every uint32 r:  T = ceil(p * 2**32)."""
    return max(0, min(TWO32, math.ceil(p * TWO32)))


def gt_threshold(p):
    """G such that  (r > G)  <=>  (r / 2**32 > p):  G = floor(p * 2**32)."""
    return math.floor(p * TWO32)


def rand_below(r, n):
    """np.random.randint(n) / choice index under the coupling rule."""
    return (r * n) >> 32


# ---------------------------------------------------------------------------------------
# Geometry (coord.py:7-114)
# ---------------------------------------------------------------------------------------
# Moves enum order, coord.py:101-106: 0 N(0,+1) 1 E(+1,0) 2 S(0,-1) 3 W(-1,0) 4 NULL
MOVES = ((0, 1), (1, 0), (0, -1), (-1, 0), (0, 0))
# Compass enum order, battleship.py:12-21
COMPASS = ((0, 1), (1, 0), (0, -1), (-1, 0), (0, 0), (1, 1), (1, -1), (-1, -1), (-1, 1))


def grid_get_index(x_size, x, y):
    """coord.py:58-59"""
    return x_size * y + x


def grid_get_coord(x_size, idx):
    """coord.py:64-66"""
    return idx % x_size, idx // x_size


def grid_is_inside(x_size, y_size, x, y):
    """coord.py:18-19, 61-62"""
    return x >= 0 and y >= 0 and x < x_size and y < y_size


def grid_opposite(move):
    """coord.py:75-77"""
    return (move + 2) % 4


def l1_distance(x0, y0, x1, y1):
    """coord.py:79-81: ``euclidean_distance`` is np.linalg.norm(., 1) == the L1 norm."""
    return abs(x0 - x1) + abs(y0 - y1)


# ---------------------------------------------------------------------------------------
# RockSample (rock.py)
# ---------------------------------------------------------------------------------------
# rock.py:43-64 (constants of the benchmark; n=15 lists 16 rocks with [1,2] twice)
ROCK_CONFIG = {
    2: ((2, 1), (0, 0), ((1, 0),)),
    4: ((4, 3), (0, 0), ((1, 0), (3, 1), (2, 3))),
    7: ((7, 8), (0, 3), ((2, 0), (0, 1), (3, 1), (6, 3), (2, 4), (3, 4), (5, 5), (1, 6))),
    11: ((11, 11), (0, 5), ((0, 3), (0, 7), (1, 8), (2, 4), (3, 3), (3, 8), (4, 3), (5, 8),
                            (6, 1), (9, 3), (9, 9))),
    15: ((15, 15), (0, 5), ((0, 7), (0, 3), (1, 2), (1, 2), (2, 6), (3, 7), (3, 2), (4, 7),
                            (5, 2), (6, 9), (9, 7), (9, 1), (11, 8), (13, 10), (14, 9),
                            (12, 2))),
}


class RockCfg:
    """rock.py:99-118 (RockEnv.__init__) and rock.py:429-432 (StochasticRockEnv)."""

    def __init__(self, board_size=7, num_rocks=8, stochastic=False, p_move=0.8):
        sizes, start, rocks = ROCK_CONFIG[board_size]
        assert num_rocks in sizes  # rock.py:101
        self.n = board_size
        self.k = num_rocks
        self.start = start
        self.rock_pos = rocks  # ALL listed rocks are written to the grid (rock.py:110-111)
        self.grid = [[-1] * board_size for _ in range(board_size)]  # [x][y], coord.py:51-52
        for idx, (rx, ry) in enumerate(rocks):
            self.grid[rx][ry] = idx  # later ids overwrite earlier ones
        self.stochastic = stochastic
        self.p_move = p_move
        self.penalization = 0 if stochastic else -100  # rock.py:117, 432
        self.n_actions = 5 + num_rocks  # rock.py:113
        self.max_dist = 2 * (board_size - 1)

    def efficiency(self, d):
        """rock.py:383-387"""
        return (1 + pow(2, -d / 20)) * .5


ROCK_ERR_DANGLING = 8  # sampled a grid id >= num_rocks (reference raises IndexError)


def rock_step(cfg, x, y, status, action, draw):
    """rock.py:123-194 / rock.py:434-504.  Returns (x, y, status, obs, reward, done, err).

    Draw slots: 0 = p_move gate (StochasticRock only, rock.py:443); 1 = sensor
    (rock.py:404).  ``status`` is a list of k ints in {-1,0,+1}; a new list is returned.
    """
    status = list(status)
    reward, ob, err = 0, 0, 0
    n = cfg.n
    if cfg.stochastic and not (draw(0) < bern_threshold(cfg.p_move)):
        return x, y, status, ob, reward, False, err  # rock.py:443, 504
    if action < 4:
        if action == 1:  # EAST, rock.py:135-141
            if x + 1 < n:
                x += 1
            else:
                return x, y, status, ob, 10, True, err
        elif action == 0:  # NORTH
            if y + 1 < n:
                y += 1
            else:
                reward = cfg.penalization
        elif action == 2:  # SOUTH
            if y - 1 >= 0:
                y -= 1
            else:
                reward = cfg.penalization
        else:  # WEST
            if x - 1 >= 0:
                x -= 1
            else:
                reward = cfg.penalization
    elif action == 4:  # SAMPLE, rock.py:160-169
        rock = cfg.grid[x][y]
        if rock >= cfg.k:
            # reference: IndexError at rock.py:162 (Rock(15,15) at (12,2); Rock(7,7) at (1,6)).
            # Defined behaviour here: flag, and treat the cell as holding no rock.
            err |= ROCK_ERR_DANGLING
            rock = -1
        if rock >= 0 and status[rock] != 0:
            reward = 10 if status[rock] == 1 else -10
            status[rock] = 0
        else:
            reward = cfg.penalization
    else:  # CHECK, rock.py:171-175, 401-407
        rock = action - 5
        rx, ry = cfg.rock_pos[rock]
        eff = cfg.efficiency(l1_distance(x, y, rx, ry))
        if draw(1) < bern_threshold(eff):
            ob = 2 if status[rock] == 1 else 1
        else:
            ob = 1 if status[rock] == 1 else 2
    if cfg.stochastic:
        done = False  # rock.py:503 (commented out) -- only the EAST exit terminates
    else:
        done = reward == cfg.penalization  # rock.py:193
    return x, y, status, ob, reward, done, err


def rock_reset_word(slot_word, rock):
    """Rock i's uniform is u = r_i / 2**32, r_i = rotl32(word of reset slot 0, 30 - 2 i): all rocks (k <= 16) share
    one draw word -- the deciding top bit of r_i is bit 2 i + 1 of it, sixteen different bits."""
    sh = (30 - 2 * rock) & 31
    return ((slot_word << sh) | (slot_word >> (32 - sh))) & 0xFFFFFFFF if sh else slot_word


def rock_reset(cfg, draw):
    """rock.py:236-241, 266-271, 78-80.  Rock i's ``uniform(0,1)`` comes from reset slot 0 (rock_reset_word).

    status = int(sign(u - .5)): -1 below one half, +1 above, 0 exactly at u == 0.5.
    """
    status = []
    for i in range(cfg.k):
        r = rock_reset_word(draw(0), i)
        status.append((r > (1 << 31)) - (r < (1 << 31)))
    return cfg.start[0], cfg.start[1], status, 0


def rock_belief_update(cfg, x, y, action, ob, side):
    """rock.py:177-191: the checked rock's side-statistics (``side`` = list of dicts with count, measured, lkv, lkw,
    prob_valuable; mutated).  (x, y) is the agent position, which a check does not change."""
    if action <= 4 or ob == 0:
        return
    r = side[action - 5]
    eff = cfg.efficiency(l1_distance(x, y, *cfg.rock_pos[action - 5]))
    r["measured"] += 1
    if ob == 2:
        r["count"] += 1
        r["lkv"] *= eff
        r["lkw"] *= (1 - eff)
    else:
        r["count"] -= 1
        r["lkw"] *= eff
        r["lkv"] *= (1 - eff)
    denom = (.5 * r["lkv"]) + (.5 * r["lkw"])
    r["prob_valuable"] = (.5 * r["lkv"]) / denom if denom != 0 else float("nan")   # numpy: 0/0 -> nan (+ warning)


def rock_compute_prob(cfg, action, x, y, status, ob):
    """rock.py:250-264 (evaluated on the post-step state)."""
    if action <= 4:
        return float(ob == 0)
    rock = action - 5
    eff = cfg.efficiency(l1_distance(x, y, *cfg.rock_pos[rock]))
    if ob == 2 and status[rock] == 1:
        return eff
    if ob == 1 and status[rock] == -1:
        return eff
    return 1 - eff


def rock_generate_legal(cfg, x, y, status):
    """rock.py:273-291 (order of the reference's list preserved)."""
    legal = [1]
    if y + 1 < cfg.n:
        legal.append(0)
    if y - 1 >= 0:
        legal.append(2)
    if x - 1 >= 0:
        legal.append(3)
    rock = cfg.grid[x][y]
    if 0 <= rock < cfg.k and status[rock] != 0:
        legal.append(4)
    for i in range(cfg.k):
        if status[i] != 0:
            legal.append(cfg.grid[cfg.rock_pos[i][0]][cfg.rock_pos[i][1]] + 5)
    return legal


# ---------------------------------------------------------------------------------------
# Tag (tag.py)
# ---------------------------------------------------------------------------------------
TAG_CELLS = 29


def tag_is_inside(x, y):
    """tag.py:46-50"""
    if y >= 2:
        return 5 <= x < 8 and y < 5
    return 0 <= x < 10 and y >= 0


def tag_get_coord(idx):
    """tag.py:52-57"""
    if idx < 20:
        return idx % 10, idx // 10
    idx -= 20
    return idx % 3 + 5, idx // 3 + 2


def tag_get_index(x, y):
    """tag.py:59-66"""
    if y < 2:
        return y * 10 + x
    return 20 + (y - 2) * 3 + x - 5


def tag_admissible(ax, ay, ox, oy):
    """tag.py:260-280: the MULTISET of moves (as Moves indices) the opponent picks from."""
    acts = []
    if ox >= ax:
        acts.append(1)  # EAST
    if oy >= ay:
        acts.append(0)  # NORTH
    if ox <= ax:
        acts.append(3)  # WEST
    if oy <= ay:
        acts.append(2)  # SOUTH
    if ox == ax and oy > ay:
        acts.append(0)
    if oy == ay and ox > ax:
        acts.append(1)
    if ox == ax and oy < ay:
        acts.append(2)
    if oy == ay and ox < ax:
        acts.append(3)
    return acts


def tag_sample_ob(ax, ay, opps, action):
    """tag.py:219-226"""
    ob = tag_get_index(ax, ay)
    if action < 4:
        for (ox, oy) in opps:
            if (ox, oy) == (ax, ay):
                ob = TAG_CELLS
    return ob


def tag_step(ax, ay, opps, num_opp, action, draw, move_prob=0.8):
    """tag.py:108-143.  Returns (ax, ay, opps, num_opp, obs, reward, done).

    Draw slot j is opponent j's word: ``binomial(1, move_prob)`` (tag.py:204) reads it whole, ``choice(actions)``
    (tag.py:205; consumed only when the first says move) reads its low half, ``tag_pick_word``.
    """
    opps = [tuple(o) for o in opps]
    if action == 4:
        tagged = False
        reward = 0.
        for j, (ox, oy) in enumerate(opps):
            if (ox, oy) == (ax, ay):
                reward = 10.
                tagged = True
                num_opp -= 1
            elif tag_is_inside(ox, oy) and num_opp > 0:
                acts = tag_admissible(ax, ay, ox, oy)  # tag.py:201-207
                w = draw(j)
                if w < bern_threshold(move_prob):
                    dx, dy = MOVES[acts[rand_below(tag_pick_word(w), len(acts))]]
                    if tag_is_inside(ox + dx, oy + dy):
                        opps[j] = (ox + dx, oy + dy)
        if not tagged:
            reward = -10.
    else:
        reward = -1.
        dx, dy = MOVES[action]
        if tag_is_inside(ax + dx, ay + dy):
            ax, ay = ax + dx, ay + dy
    ob = tag_sample_ob(ax, ay, opps, action)
    return ax, ay, opps, num_opp, ob, reward, num_opp == 0


def tag_pick_word(w):
    """The word np.random.choice (tag.py:205) is scripted with: the low half of the opponent's draw word."""
    return (int(w) << 16) & 0xFFFFFFFF


def tag_reset_word(slot_word, digit):
    """The word np.random.randint(29) is scripted with for the ``digit``-th (0..2) cell drawn from one reset slot:
    floor(u' * 29) with u' = frac(u * 29**digit) is the digit-th base-29 digit of u = slot_word / 2**32."""
    return (slot_word * 29 ** digit) & 0xFFFFFFFF


def tag_reset(num_opponents, draw):
    """tag.py:97-102, 181-193, 43-44.  The j-th randint(29) (j = 0 agent, 1 + i opponent i) is digit j % 3 of
    reset slot j // 3 (tag_reset_word)."""
    cell = lambda j: rand_below(tag_reset_word(draw(j // 3), j % 3), TAG_CELLS)
    ax, ay = tag_get_coord(cell(0))
    opps = [tag_get_coord(cell(1 + j)) for j in range(num_opponents)]
    return ax, ay, opps, num_opponents, tag_sample_ob(ax, ay, opps, 0)


def tag_compute_prob(ax, ay, opps, ob):
    """tag.py:209-217"""
    p = int(ob == tag_get_index(ax, ay))
    if ob == TAG_CELLS:
        for o in opps:
            if tuple(o) == (ax, ay):
                return 1.
    return p


# ---------------------------------------------------------------------------------------
# BattleShip (battleship.py)
# ---------------------------------------------------------------------------------------


class ShipBoard:
    """battleship.py:46-61: per-cell occupied / visited flags indexed [x][y]."""

    def __init__(self, x_size, y_size):
        self.x_size, self.y_size = x_size, y_size
        self.occupied = [[False] * y_size for _ in range(x_size)]
        self.visited = [[False] * y_size for _ in range(x_size)]
        self.total_remaining = 0

    def copy(self):
        b = ShipBoard(self.x_size, self.y_size)
        b.occupied = [col[:] for col in self.occupied]
        b.visited = [col[:] for col in self.visited]
        b.total_remaining = self.total_remaining
        return b


def ship_collision(board, x, y, direction, length):
    """battleship.py:195-211.  Quirks kept: ``length + 1`` cells are walked and the cell
    AFTER each must be inside; the adjacency loop is ``range(8)`` = N,E,S,W,Null,NE,SE,SW
    (NorthWest is never looked at)."""
    dx, dy = COMPASS[direction]
    for _ in range(length + 1):
        if not grid_is_inside(board.x_size, board.y_size, x + dx, y + dy):
            return True
        if board.occupied[x][y]:
            return True
        for adj in range(8):
            cx, cy = x + COMPASS[adj][0], y + COMPASS[adj][1]
            if grid_is_inside(board.x_size, board.y_size, cx, cy) and board.occupied[cx][cy]:
                return True
        x, y = x + dx, y + dy
    return False


def ship_mark(board, x, y, direction, length):
    """battleship.py:182-193"""
    dx, dy = COMPASS[direction]
    for _ in range(length):
        assert not board.occupied[x][y]
        board.occupied[x][y] = True
        if not board.visited[x][y]:
            board.total_remaining += 1
        x, y = x + dx, y + dy


def ship_lengths(max_len):
    """battleship.py:74-75, 171: ``self.max_len = max_len + 1``; reversed(range(2, self.max_len))."""
    return list(reversed(range(2, max_len + 1)))


def battleship_reset_rejection(x_size, y_size, max_len, draw, max_attempts=4096):
    """battleship.py:131-137, 167-180, 33-37, coord.py:68-69 -- the reference's own loop.

    Attempt a (counted across ships) uses slot 2a = ``randint(n_tiles)`` for the position
    and slot 2a+1 = ``randint(4)`` for the direction.  Returns (board, attempts).
    """
    board = ShipBoard(x_size, y_size)
    a = 0
    for length in ship_lengths(max_len):
        while True:
            assert a < max_attempts
            x, y = grid_get_coord(x_size, rand_below(draw(2 * a), x_size * y_size))
            d = rand_below(draw(2 * a + 1), 4)
            a += 1
            if not ship_collision(board, x, y, d, length):
                break
        ship_mark(board, x, y, d, length)
    return board, a


def battleship_valid_placements(board, length):
    """All candidates c = 4*pos + dir (pos = x_size*y + x) the rejection loop would accept."""
    out = []
    for pos in range(board.x_size * board.y_size):
        x, y = grid_get_coord(board.x_size, pos)
        for d in range(4):
            if not ship_collision(board, x, y, d, length):
                out.append(4 * pos + d)
    return out


def battleship_reset_scan(x_size, y_size, max_len, draw):
    """Same distribution as the rejection loop (uniform over the accepted set, given the
    earlier ships) in fixed time: ship s takes the k-th valid candidate in increasing
    c = 4*pos + dir order, k = floor(u * count) from draw slot s.  Returns (board, ok)."""
    board = ShipBoard(x_size, y_size)
    for s, length in enumerate(ship_lengths(max_len)):
        valid = battleship_valid_placements(board, length)
        if not valid:
            return board, False  # the reference would loop forever
        c = valid[rand_below(draw(s), len(valid))]
        x, y = grid_get_coord(x_size, c >> 2)
        ship_mark(board, x, y, c & 3, length)
    return board, True


def battleship_step(board, action):
    """battleship.py:91-122 (the ``diagonal`` writes at 111-113 are dead state).
    Mutates ``board``; returns (obs, reward, done)."""
    x, y = grid_get_coord(board.x_size, action)
    reward = 0
    if board.visited[x][y]:
        reward -= 10
        obs = 0
    else:
        if board.occupied[x][y]:
            reward -= 1
            obs = 1
            board.total_remaining -= 1
        else:
            reward -= 1
            obs = 0
        board.visited[x][y] = True
    done = False
    if board.total_remaining == 0:
        reward += board.x_size * board.y_size
        done = True
    return obs, reward, done


def battleship_compute_prob(board, action, ob):
    """battleship.py:80-89"""
    x, y = grid_get_coord(board.x_size, action)
    if ob == 0 and board.visited[x][y]:
        return 1
    if ob == 1 and board.occupied[x][y]:
        return 1
    return int(ob == 0)


# ---------------------------------------------------------------------------------------
# Tiger (tiger.py)
# ---------------------------------------------------------------------------------------


def tiger_step(state, action, draw, listen_prob=0.85):
    """tiger.py:72-88, 117-119, 140-172.  Returns (state, obs, reward, done).

    Slot 0 serves both ``state_space.sample()`` (gym's RNG; only for a in {0,1}, tiger.py:118-119) and
    ``np.random.uniform()`` (always drawn when not terminal, tiger.py:143, but read only after LISTEN): a step never
    consumes both.
    ``_sample_ob`` ignores ``self.correct_prob`` (default argument .85, tiger.py:141).
    """
    if action == 2:
        reward = -1
    elif action != state:
        reward = 10
    else:
        reward = -20
    if action != 2 and action == state:
        return state, state, reward, True  # tiger.py:81-83: ob = state on terminal
    if action in (0, 1):
        state = rand_below(draw(0), 2)
    ob = 2
    flip = draw(0) > gt_threshold(listen_prob)
    if action == 2:
        if state == 0:
            ob = 1 if flip else 0
        else:
            ob = 0 if flip else 1
    return state, ob, reward, False


def tiger_reset(draw):
    """tiger.py:60-66: state from gym's Discrete.sample (slot 0); ob = NULL (2)."""
    return rand_below(draw(0), 2), 2


def tiger_compute_prob(action, next_state, ob, correct_prob=.85):
    """tiger.py:125-138"""
    p = 0.0
    if action == 2 and ob != 2:
        p = correct_prob if next_state == ob else 1 - correct_prob
    elif action != 2 and ob == 2:
        p = 1.
    return p


# ---------------------------------------------------------------------------------------
# Network (network.py)
# ---------------------------------------------------------------------------------------


def network_neighbours(n_machines, problem_type=3):
    """network.py:144-168 (3-legs is asymmetric and links idx <= 4, not 3, to machine 0)."""
    nb = [[] for _ in range(n_machines)]
    if problem_type != 3:
        for i in range(n_machines):
            nb[i].append((i + 1) % n_machines)
            nb[i].append((i + n_machines - 1) % n_machines)
        return nb
    assert n_machines >= 4 and n_machines % 3 == 1
    nb[0] += [1, 2, 3]
    for i in range(1, n_machines):
        if i < n_machines - 3:
            nb[i].append(i + 3)
        if i <= 4:
            nb[i].append(0)
        else:
            nb[i].append(i - 3)
    return nb


def network_step(state, action, draw, neighbours, p=0.1, q=0.33, p_ob=0.95):
    """network.py:71-114.  Returns (state, obs, reward_tenths, done).

    Slot m = machine m's failure draw (consumed only if m is up, network.py:94-99),
    slot n = the action's observation draw (network.py:101-112).  The reward is returned
    as an exact integer number of tenths; the reference's double is ``tenths / 10.0``
    (verified value-for-value by the golden fixtures).
    """
    n = len(state)
    state = list(state)
    n_fail = [0] * n
    for i in range(n):
        for j in neighbours[i]:
            if state[j] == 0:
                n_fail[i] = 1
    tenths = 0
    for i in range(n):
        if state[i] == 1:
            tenths += 20 if len(neighbours[i]) > 2 else 10
    for i in range(n):
        if state[i]:
            pr = q if n_fail[i] else p
            state[i] = 1 - (1 if draw(i) < bern_threshold(pr) else 0)
    ob = 2
    if action < 2 * n:
        machine, reboot = divmod(action, 2)
        hit = 1 if draw(n) < bern_threshold(p_ob) else 0
        if reboot:
            tenths -= 25
            state[machine] = 1
            ob = hit
        else:
            tenths -= 1
            ob = state[machine] if hit else 1 - state[machine]
    return state, ob, tenths, False


def network_reset(n_machines):
    """network.py:61-69: all up; returns Obs.OFF (0), not NULL."""
    return [1] * n_machines, 0


def network_compute_prob(action, next_state, ob, p_ob=0.95):
    """network.py:43-55"""
    n = len(next_state)
    if action < 2 * n:
        return p_ob if next_state[action // 2] == ob else 1 - p_ob
    return 1. if ob == 2 else 0


# ---------------------------------------------------------------------------------------
# Legal actions and uniform-legal rollouts
# ---------------------------------------------------------------------------------------
def tag_generate_legal():
    """tag.py:228-229"""
    return list(range(5))


def tiger_generate_legal():
    """tiger.py:111-112"""
    return list(range(3))


def network_generate_legal(n_machines):
    """network.py:129-130"""
    return list(range(2 * n_machines + 1))


def battleship_generate_legal(board):
    """battleship.py:157-165: the unvisited cells, in increasing action order"""
    legal = []
    for action in range(board.x_size * board.y_size):
        x, y = grid_get_coord(board.x_size, action)
        if not board.visited[x][y]:
            legal.append(action)
    return legal


def rollout(legal, step, is_done, max_steps, gamma, draws):
    """The loop at rock.py:563-572 / tag.py:310-316 for ONE env instance::

        while not done and t < max_steps:
            a = np.random.choice(env._generate_legal())
            ob, rw, done, _ = env.step(a)
            r += rw * discount;  discount *= env._discount

    ``legal()`` -> list, ``step(a, draw)`` -> (reward, done) (mutating the caller's state),
    ``is_done()`` -> bool; ``draws(t)`` -> (policy_word, draw) for rollout step t, i.e. the
    words of step counter c + t: domain POLICY slot 0 for ``choice`` (index = floor(u * len)),
    domain STEP for the step's own slots.  Returns (ret, steps, actions taken).
    """
    ret, disc, t, actions = 0.0, 1.0, 0, []
    while t < max_steps and not is_done():
        w, draw = draws(t)
        lst = legal()
        a = lst[rand_below(w, len(lst))]
        rw, _ = step(a, draw)
        ret += rw * disc
        disc *= gamma
        actions.append(a)
        t += 1
    return ret, t, actions
