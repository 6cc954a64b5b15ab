"""Level model and bit-plane packers (host side).

A level is the static part of a GridUniverse: shape, walls, goal / lava terminals and
start states.  The text format and its error behaviour follow the reference parser
(core/envs/griduniverse_env.py:246-300); constructor-style levels follow :44-90,120-134.
The packers produce the two device layouts of include/gu_b200.h:

* dense planes for env batches -- bit (s & 31) of word (s >> 5) is cell s; per-env
  batches are WORD-MAJOR uint32[words][N];
* row-pitched planes with one ghost row above and below for the sweep kernels.
"""
import numpy as np


def _reject_negative(t):
    # Deliberate deviation: the reference lets a negative goal / lava index wrap around in
    # reward_matrix while `state in goal_states` never matches it (a +10 cell that is not
    # terminal).  Rewards are derived from the terminal masks here, so reject it instead.
    if isinstance(t, (int, np.integer)) and t < 0:
        raise IndexError("negative terminal state index {}".format(t))


class Level(object):
    """Shape + terminals + walls + starts of one grid (no device state)."""

    def __init__(self, X, Y, walls=(), goals=None, lavas=(), starts=(0,)):
        self.X, self.Y = int(X), int(Y)
        self.N = self.X * self.Y
        # default goal = last cell when none / empty given (griduniverse_env.py:66-67)
        if goals is None or len(goals) == 0:
            goals = [self.N - 1]
        self.goal_states = goals
        self.lava_states = list(lavas) if lavas is not None else []
        self.starting_states = list(starts)
        self.wall_indices = []
        self.wall = np.zeros(self.N, dtype=bool)
        self.goal = np.zeros(self.N, dtype=bool)
        self.lava = np.zeros(self.N, dtype=bool)
        if walls is not None:
            for w in walls:
                if w < 0 or w > (self.N - 1):  # :130-131
                    raise ValueError("Wall state {} is out of grid bounds".format(w))
                self.wall[w] = True
                self.wall_indices.append(w)
        # reward_matrix (:80-90): -1, goals +10, then lava -10; a bad index raises IndexError
        self.reward_matrix = np.full(self.N, -1)
        for t in self.goal_states:
            try:
                _reject_negative(t)
                self.reward_matrix[t] = 10
                self.goal[t] = True
            except IndexError:
                raise IndexError("Terminal goal state {} is out of grid bounds or is wrong type. "
                                 "Should be an integer.".format(t))
        for t in self.lava_states:
            try:
                _reject_negative(t)
                self.reward_matrix[t] = -10
                self.lava[t] = True
            except IndexError:
                raise IndexError("Lava terminal state {} is out of grid bounds or is wrong type. "
                                 "Should be an integer.".format(t))

    @classmethod
    def from_masks(cls, X, Y, wall, goal, lava, starts=(0,)):
        """Build from boolean masks without materialising index lists (large grids)."""
        lv = cls.__new__(cls)
        lv.X, lv.Y, lv.N = int(X), int(Y), int(X) * int(Y)
        lv.wall = np.ascontiguousarray(wall, dtype=bool).reshape(-1)
        lv.goal = np.ascontiguousarray(goal, dtype=bool).reshape(-1)
        lv.lava = np.ascontiguousarray(lava, dtype=bool).reshape(-1)
        assert lv.wall.size == lv.N and lv.goal.size == lv.N and lv.lava.size == lv.N
        lv.starting_states = list(starts)
        lv.goal_states = lv.lava_states = lv.wall_indices = None   # materialise on demand
        lv.reward_matrix = None
        return lv

    def index_lists(self):
        """(goal_states, lava_states, wall_indices) as python lists."""
        if self.goal_states is None:
            self.goal_states = [int(i) for i in np.flatnonzero(self.goal)]
            self.lava_states = [int(i) for i in np.flatnonzero(self.lava)]
            self.wall_indices = [int(i) for i in np.flatnonzero(self.wall)]
        return self.goal_states, self.lava_states, self.wall_indices

    def rewards(self):
        if self.reward_matrix is None:
            r = np.full(self.N, -1)
            r[self.goal] = 10
            r[self.lava] = -10
            self.reward_matrix = r
        return self.reward_matrix

    def to_text_lines(self):
        cells = np.full(self.N, 'o', dtype='<U1')
        cells[self.starting_states] = 'x'
        cells[self.goal] = 'G'
        cells[self.lava] = 'L'
        cells[self.wall] = '#'
        return [''.join(cells[y * self.X:(y + 1) * self.X]) for y in range(self.Y)]


def read_level_file(fp):
    """A level file as the list of its non-blank rows with every whitespace character removed
    (what griduniverse_env.py:246-251 feeds the text parser)."""
    with open(fp, 'r') as f:
        rows = ("".join(raw.split()) for raw in f)
        return [row for row in rows if row]


# cell classes of the level alphabet (griduniverse_env.py:282-291); anything else is invalid
_FLOOR, _WALL, _GOAL, _LAVA, _START, _INVALID = 0, 1, 2, 3, 4, 255
_CELL_CLASS = np.full(256, _INVALID, dtype=np.uint8)
for _ch, _cls in (("o", _FLOOR), ("#", _WALL), ("G", _GOAL), ("L", _LAVA), ("x", _START)):
    _CELL_CLASS[ord(_ch)] = _cls


def parse_level_text(rows):
    """Rows of 'o' floor / '#' wall / 'G' goal / 'L' lava / 'x' start -> Level, with the reference's
    ValueErrors in the order its row-major scan meets them (griduniverse_env.py:253-300): an invalid
    character in an earlier row wins over a ragged later row; missing 'x' / 'G' are reported after the
    scan.  The whole grid is classified with one byte look-up table instead of a per-character loop
    (a 16384 x 16384 level is 268 M characters)."""
    X = len(rows[0])
    ragged = next((i for i, row in enumerate(rows) if len(row) != X), None)
    scanned = rows if ragged is None else rows[:ragged]
    text = "".join(scanned)
    cls = _CELL_CLASS[np.frombuffer(text.encode("latin-1", "replace"), dtype=np.uint8)]
    bad = np.flatnonzero(cls == _INVALID)
    if bad.size:
        raise ValueError('Invalid Character "{}". Returning'.format(text[int(bad[0])]))
    if ragged is not None:
        raise ValueError("Input text file is not a rectangle")
    cells = {c: np.flatnonzero(cls == c).tolist() for c in (_WALL, _GOAL, _LAVA, _START)}
    if not cells[_START]:
        raise ValueError("No starting states set in text file. Place \"x\" within grid. ")
    if not cells[_GOAL]:
        raise ValueError("No terminal goal states set in text file. Place \"T\" within grid. ")
    return Level(X, len(rows), walls=cells[_WALL], goals=cells[_GOAL], lavas=cells[_LAVA], starts=cells[_START])


# --------------------------------------------------------------------------
# packers
# --------------------------------------------------------------------------
def pack_dense(mask):
    """bool[..., cells] -> uint32[..., ceil(cells/32)], bit (s & 31) of word (s >> 5) = cell s."""
    mask = np.asarray(mask, dtype=bool)
    cells = mask.shape[-1]
    words = (cells + 31) // 32
    pad = words * 32 - cells
    if pad:
        mask = np.concatenate([mask, np.zeros(mask.shape[:-1] + (pad,), dtype=bool)], axis=-1)
    by = np.packbits(mask, axis=-1, bitorder='little')
    return np.ascontiguousarray(by).view(np.uint32).reshape(mask.shape[:-1] + (words,))


def pack_env_planes(masks):
    """bool[N, cells] -> WORD-MAJOR uint32[words, N] (include/gu_b200.h, gu_levels)."""
    return np.ascontiguousarray(pack_dense(masks).T)


def grid_pitch(X):
    """Row pitch in elements of every per-cell array: X rounded up to 32, so each row of any
    element type (uint8 tie masks included) starts on a 16-byte boundary and lines up with
    the bit-plane words (the tiled kernels move rows as 16-byte vectors / TMA boxes)."""
    return ((X + 31) // 32) * 32


def grid_pitch_words(X):
    """uint32 words per bit-plane row, padded to a multiple of 4 words (16 bytes)."""
    w = (X + 31) // 32
    return ((w + 3) // 4) * 4


def pack_grid_plane(mask2d, row_begin, row_end, pitch_words):
    """bool[Y, X] -> uint32[(rows+2) * pitch_words] holding rows [row_begin-1, row_end+1)
    (ghost rows outside the grid are zero); bit (x & 31) of word (x >> 5)."""
    Y, X = mask2d.shape
    rows = row_end - row_begin
    out = np.zeros((rows + 2, pitch_words), dtype=np.uint32)
    lo, hi = max(row_begin - 1, 0), min(row_end + 1, Y)
    packed = pack_dense(mask2d[lo:hi])
    out[lo - (row_begin - 1):hi - (row_begin - 1), :packed.shape[1]] = packed
    return out.reshape(-1)
