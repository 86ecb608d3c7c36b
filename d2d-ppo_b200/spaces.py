"""Shape metadata for observation / action / state spaces.

The reference uses ``gym.spaces`` purely as shape containers (envs/env.py:43-48,
envs/combinatorial_env.py:47-58, envs/channel_selection_env.py:41-46; read back at
algorithms/d2d_ppo.py:252-253,265).  ``gym`` is not a dependency here; these classes expose the
attributes the agents read: ``[k]``, ``.shape``, ``.n``.
"""


class Box:
    def __init__(self, low=-float("inf"), high=float("inf"), shape=()):
        self.low, self.high, self.shape = low, high, tuple(int(s) for s in shape)

    def __repr__(self):
        return f"Box(shape={self.shape})"


class Discrete:
    def __init__(self, n):
        self.n, self.shape = int(n), ()

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiBinary:
    def __init__(self, n):
        self.n = int(n)
        self.shape = (self.n,)

    def __repr__(self):
        return f"MultiBinary({self.n})"


class Tuple:
    def __init__(self, spaces):
        self.spaces = list(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)
