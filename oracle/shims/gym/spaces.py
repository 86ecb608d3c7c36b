"""Shape-metadata-only stand-ins for the four gym.spaces classes the reference touches."""


class Box:
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(int(s) for s in shape)  # as gym.spaces.Box does


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


class MultiBinary:
    def __init__(self, n):
        self.n = int(n)
        self.shape = (self.n,)


class Tuple:
    def __init__(self, spaces):
        self.spaces = list(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)
