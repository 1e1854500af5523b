"""Batched ``Environment`` base class -- same bookkeeping as the reference's envs/base.py:6-55,
with every per-canvas quantity gaining a leading batch axis B.

The action-column order (``self.acs``) reproduces the CPython-2.7 dict iteration order the
reference relies on at envs/base.py:32 (it is also the autoregressive order of the policy
heads, models/policy.py:224), so ``actions[..., k]`` means the same head here and there.
"""
from collections import OrderedDict

import numpy as np


def py2_dict_order(names):
    """Iteration order of a CPython-2.7 (64-bit, hash randomisation off) dict literal whose string
    keys appear in this source order."""
    M = 0xFFFFFFFFFFFFFFFF

    def strhash(s):
        if not s:
            return 0
        x = (ord(s[0]) << 7) & M
        for ch in s:
            x = ((1000003 * x) ^ ord(ch)) & M
        x ^= len(s)
        if x == M:          # -1 is reserved
            x = M - 1
        return x

    def insert(table, key):
        mask = len(table) - 1
        h = strhash(key)
        i = h & mask
        perturb = h
        while table[i & mask] is not None:
            i = ((i << 2) + i + perturb + 1) & M
            perturb >>= 5
        table[i & mask] = key

    size = 8
    if len(names) > 5:      # BUILD_MAP n -> _PyDict_NewPresized(n)
        while size <= len(names):
            size <<= 1
    table = [None] * size
    fill = 0
    for k in names:
        insert(table, k)
        fill += 1
        if fill * 3 >= len(table) * 2:      # dictresize(used * 4)
            new = 8
            while new <= fill * 4:
                new <<= 1
            old = [t for t in table if t is not None]
            table = [None] * new
            for kk in old:
                insert(table, kk)
    return [t for t in table if t is not None]


class Environment(object):
    """Subclasses define ``action_sizes`` as a list of (name, size) in the SOURCE order of the
    reference's class-level dict literal; ``None`` sizes become ``[L, L]``."""

    action_sizes = []

    def __init__(self, args, num_envs=None):
        self.args = args
        self.num_envs = int(num_envs if num_envs is not None else getattr(args, "num_envs", 1))

        sizes = dict(self.action_sizes)
        order = py2_dict_order([n for n, _ in self.action_sizes])
        if not args.jump and 'jump' in sizes:
            del sizes['jump']
        if not args.curve and 'control' in sizes:
            del sizes['control']

        # terminal
        self.episode_length = args.episode_length

        # screen
        self.screen_size = args.screen_size
        self.height, self.width = self.screen_size, self.screen_size
        self.observation_shape = [self.height, self.width, args.color_channel]

        # location
        self.location_size = args.location_size
        self.location_shape = [self.location_size, self.location_size]

        self.action_sizes = OrderedDict()
        for name in order:
            if name not in sizes:
                continue
            value = sizes[name]
            self.action_sizes[name] = list(self.location_shape) if value is None else list(value)

        self.acs = list(self.action_sizes.keys())
        self.ac_idx = {ac: idx for idx, ac in enumerate(self.acs)}
        self.head_sizes = [int(np.prod(self.action_sizes[ac])) for ac in self.acs]

        self.conditional = args.conditional

    def random_action(self, rng=None):
        """int array [B, K]; head k ~ randint(prod(size_k)) (reference base.py:39-45)."""
        rng = np.random if rng is None else rng
        return np.stack([rng.randint(n, size=self.num_envs) for n in self.head_sizes], 1).astype(np.int32)

    @property
    def initial_action(self):
        """[-1] * K per canvas (reference base.py:47-49)."""
        return np.full((self.num_envs, len(self.acs)), -1, np.int32)

    def norm(self, img):
        return (np.array(img) - 127.5) / 127.5 if not hasattr(img, "sub") else (img - 127.5) / 127.5

    def denorm(self, img):
        return img * 127.5 + 127.5
