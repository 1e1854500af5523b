"""``Simple`` -- the reference's PIL toy environment (envs/simple.py:12-143), the one BASELINE
configuration that "runs on CPU" (README debug command).  It is not on the GPU hot path: the
canvases are PIL images, batched by a plain Python loop, and the arrays are numpy float64 in
[-1, 1] exactly like the reference's ``state``."""
import numpy as np
from PIL import Image, ImageDraw

from . import utils
from .base import Environment


class Simple(Environment):

    # source order of the reference's dict literal (envs/simple.py:14-18)
    action_sizes = [
        ('color', [1]),
        ('shape', [2]),
    ]

    def __init__(self, args, num_envs=None):
        super(Simple, self).__init__(args, num_envs=num_envs)
        assert self.conditional, "Don't train a Simple env with --conditional=False"

        self.object_radius = 3
        self.object_size = self.object_radius * 2
        self.background_color = (255, 255, 255)
        self.colors = [
            (0, 0, 0),         # black
            (173, 181, 189),   # gray 5
            (255, 224, 102),   # yellow 3
        ]
        if 'color' in self.action_sizes:
            self.colors = self.colors[:self.action_sizes['color'][0]]
        self.shapes = ['circle', 'rectangle']
        if 'shape' in self.action_sizes:
            self.shapes = self.shapes[:self.action_sizes['shape'][0]]
        self.locations = utils.uniform_locations(self.screen_size, self.location_size, self.object_radius)
        self.images = None
        self.random_target = None
        self._np_rng = np.random.RandomState(getattr(args, "seed", 123))

    def _new_image(self):
        return Image.new('RGB', (self.width, self.height), self.background_color)

    def reset(self):
        self.random_target = np.stack([self._one_target() for _ in range(self.num_envs)])
        self.images = [self._new_image() for _ in range(self.num_envs)]
        self._step = 0
        self.z = None
        return self.state, self.random_target, self.z

    def _draw_one(self, ac, image):
        r = self.object_radius
        x, y = self.locations[0]
        color, shape = self.colors[0], self.shapes[0]
        for name in self.action_sizes:
            named_ac = int(ac[self.ac_idx[name]])
            value = getattr(self, name + "s")[named_ac]
            if name == 'location':
                x, y = value
            elif name == 'color':
                color = value
            elif name == 'shape':
                shape = value
        drawer = ImageDraw.Draw(image)
        if shape == 'circle':
            drawer.ellipse((x - r, y - r, x + r, y + r), fill=color)
        elif shape == 'rectangle':
            drawer.rectangle((x - r, y - r, x + r, y + r), fill=color)
        else:
            raise Exception("Unkown shape: {}".format(shape))

    def _one_target(self):
        image = self._new_image()
        for _ in range(self.episode_length):
            ac = [self._np_rng.randint(n) for n in self.head_sizes]
            self._draw_one(ac, image)
        return np.array(self.norm(image))

    def get_random_target(self, num=None):
        return np.stack([self._one_target() for _ in range(num or self.num_envs)])

    def step(self, acs):
        acs = np.asarray(acs).reshape(self.num_envs, len(self.acs))
        for ac, image in zip(acs, self.images):
            self._draw_one(ac, image)
        self._step += 1
        terminal = (self._step == self.episode_length)
        if terminal:
            state = self.state
            n = np.prod(self.observation_shape)
            reward = np.array([-utils.l2(s, t) / n * 100 for s, t in zip(state, self.random_target)])
        else:
            reward = np.zeros(self.num_envs)
        # XXX: DEBUG override kept from the reference (simple.py:132-133)
        reward = np.where(reward == 0, 1.0, reward)
        return self.state, reward, terminal, {}

    @property
    def state(self):
        return np.stack([np.array(self.norm(image)) for image in self.images])
