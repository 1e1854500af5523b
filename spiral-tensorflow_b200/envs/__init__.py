"""``create_env(args)`` -- same factory as the reference's envs/__init__.py:10-20, returning the
batched, GPU-resident environments."""
from .mnist import MNIST, SimpleMNIST, ColorMNIST
from .simple import Simple


def create_env(args, num_envs=None, device=None):
    env = args.env.lower()
    if env == 'simple':
        env = Simple(args, num_envs=num_envs)
    elif env == 'simple_mnist':
        env = SimpleMNIST(args, num_envs=num_envs, device=device)
    elif env == 'mnist':
        env = MNIST(args, num_envs=num_envs, device=device)
    elif env == 'color_mnist':      # EXTENSION: MNIST heads + colour head (BASELINE rgb64 config)
        env = ColorMNIST(args, num_envs=num_envs, device=device)
    else:
        raise Exception("Unkown environment: {}".format(args.env))
    return env
