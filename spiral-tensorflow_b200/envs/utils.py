"""envs/utils.py of the reference (rgb2gray, l2, uniform_locations), Python-3 safe."""
import numpy as np


def rgb2gray(rgb):
    """BT.601 luma, keeps a trailing channel axis (reference envs/utils.py:3-5)."""
    rgb = np.dot(rgb[..., :3], [0.299, 0.587, 0.114])
    return np.expand_dims(rgb, -1)


def l2(mat1, mat2):
    """Frobenius norm of the difference (reference envs/utils.py:7-8)."""
    return np.sqrt(np.sum((mat1 - mat2) ** 2))


def uniform_locations(screen_size, location_size, object_radius, normalize=False):
    """L*L grid, index k -> (lin[k % L], lin[k // L]) (reference envs/utils.py:10-18; the py2
    ``np.array(zip(...))`` idiom is spelled out so it also runs on Python 3)."""
    x = np.linspace(object_radius, screen_size - object_radius, location_size)
    grid = np.meshgrid(x, x)
    out = np.array(list(zip(*np.vstack(list(map(np.ravel, grid))))))
    if normalize:
        div = location_size ** 2 / 2
        out = (out - div) / div
    return out
