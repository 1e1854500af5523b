"""GPU tests of K4 (categorical sampling for the action heads, models/policy.py:364-368)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sp():
    import spiral_b200 as sp
    return sp


def _sample(logits, seed, offset):
    N = _sp()._native
    lg = torch.tensor(logits, device="cuda")
    out = torch.empty((lg.shape[0],), dtype=torch.int32, device="cuda")
    N.check(N.lib().spiral_categorical_sample(N.ptr(lg), lg.shape[0], lg.shape[1], C.c_uint64(seed), C.c_uint64(offset), None,
                                              N.ptr(out), N.stream_ptr()))
    return out.cpu().numpy()


def test_philox_block_matches_random123_vectors():
    N = _sp()._native
    out = torch.zeros(4, dtype=torch.int32, device="cuda")
    for ck, want in (([0] * 6, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
                     ([0xffffffff] * 6, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
                     ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0],
                      [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])):
        arr = (C.c_uint32 * 6)(*ck)
        N.check(N.lib().spiral_debug_philox(arr, N.ptr(out), N.stream_ptr()))
        got = [int(v) & 0xffffffff for v in out.cpu().numpy()]
        assert got == want


@pytest.mark.parametrize("rows,n", [(64, 2), (257, 8), (33, 1024), (16, 4096), (5, 4097)])
def test_sampler_matches_oracle_restatement(rows, n):
    from oracle import sample_oracle as so
    rng = np.random.RandomState(rows + n)
    logits = (rng.randn(rows, n) * 2).astype(np.float32)
    got = _sample(logits, seed=1234567890123, offset=7)
    want = so.categorical_sample(logits, 1234567890123, 7)
    # identical uniforms; logf on the device and numpy's log may differ in the last ulp, which can only flip
    # an arg-max between two near-tied classes
    assert (got == want).mean() >= 0.99
    assert (got >= 0).all() and (got < n).all()
    assert (_sample(logits, 1234567890123, 7) == got).all()              # deterministic
    assert (_sample(logits, 1234567890123, 8) != got).any() or n == 1    # the offset changes the stream


def test_sampler_distribution_is_softmax():
    rng = np.random.RandomState(0)
    n, rows = 6, 200000
    base = rng.randn(n).astype(np.float32) * 1.5
    logits = np.tile(base, (rows, 1))
    idx = _sample(logits, seed=99, offset=0)
    freq = np.bincount(idx, minlength=n) / rows
    p = np.exp(base - base.max())
    p /= p.sum()
    # chi-square with n-1 = 5 degrees of freedom: 99.9 % quantile = 20.5
    chi2 = rows * ((freq - p) ** 2 / p).sum()
    assert chi2 < 20.5, (chi2, freq, p)


def test_device_offset_counter_equals_host_offset():
    """The device-resident call counter (CUDA-graph replays) addresses the same Philox stream as the host offset."""
    N = _sp()._native
    rng = np.random.RandomState(3)
    logits = (rng.randn(40, 512) * 2).astype(np.float32)
    lg = torch.tensor(logits, device="cuda")
    out = torch.empty((40,), dtype=torch.int32, device="cuda")
    ctr = torch.tensor([5], dtype=torch.int64, device="cuda")
    N.check(N.lib().spiral_categorical_sample(N.ptr(lg), 40, 512, C.c_uint64(11), C.c_uint64(2), N.ptr(ctr), N.ptr(out),
                                              N.stream_ptr()))
    assert (out.cpu().numpy() == _sample(logits, 11, 7)).all()
