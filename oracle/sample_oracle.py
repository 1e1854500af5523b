"""oracle/sample_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), NOT product code.

numpy restatement of the categorical sampler behind the policy's action heads
(models/policy.py:364-368: ``tf.multinomial(logits - max, 1)``).  TensorFlow's kernel is the Gumbel-max
trick over Philox4x32-10 uniforms; the op-seeded Philox schedule of TF 1.6 cannot be reproduced
outside TF, so what is pinned here is (1) the Philox4x32-10 block function against the Random123
known-answer vectors and (2) the sampler's distribution (softmax of the logits).
PARITY UNPINNED against TensorFlow's sample streams (by construction).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """ctr: uint32 [..., 4], key: uint32 [..., 2] -> uint32 [..., 4]  (Random123 philox4x32-10)."""
    c = [np.asarray(ctr[..., i], np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], np.uint64)
    k1 = np.asarray(key[..., 1], np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return np.stack(c, -1).astype(np.uint32)


def uniforms(rows, n, seed, offset):
    """The kernel's u[row, i]: counter = (i // 4, row, offset_lo, offset_hi), key = (seed_lo, seed_hi)."""
    nq = (n + 3) // 4
    ctr = np.zeros((rows, nq, 4), np.uint32)
    ctr[..., 0] = np.arange(nq, dtype=np.uint32)[None, :]
    ctr[..., 1] = np.arange(rows, dtype=np.uint32)[:, None]
    ctr[..., 2] = np.uint32(offset & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32((offset >> 32) & 0xFFFFFFFF)
    key = np.zeros((rows, nq, 2), np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key).reshape(rows, nq * 4)[:, :n]
    return ((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def categorical_sample(logits, seed, offset):
    """Gumbel-max: argmax_i logit_i - log(-log(u_i)), ties to the lowest index."""
    logits = np.asarray(logits, np.float32)
    u = uniforms(logits.shape[0], logits.shape[1], seed, offset)
    v = logits - np.log(-np.log(u)).astype(np.float32)
    return np.argmax(v, axis=1).astype(np.int32)
