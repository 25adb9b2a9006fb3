"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the engine's dropout RNG.

The reference's dropout (`Dropout(0.25)` T1H:863 etc.) uses TensorFlow's RNG stream, which
cannot be matched by any other implementation (hard part 6).  The engine
therefore defines its own counter-based stream -- Philox4x32-10 (Salmon et al., SC'11; the
published constants below) -- and this file restates it on the CPU so that a training step
*with dropout on* can still be compared bit-for-bit on the mask.

Stream definition (shared with csrc/common.cuh `philox_keep`):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (e >> 2  [low 32 bits], step [low 32 bits], op_id, e >> 34)
    where e is the logical NHWC element index of the dropout output tensor.
    word    = philox4x32_10(counter, key)[e & 3]
    keep    = word >= floor(float32(p) * 2**32)   (P[keep] = 1 - p)
Kept elements are scaled by 1/(1-p) (Keras `Dropout`, training only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint64(0x9E3779B9)
_W1 = np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs uint64 arrays/scalars holding 32-bit values."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64) + np.zeros_like(c0)
    c2 = np.asarray(c2, dtype=np.uint64) + np.zeros_like(c0)
    c3 = np.asarray(c3, dtype=np.uint64) + np.zeros_like(c0)
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = (hi1 ^ c1 ^ k0) & _MASK
        n1 = lo1
        n2 = (hi0 ^ c3 ^ k1) & _MASK
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def dropout_keep_mask(n_elem, p, seed, step, op_id):
    """Boolean keep-mask of length n_elem (logical NHWC order)."""
    e = np.arange(n_elem, dtype=np.uint64)
    blk = e >> np.uint64(2)
    out = philox4x32_10(blk & _MASK, np.uint64(step) & _MASK, np.uint64(op_id), blk >> np.uint64(32),
                        np.uint64(seed) & _MASK, (np.uint64(seed) >> np.uint64(32)) & _MASK)
    lane = (e & np.uint64(3)).astype(np.int64)
    words = np.stack(out, axis=0)                      # (4, n)
    w = words[lane, np.arange(n_elem)]
    thr = np.uint64(int(np.floor(float(np.float32(p)) * 4294967296.0)))   # p is a C float on the device
    return w >= thr
