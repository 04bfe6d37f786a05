"""CPU ORACLE — test infrastructure only. numpy restatement of the production eps stream.

The reference draws eps with torch's global RNG (`eps.data.normal_()`, bayesian-torch 0.5.0
*_variational.py forward); that stream cannot be reproduced inside a kernel, so the CUDA path
defines its own: Philox4x32-10 (Salmon et al., SC'11 — published algorithm, constants below)
followed by Box-Muller. This file restates that spec so tests can regenerate, on the CPU, the
exact eps the kernels used and feed it to oracle/bnn_oracle.py.

  counter = (elem//4 low 32, elem//4 high 32, sample_id, layer_id)   key = (seed low, seed high)
  (r0,r1,r2,r3) -> u_i = ((r_i >> 8) + 0.5) / 2^24
  z0 = sqrt(-2 ln u0) cos(2 pi u1), z1 = .. sin(..), z2/z3 likewise from (u2,u3); eps[elem] = z[elem % 4]
  bias tensors use layer_id | 0x80000000.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        with np.errstate(over="ignore"):
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_normal(n: int, seed: int, layer_id: int, sample_id: int) -> np.ndarray:
    """float32[n]: mauv_philox_normal_f32 with exact libm (the device evaluates log / sqrt / sin / cos on the SFU: the two
    agree to ~1e-6 absolute, checked by tests/gpu_bringup.py t_philox at 1e-5 of the maximum)."""
    quads = (n + 3) // 4
    q = np.arange(quads, dtype=np.uint64)
    c0 = (q & MASK).astype(np.uint32)
    c1 = (q >> np.uint64(32)).astype(np.uint32)
    c2 = np.full(quads, sample_id & 0xFFFFFFFF, dtype=np.uint32)
    c3 = np.full(quads, layer_id & 0xFFFFFFFF, dtype=np.uint32)
    r = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    k = np.float32(1.0 / 16777216.0)
    u = [((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * k for x in r]
    ra = np.sqrt(np.float32(-2.0) * np.log(u[0]))
    rb = np.sqrt(np.float32(-2.0) * np.log(u[2]))
    two_pi = np.float32(2.0 * np.pi)
    z = np.stack([ra * np.cos(two_pi * u[1]), ra * np.sin(two_pi * u[1]),
                  rb * np.cos(two_pi * u[3]), rb * np.sin(two_pi * u[3])], axis=1).astype(np.float32)
    return z.reshape(-1)[:n]
