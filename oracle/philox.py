"""
NumPy Philox4x32-10 and the counter layout the CUDA kernels draw from.  TEST INFRASTRUCTURE.

Lets the oracle reproduce, bit for bit, the uniforms that `csrc/b2c_rng.cuh` generates on
the device, so the Philox (throughput) mode of the kernels is parity-checked exactly like
the injected mode: oracle(draws = philox draws) vs GPU(philox).

Philox4x32-10 is the counter-based generator of Salmon et al., "Parallel random numbers: as
easy as 1, 2, 3" (SC'11); known-answer vectors from the Random123 distribution are checked in
tests/test_philox.py.  The reference itself uses numpy's global MT19937 stream --
RNG parity with it is by injection only.

Counter layout (shared with include/b2c.h):
    key  = (seed & 0xffffffff, seed >> 32)
    ctr  = (index, stream, slot & 0xffffffff, slot >> 32)         slot = global sample index
    (bin k has frequency offset f = k - half (k < half) or k - half + 1, half = (nsc+1)//2;
     it uses lane l = |f| - 1 and word half h = (f > 0))
    stream 0 SYMBOLS: index = (s >> 1) * 320 + l ; word (s & 1)*2 + h     -> phase of RE (s, k)
    stream 1 JAKES  : index = ((p*ntx + tx)*nrx + rx)*10 + (n >> 1)
                      words (0,1) for even n, (2,3) for odd n            -> (angle, phase) of oscillator n
    stream 2 NOISE  : index = (s*nrx + rx)*320 + l ; words (2h, 2h+1)     -> Box-Muller (u1, u2) of rx[s, rx, k]
    stream 3 PARAMS : index = 0 ; words 0..3 -> model, doppler, snr, density choice: (word * n) >> 32
    uniform u = ((word >> 9) + 0.5) * 2**-23   (exact in fp32, never 0 or 1)
    noise   = sqrt(-2 ln u1) * (cos 2 pi u2 + j sin 2 pi u2)
"""

from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
STREAM_SYMBOLS, STREAM_JAKES, STREAM_NOISE, STREAM_PARAMS = 0, 1, 2, 3
RNG_LANES = 320


def philox4x32_10(ctr, key):
    """ctr[..., 4] uint32, key (k0, k1) -> out[..., 4] uint32."""
    c = np.asarray(ctr, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def u01(words):
    """23-bit uniform in (0,1): exactly what the device computes in fp32."""
    return ((np.asarray(words, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0 ** -23


def _block(seed, slot, stream, index):
    index = np.asarray(index, dtype=np.uint64)
    ctr = np.empty(index.shape + (4,), dtype=np.uint64)
    ctr[..., 0] = index
    ctr[..., 1] = stream
    ctr[..., 2] = int(slot) & 0xFFFFFFFF
    ctr[..., 3] = (int(slot) >> 32) & 0xFFFFFFFF
    return philox4x32_10(ctr, (int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF))


def _lane(k, nsc):
    """(lane, word half) of used bin k: mirror bins -f / +f share a Philox call."""
    half = (nsc + 1) // 2
    h = (k >= half).astype(np.int64)
    return np.where(h == 1, k - half, half - 1 - k), h


def symbol_u(seed, slot, nsym, nsc, qpsk=False):
    """u[nsym, nsc]: phase (in turns) of every resource element.  qpsk (b2c_slots.qpsk): the word keeps its top two
    bits (the quadrant) and the phase sits at the quadrant's centre, u = k/4 + 1/8 (+ 2^-24)."""
    s, k = np.meshgrid(np.arange(nsym), np.arange(nsc), indexing="ij")
    l, h = _lane(k, nsc)
    w = _block(seed, slot, STREAM_SYMBOLS, (s >> 1) * RNG_LANES + l)
    w = np.take_along_axis(w, ((s & 1) * 2 + h)[..., None], axis=-1)[..., 0]
    if qpsk:
        w = (w & np.uint32(0xC0000000)) | np.uint32(0x20000000)
    return u01(w)


def jakes_u(seed, slot, npaths, ntx, nrx, nosc=20):
    """u[P, ntx, nrx, 2, nosc] in the oracle's jakes_u layout (angles, phases)."""
    p, t, r, n = np.meshgrid(np.arange(npaths), np.arange(ntx), np.arange(nrx), np.arange(nosc), indexing="ij")
    w = _block(seed, slot, STREAM_JAKES, ((p * ntx + t) * nrx + r) * (nosc // 2) + (n >> 1))
    off = (n & 1) * 2
    ang = np.take_along_axis(w, off[..., None], axis=-1)[..., 0]
    ph = np.take_along_axis(w, (off + 1)[..., None], axis=-1)[..., 0]
    return np.stack([u01(ang), u01(ph)], axis=3)


def noise(seed, slot, nsym, nrx, nsc):
    """(re, im)[nsym, nrx, nsc] unit-variance-per-component Box-Muller normals."""
    s, r, k = np.meshgrid(np.arange(nsym), np.arange(nrx), np.arange(nsc), indexing="ij")
    l, h = _lane(k, nsc)
    w = _block(seed, slot, STREAM_NOISE, (s * nrx + r) * RNG_LANES + l)
    off = h * 2
    u1 = u01(np.take_along_axis(w, off[..., None], axis=-1)[..., 0])
    u2 = u01(np.take_along_axis(w, (off + 1)[..., None], axis=-1)[..., 0])
    rad = np.sqrt(-2.0 * np.log(u1))
    return rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)


def param_choice(seed, slot, n_model, n_doppler, n_snr, n_density):
    """Indices of the per-slot (model, doppler, snr, density) choice."""
    w = _block(seed, slot, STREAM_PARAMS, np.zeros((), dtype=np.uint64)).astype(np.uint64)
    return tuple(int((w[i] * np.uint64(n)) >> np.uint64(32)) for i, n in
                 enumerate((n_model, n_doppler, n_snr, n_density)))


def slot_draws(seed, slot, cfg_nsym, nsc, npaths, ntx, nrx, pilot_mask, qpsk=False):
    """Draw dictionary for oracle.simulate(): same keys as a recorded reference run, but
    `perm` is replaced by an explicit pilot mask (the pattern comes from the host pool)."""
    u = symbol_u(seed, slot, cfg_nsym, nsc, qpsk)
    ph = 2 * np.pi * u
    nre, nim = noise(seed, slot, cfg_nsym, nrx, nsc)
    return {"pilot_phase": ph[pilot_mask], "data_phase": ph[~pilot_mask],
            "jakes_u": jakes_u(seed, slot, npaths, ntx, nrx), "noise_re": nre, "noise_im": nim}
