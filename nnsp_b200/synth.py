"""Deterministic synthetic 16 kHz PCM for parity tests and benchmarks (no audio files needed).

Stream ``s`` is built from a small pool of speech-like "programmes" (voiced harmonic segments,
noise bursts, pauses), read at a per-stream offset, scaled by a per-stream gain and mixed with
per-stream uniform noise. Every 16th stream is an edge case: full-range noise, digital silence,
a +-12000 square wave at 1 kHz, or a constant DC level.
Only numpy; the same call returns the same samples everywhere.
"""
import numpy as np

FRAME = 160
_SINE = np.round(32767.0 * np.sin(2.0 * np.pi * np.arange(4096) / 4096.0)).astype(np.int64)


def _programme(rng, n):
    out = np.zeros(n, np.int64)
    pos = 0
    while pos < n:
        seg = int(rng.integers(800, 6400))
        kind = int(rng.integers(0, 10))
        m = min(seg, n - pos)
        t = np.arange(m, dtype=np.int64)
        env = _SINE[((t * 2048) // max(m, 1)) % 4096]          # half-sine envelope, 0..32767
        if kind < 5:                                            # voiced: harmonics of f0
            f0 = int(rng.integers(90, 260))
            sig = np.zeros(m, np.int64)
            for h, w in ((1, 16), (2, 10), (3, 12), (5, 6), (8, 4), (13, 3)):
                step = (f0 * h * 4096) // 16000
                sig += w * _SINE[(t * step + int(rng.integers(0, 4096))) % 4096]
            sig = (sig * env) >> 21
        elif kind < 7:                                          # unvoiced burst
            sig = (rng.integers(-12000, 12000, m) * env) >> 15
        else:                                                   # pause
            sig = np.zeros(m, np.int64)
        out[pos:pos + m] = sig
        pos += m
    return out


def synth_pcm(n_streams, n_frames, seed=0x6E6E7370, first_stream=0, pool=32):
    """int16 array [n_streams, n_frames*160]; stream index = first_stream + row."""
    n = n_frames * FRAME
    rng = np.random.default_rng(seed)
    plen = n + 16000
    progs = np.stack([_programme(rng, plen) for _ in range(pool)])
    out = np.empty((n_streams, n), np.int16)
    for r in range(n_streams):
        s = first_stream + r
        srng = np.random.default_rng([seed, s])
        cls = s % 16
        if cls == 0:
            x = srng.integers(-32768, 32768, n)
        elif cls == 1:
            x = np.zeros(n, np.int64)
        elif cls == 2:
            x = np.where((np.arange(n) // 8) % 2 == 0, 12000, -12000)
        elif cls == 3:
            x = np.full(n, 3000 + 500 * (s % 7), np.int64) + srng.integers(-4, 5, n)
        else:
            off = (s * 7919 * FRAME) % 16000
            base = progs[s % pool, off:off + n]
            amp = 1 << (4 + ((s >> 2) % 8))
            x = (base >> (s % 4)) + srng.integers(-amp, amp + 1, n)
        out[r] = np.clip(x, -32768, 32767).astype(np.int16)
    return out


def adversarial_windows():
    """480-sample analysis windows that stress the fixed-point front end (full scale, DC, alternating)."""
    n = 480
    w = [np.full(n, 32767), np.full(n, -32768), np.zeros(n), np.where(np.arange(n) % 2 == 0, 32767, -32768),
         np.where(np.arange(n) % 2 == 0, -32768, 32767), np.where((np.arange(n) // 2) % 2 == 0, 32767, -32768),
         np.concatenate([np.full(240, 32767), np.full(240, -32768)]), np.arange(n) * 136 - 32640,
         np.where(np.arange(n) == 200, 32767, 0), np.where(np.arange(n) % 4 < 2, -32768, 32767)]
    rng = np.random.default_rng(7)
    for k in range(6):
        w.append(rng.integers(-32768, 32768, n))
    for k in range(4):
        f = [1, 17, 128, 255][k]
        w.append(np.round(32767 * np.cos(2 * np.pi * f * np.arange(n) / 512.0)))
    return np.stack(w).astype(np.int16)
