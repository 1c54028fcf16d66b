"""Deterministic synthetic 16 kHz audio for tests, golden fixtures and bench.

Everything is generated with integer hashes / float64 closed forms and then
quantised to int16, i.e. exactly the value set the reference's dataset yields
(``audio.float().div_(32768.0)``, model/dataset.py:62), and does not depend on
any framework RNG stream.
"""
import numpy as np

SEGMENT_SAMPLES = 327680          # train_UNet_VAT.py:55  sequence_length
SAMPLE_RATE = 16000               # model/constants.py:4
SEGMENT_SECONDS = SEGMENT_SAMPLES / SAMPLE_RATE   # 20.48


def _splitmix64(idx, seed):
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x632BE59BD9B4E019)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def white_int16(n, seed):
    """Uniform int16 noise in [-32768, 32767]."""
    h = _splitmix64(np.arange(n, dtype=np.uint64), seed)
    return (h >> np.uint64(48)).astype(np.uint16).view(np.int16).copy()


def uniform01(n, seed):
    """float64 uniforms in [0,1) from the same hash (for phases / pitches)."""
    h = _splitmix64(np.arange(n, dtype=np.uint64), seed)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def music_int16(n, seed, n_notes=8, noise_db=-40.0, gain=0.95):
    """Sum of harmonic stacks at piano pitches (MIDI 21..108) with 1/h
    amplitudes plus white noise `noise_db` below the signal RMS: strong
    partials next to weak bins, the hard case for the STFT's dynamic range."""
    u = uniform01(64 * n_notes + 8, seed * 7919 + 13)
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    x = np.zeros(n, dtype=np.float64)
    for j in range(n_notes):
        midi = 21 + int(u[64 * j] * 88)
        f0 = 440.0 * 2.0 ** ((midi - 69) / 12.0)
        h = 1
        while f0 * h < 0.5 * SAMPLE_RATE and h < 32:
            x += np.sin(2 * np.pi * f0 * h * t + 2 * np.pi * u[64 * j + h]) / h
            h += 1
    x *= gain / np.abs(x).max()
    rms = np.sqrt(np.mean(x * x))
    noise = white_int16(n, seed + 1000003).astype(np.float64) / 32768.0 * np.sqrt(3.0)   # unit variance
    x += noise * rms * 10.0 ** (noise_db / 20.0)
    return np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)


def impulse_int16(n, pos, amp=16384):
    x = np.zeros(n, dtype=np.int16)
    x[pos] = amp
    return x


def to_float(x_int16):
    """model/dataset.py:62"""
    return x_int16.astype(np.float32) / np.float32(32768.0)


def segments(batch, kind="mixed", seed=0, n=SEGMENT_SAMPLES):
    """(batch, n) float32 in [-1, 1).  kind: 'white' | 'music' | 'mixed'."""
    out = np.empty((batch, n), dtype=np.float32)
    for b in range(batch):
        k = kind if kind != "mixed" else ("music" if b % 2 else "white")
        out[b] = to_float(white_int16(n, seed + b) if k == "white" else music_int16(n, seed + b))
    return out
