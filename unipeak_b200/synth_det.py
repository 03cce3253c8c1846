"""Bench-scale synthetic data that is BIT-IDENTICAL on the CPU and on the GPU — set-up only, never timed.

Same recipe as unipeak_b200.synth.make_synthetic (70 % peak tags around shared centres with Pareto weights, 30 % uniform
background; directional: independent peak sets per strand; non-directional: fragment midpoints, '+' tag at mid - len/2 and
'-' tag at mid + len/2, true half-shift 90), but every random number comes from a counter-based integer hash (splitmix64)
evaluated with torch int64 operations, and every per-tag quantity is integer arithmetic: the same call gives the same arrays
with device="cpu" and device="cuda:0".  That is what lets the full-size configurations (16 x 30 M tags on hg19) be run
through the compiled reference in the CPU container once (tools/pin_fullsize.py -> tests/golden/fullsize_*.json) and compared
with the GPU result on the B200 box, where the inputs are regenerated instead of shipped.

The only floating point is the Pareto weight table of the peak centres (numpy on the host, ~3 x 10^5 values, always computed on
the CPU whatever `device` is) which is turned into an integer cumulative table before any tag is drawn.
Normal deviates are Irwin-Hall sums of four 16-bit uniforms (sd ~ 39 / 29.3 / 9.8 instead of 40 / 30 / 10).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .synth import TagData


def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


_GOLD, _M1, _M2 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (two's-complement wrap-around on both back ends)."""
    z = x + _GOLD
    z = (z ^ _lsr(z, 30)) * _M1
    z = (z ^ _lsr(z, 27)) * _M2
    return z ^ _lsr(z, 31)


def _mix_int(x: int) -> int:
    m = (1 << 64) - 1
    z = (x + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


def _key(*parts: int) -> int:
    k = 0x243F6A8885A308D3
    for p in parts:
        k = _mix_int(k ^ (p & ((1 << 64) - 1)))
    return k


def _bits(key: int, n: int, dev) -> torch.Tensor:
    """n 64-bit hash values of stream `key` (int64 tensor, all 64 bits random)."""
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    return _mix(idx * _GOLD + _s64(key))


def _uniform(bits: torch.Tensor, N: int) -> torch.Tensor:
    """Integers in [0, N), N <= 2^32, from the top 62 bits of `bits`: floor(r62 * N / 2^62) without leaving int64."""
    r = _lsr(bits, 2)
    hi, lo = r >> 31, r & ((1 << 31) - 1)
    return (hi * N + ((lo * N) >> 31)) >> 31


def _normal(bits: torch.Tensor, mult: int) -> torch.Tensor:
    """Irwin-Hall(4) of the four 16-bit fields, centred, scaled by mult / 2^20 and rounded to the nearest integer."""
    s = (bits & 0xFFFF) + (_lsr(bits, 16) & 0xFFFF) + (_lsr(bits, 32) & 0xFFFF) + _lsr(bits, 48) - 131070
    return (s * mult + (1 << 19)) >> 20


SD40, SD30, SD10 = 1083, 812, 271          # 37837 * mult / 2^20  ->  39.1, 29.3, 9.8


def _peak_set(seed: int, which: int, G: int, npk: int, dev):
    cpu = torch.device("cpu")
    g = _uniform(_bits(_key(seed, 101, which), npk, cpu), G)
    r53 = _lsr(_bits(_key(seed, 102, which), npk, cpu), 11).numpy().astype(np.float64)
    u = (r53 + 0.5) / 9007199254740992.0
    w = np.power(u, -1.0 / 1.2)                                        # Pareto(1.2) + 1
    cw = np.cumsum(w)
    cwi = np.floor(cw / cw[-1] * 4611686018427387904.0).astype(np.int64)   # scaled to 2^62
    cwi[-1] = 1 << 62
    return g.to(dev), torch.from_numpy(cwi).to(dev)


def make_synthetic_det(contig_sizes: Sequence[int], n_samples: int, tags_per_sample: int, *, directional: bool = True,
                       seed: int = 1234, bandwidth: int = 50, peak_fraction: float = 0.7, peak_spacing: int = 10000,
                       contigs: Sequence[int] | None = None, device: str = "cpu", contig_names=None, first_sample: int = 0) -> TagData:
    """Returns host (numpy) arrays.  `contigs`: keep only these contigs' tags (a rank's shard); the other segments are left
    empty so contig indices stay global — the kept segments are exactly those of the unsharded call.  `first_sample`: the
    samples generated are numbers first_sample .. first_sample + n_samples - 1 of the (unbounded) series this seed defines — the
    same arrays a larger call would hold for them (a rank that holds a block of samples)."""
    dev = torch.device(device)
    sizes = torch.tensor(list(contig_sizes), dtype=torch.int64, device=dev)
    C = int(sizes.shape[0])
    starts = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(sizes, 0)])
    G = int(starts[-1].item())
    assert G <= (1 << 32)
    keep_mask = torch.zeros(C, dtype=torch.bool, device=dev)
    keep_mask[list(range(C)) if contigs is None else list(contigs)] = True
    npk = max(1, G // peak_spacing)
    peak_sets = [_peak_set(seed, k, G, npk, dev) for k in range(2 if directional else 1)]

    pos_parts, cnt_parts = [], []
    off = np.zeros(n_samples * 2 * (C + 1), dtype=np.uint64)
    total = 0
    n_peak = int(round(tags_per_sample * peak_fraction))
    for s_local in range(n_samples):
        s = s_local + first_sample                                      # sample number in the seed's series (keys); s_local: slot in the arrays
        glob = [[], []]
        n_bg = tags_per_sample - n_peak
        if directional:
            for d in range(2):
                m = n_peak // 2 if d == 0 else n_peak - n_peak // 2
                g, cw = peak_sets[d]
                which = torch.searchsorted(cw, _lsr(_bits(_key(seed, 1, s, d), m, dev), 2), right=True).clamp_(max=npk - 1)
                glob[d].append(g[which] + _normal(_bits(_key(seed, 2, s, d), m, dev), SD40))
                del which
        else:
            g, cw = peak_sets[0]
            m = n_peak // 2
            which = torch.searchsorted(cw, _lsr(_bits(_key(seed, 1, s, 0), m, dev), 2), right=True).clamp_(max=npk - 1)
            mid = g[which] + _normal(_bits(_key(seed, 2, s, 0), m, dev), SD30)
            half = 90 + _normal(_bits(_key(seed, 3, s, 0), m, dev), SD10)
            glob[0].append(mid - half)
            glob[1].append(mid + half)
            n_bg = tags_per_sample - 2 * m
            del which, mid, half
        for d in range(2):
            m = n_bg // 2 if d == 0 else n_bg - n_bg // 2
            glob[d].append(_uniform(_bits(_key(seed, 4, s, d), m, dev), G))
        for d in range(2):
            x = torch.cat(glob[d])
            glob[d] = None
            x = x[(x >= 0) & (x < G)]
            c = torch.searchsorted(starts, x, right=True) - 1
            local = x - starts[c] + 1
            ok = (local >= bandwidth + 1) & (local <= sizes[c] - bandwidth) & keep_mask[c]
            key = (c[ok] << 32) | local[ok]
            del x, c, local, ok
            key, _ = torch.sort(key)
            uk, k = torch.unique_consecutive(key, return_counts=True)
            del key
            uc = uk >> 32
            bounds = torch.searchsorted(uc, torch.arange(C + 1, device=dev, dtype=torch.int64)).cpu().numpy().astype(np.uint64)
            base = (s_local * 2 + d) * (C + 1)
            off[base:base + C + 1] = bounds + np.uint64(total)
            pos_parts.append((uk & 0xFFFFFFFF).cpu().numpy().astype(np.uint32))
            cnt_parts.append(k.cpu().numpy().astype(np.uint32))
            total += int(uk.shape[0])
            del uk, k, uc
    pos = np.concatenate(pos_parts) if pos_parts else np.zeros(0, np.uint32)
    cnt = np.concatenate(cnt_parts) if cnt_parts else np.zeros(0, np.uint32)
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    return TagData(np.asarray(list(contig_sizes), dtype=np.uint32), n_samples, off, pos, cnt, list(contig_names) if contig_names else [])


def fingerprint(data: TagData) -> str:
    """SHA-256 over the offsets, positions and counts (what the GPU box recomputes to prove it holds the same inputs)."""
    import hashlib
    h = hashlib.sha256()
    for a in (data.contig_sizes.astype("<u4"), data.off.astype("<u8"), data.pos.astype("<u4"), data.cnt.astype("<u4")):
        h.update(np.ascontiguousarray(a).view(np.uint8).tobytes())
    return h.hexdigest()
