"""Bench-scale synthetic data generated with torch on the GPU — set-up only, never timed.

Same recipe as unipeak_b200.synth.make_synthetic (70 % peak tags around shared centres with Pareto weights,
30 % uniform background; directional: independent peak sets per strand), but vectorised so that 16 x 30 M
tags on hg19 take seconds instead of minutes.  The random streams differ from the numpy generator's; parity
checks always run both sides on the SAME generated arrays.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .synth import TagData


def make_synthetic_gpu(contig_sizes: Sequence[int], n_samples: int, tags_per_sample: int, *, directional: bool = True,
                       seed: int = 1234, bandwidth: int = 50, peak_fraction: float = 0.7, peak_spacing: int = 10000,
                       contigs: Sequence[int] | None = None, device: str = "cuda:0", contig_names=None) -> TagData:
    """Returns host (numpy) arrays.  `contigs`: keep only these contigs' tags (a rank's shard); the other
    segments are left empty so contig indices stay global."""
    dev = torch.device(device)
    sizes = torch.tensor(list(contig_sizes), dtype=torch.int64, device=dev)
    C = int(sizes.shape[0])
    starts = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(sizes, 0)])
    G = int(starts[-1].item())
    keep_mask = torch.zeros(C, dtype=torch.bool, device=dev)
    keep_mask[list(range(C)) if contigs is None else list(contigs)] = True

    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    npk = max(1, G // peak_spacing)

    def peak_set():
        g = torch.randint(0, G, (npk,), generator=gen, device=dev, dtype=torch.int64)
        u = torch.rand(npk, generator=gen, device=dev, dtype=torch.float64)
        w = (1.0 - u).pow(-1.0 / 1.2)            # Pareto(1.2) + 1
        cw = torch.cumsum(w / w.sum(), 0)
        return g, cw

    peak_sets = [peak_set(), peak_set()] if directional else [peak_set()]

    pos_parts, cnt_parts = [], []
    off = np.zeros(n_samples * 2 * (C + 1), dtype=np.uint64)
    total = 0
    for s in range(n_samples):
        gs = torch.Generator(device=dev)
        gs.manual_seed(seed + 1 + s)
        n_peak = int(round(tags_per_sample * peak_fraction))
        n_bg = tags_per_sample - n_peak
        glob = [[], []]
        if directional:
            for d in range(2):
                m = n_peak // 2 if d == 0 else n_peak - n_peak // 2
                g, cw = peak_sets[d]
                which = torch.searchsorted(cw, torch.rand(m, generator=gs, device=dev, dtype=torch.float64)).clamp_(max=npk - 1)
                glob[d].append(torch.round(g[which].to(torch.float64) + 40.0 * torch.randn(m, generator=gs, device=dev, dtype=torch.float64)).to(torch.int64))
        else:
            g, cw = peak_sets[0]
            m = n_peak // 2
            which = torch.searchsorted(cw, torch.rand(m, generator=gs, device=dev, dtype=torch.float64)).clamp_(max=npk - 1)
            mid = g[which].to(torch.float64) + 30.0 * torch.randn(m, generator=gs, device=dev, dtype=torch.float64)
            ln = 180.0 + 20.0 * torch.randn(m, generator=gs, device=dev, dtype=torch.float64)
            glob[0].append(torch.round(mid - ln / 2).to(torch.int64))
            glob[1].append(torch.round(mid + ln / 2).to(torch.int64))
            n_bg = tags_per_sample - 2 * m
        for d in range(2):
            m = n_bg // 2 if d == 0 else n_bg - n_bg // 2
            glob[d].append(torch.randint(0, G, (m,), generator=gs, device=dev, dtype=torch.int64))
        for d in range(2):
            x = torch.cat(glob[d])
            x = x[(x >= 0) & (x < G)]
            c = torch.searchsorted(starts, x, right=True) - 1
            local = x - starts[c] + 1
            ok = (local >= bandwidth + 1) & (local <= sizes[c] - bandwidth) & keep_mask[c]
            key = (c[ok] << 32) | local[ok]
            key, _ = torch.sort(key)
            uk, k = torch.unique_consecutive(key, return_counts=True)
            uc = uk >> 32
            bounds = torch.searchsorted(uc, torch.arange(C + 1, device=dev, dtype=torch.int64)).cpu().numpy().astype(np.uint64)
            base = (s * 2 + d) * (C + 1)
            off[base:base + C + 1] = bounds + np.uint64(total)
            pos_parts.append((uk & 0xFFFFFFFF).to(torch.int64).cpu().numpy().astype(np.uint32))
            cnt_parts.append(k.cpu().numpy().astype(np.uint32))
            total += int(uk.shape[0])
            del x, c, local, ok, key, uk, k, uc
    pos = np.concatenate(pos_parts) if pos_parts else np.zeros(0, np.uint32)
    cnt = np.concatenate(cnt_parts) if cnt_parts else np.zeros(0, np.uint32)
    torch.cuda.empty_cache()
    return TagData(np.asarray(list(contig_sizes), dtype=np.uint32), n_samples, off, pos, cnt, list(contig_names) if contig_names else [])
