"""Synthetic "msprime-style" genotypes for benchmarks and large parity tests.

msprime / tskit are not available offline, so this is a self-contained two-population
generator with fixed seeds:

* positions: sorted unique integers in [1, L], density theta * a_H per bp
  (theta = 5.2e-4, a_H = harmonic number of the total haplotype count - 1);
* a fraction ``f_private`` of sites is private to the target population (reference
  haplotypes all ancestral) with a steep allele-count spectrum ~ k^-1.5, the rest are
  shared with a neutral 1/k spectrum over all haplotypes; carriers are drawn uniformly;
* diploid dosage = sum of consecutive haplotype pairs (sstar/utils/utils.py:305-306);
* no missing data; sites fixed-alt in both populations cannot occur (k <= H - 1).

On the 50 kb / 10 kb tiling this gives ~160 sites per window of which ~40 % are
reference-absent and n ~ 8 carried reference-absent sites per diploid target column,
matching tests/exp_results/test.0.vcf of the reference (88 of 228 sites absent, n ~ 7.8).
"""
from __future__ import annotations

import numpy as np

THETA = 5.2e-4


def _sfs_counts(rng, size, kmax, power):
    k = np.arange(1, kmax + 1, dtype=np.float64)
    w = k ** (-power)
    return rng.choice(np.arange(1, kmax + 1), size=size, p=w / w.sum())


def synth_positions(rng, length: int, n_sites: int) -> np.ndarray:
    n_sites = min(n_sites, length)
    draw = rng.integers(1, length + 1, size=int(n_sites * 1.05) + 16)
    pos = np.unique(draw)
    while pos.size < n_sites:
        pos = np.unique(np.concatenate([pos, rng.integers(1, length + 1, size=n_sites)]))
    keep = np.sort(rng.choice(pos.size, size=n_sites, replace=False))
    return pos[keep].astype(np.int32)


def synth_chromosome(length: int, n_ref: int, n_tgt: int, phased: bool, seed: int,
                     f_private: float = 0.38, ploidy: int = 2, density: float | None = None,
                     chunk: int = 1 << 16):
    """-> (ref_gts int8 [V,R], tgt_gts int8 [V,T], pos int32 [V]).

    R, T = individuals (unphased dosages) or individuals * ploidy (phased haplotypes).
    """
    rng = np.random.default_rng(seed)
    hr, ht = n_ref * ploidy, n_tgt * ploidy
    h = hr + ht
    if density is None:
        density = THETA * float(np.sum(1.0 / np.arange(1, h)))
    n_sites = int(rng.poisson(density * length))
    pos = synth_positions(rng, length, max(n_sites, 1))
    V = pos.size
    private = rng.random(V) < f_private
    k = np.where(private, _sfs_counts(rng, V, max(ht - 1, 1), 1.5), _sfs_counts(rng, V, h - 1, 1.0))
    ref_h = np.zeros((V, hr), dtype=np.int8)
    tgt_h = np.zeros((V, ht), dtype=np.int8)
    for a in range(0, V, chunk):  # carriers: the k smallest of H uniform keys per site
        b = min(V, a + chunk)
        pr = private[a:b]
        keys = rng.random((b - a, h), dtype=np.float32)
        keys[pr, :hr] = 2.0  # private sites never land on a reference haplotype
        order = np.argsort(keys, axis=1, kind="stable")
        rank = np.empty_like(order)
        np.put_along_axis(rank, order, np.arange(h)[None, :].repeat(b - a, 0), axis=1)
        carr = rank < k[a:b, None]
        ref_h[a:b] = carr[:, :hr]
        tgt_h[a:b] = carr[:, hr:]
    if phased:
        return ref_h, tgt_h, pos
    ref = ref_h.reshape(V, n_ref, ploidy).sum(axis=2, dtype=np.int8)
    tgt = tgt_h.reshape(V, n_tgt, ploidy).sum(axis=2, dtype=np.int8)
    return ref, tgt, pos


def synth_chromosome_device(torch, device, length: int, n_ref: int, n_tgt: int, phased: bool,
                            seed: int, f_private: float = 0.38, ploidy: int = 2,
                            density: float | None = None, chunk: int = 1 << 16):
    """Large-scale variant generated on the GPU (independent Bernoulli carriers per haplotype
    with the same per-site frequencies); returns device tensors (int8, int8, int32)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    hr, ht = n_ref * ploidy, n_tgt * ploidy
    h = hr + ht
    if density is None:
        density = THETA * float(np.sum(1.0 / np.arange(1, h)))
    n_sites = int(density * length)
    draw = torch.randint(1, length + 1, (int(n_sites * 1.03) + 16,), device=device, generator=g)
    pos = torch.unique(draw)[:n_sites].to(torch.int32)
    V = pos.numel()
    private = torch.rand(V, device=device, generator=g) < f_private

    def sfs(kmax, power):
        k = torch.arange(1, kmax + 1, device=device, dtype=torch.float64)
        w = k ** (-power)
        return torch.multinomial(w / w.sum(), V, replacement=True, generator=g) + 1

    k = torch.where(private, sfs(max(ht - 1, 1), 1.5), sfs(h - 1, 1.0)).to(torch.float32)
    f_ref = torch.where(private, torch.zeros_like(k), k / h)
    f_tgt = torch.where(private, k / ht, k / h)
    R = hr if phased else n_ref
    T = ht if phased else n_tgt
    ref = torch.empty((V, R), dtype=torch.int8, device=device)
    tgt = torch.empty((V, T), dtype=torch.int8, device=device)
    for a in range(0, V, chunk):
        b = min(V, a + chunk)
        rh = (torch.rand((b - a, hr), device=device, generator=g) < f_ref[a:b, None]).to(torch.int8)
        th = (torch.rand((b - a, ht), device=device, generator=g) < f_tgt[a:b, None]).to(torch.int8)
        if phased:
            ref[a:b], tgt[a:b] = rh, th
        else:
            ref[a:b] = rh.view(b - a, n_ref, ploidy).sum(dim=2, dtype=torch.int8)
            tgt[a:b] = th.view(b - a, n_tgt, ploidy).sum(dim=2, dtype=torch.int8)
    return ref, tgt, pos


def regular_windows(pos, win_len: int, win_step: int):
    """The reference's window rule (sstar/utils/utils.py:480-488) as int32 arrays."""
    first, last = int(pos[0]), int(pos[-1])
    start0 = max(1, (first + win_step) // win_step * win_step - win_len + 1)
    n = (last - start0) // win_step + 1 if last >= start0 else 0
    ws = start0 + win_step * np.arange(n, dtype=np.int64)
    return ws.astype(np.int32), (ws + win_len - 1).astype(np.int32)


def write_vcf(path: str, chrom: str, pos, hap, missing_rate: float = 0.0, seed: int = 0, sample_prefix: str = "ind"):
    """Phased haplotype matrix ``hap`` int8 [V, 2N] -> a plain-text VCF with N diploid samples
    (``0|1`` calls, ``.`` for alleles made missing at ``missing_rate``).  Returns the sample names.
    For tests and benchmarks of the ingest path."""
    hap = np.asarray(hap)
    V, H = hap.shape
    n = H // 2
    names = [f"{sample_prefix}{i}" for i in range(n)]
    rng = np.random.default_rng(seed)
    txt = np.where(hap == 0, "0", np.where(hap == 1, "1", "2")).astype("U1")
    if missing_rate > 0:
        txt[rng.random(hap.shape) < missing_rate] = "."
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n##source=sstar_b200.synth\n")
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
        for a in range(0, V, 4096):
            b = min(V, a + 4096)
            calls = np.char.add(np.char.add(txt[a:b, 0::2], "|"), txt[a:b, 1::2])
            rows = ["\t".join((chrom, str(int(p)), ".", "A", "T", ".", "PASS", ".", "GT") + tuple(c))
                    for p, c in zip(pos[a:b], calls)]
            fh.write("\n".join(rows) + "\n")
    return names
