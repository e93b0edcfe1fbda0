"""Synthetic founders of the BASELINE.json shapes for bench.py: n_chr linkage
groups of evenly spaced markers, per-marker allele frequency ~ U(0.05, 0.95), phased 'A'/'T'
founders, additive effects for both alleles ~ N(0,1).  Same dict layout as the test datasets."""
import numpy as np


def founders(n_founders, n_markers, n_chr, chr_len_cm=150.0, seed=2, n_effect_sets=1):
    rng = np.random.default_rng(seed)
    base = n_markers // n_chr
    sizes = np.full(n_chr, base, dtype=np.int64)
    sizes[n_chr - (n_markers - base * n_chr):] += 1
    chr_offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    dists = np.zeros(n_markers)
    pos = np.zeros(n_markers)
    lam = np.zeros(n_chr)
    for c in range(n_chr):
        lo, hi = int(chr_offset[c]), int(chr_offset[c + 1])
        n = hi - lo
        p = np.arange(n, dtype=np.float64) * (chr_len_cm / max(n - 1, 1))
        span = p[-1] - p[0]
        pos[lo:hi] = p
        lam[c] = span / 100.0                       # expected crossovers = Morgans
        with np.errstate(invalid="ignore", divide="ignore"):
            dists[lo:hi] = (p - p[0]) / span
    sym = np.frombuffer(b"AT", dtype=np.uint8)
    freq = rng.uniform(0.05, 0.95, n_markers)
    g = np.empty((n_founders, 2 * n_markers), dtype=np.uint8)
    rep = np.repeat(freq, 2)
    for i in range(n_founders):                     # row by row: bounded temporary memory
        g[i] = np.where(rng.random(2 * n_markers) < rep, sym[0], sym[1])
    effects = []
    for _ in range(n_effect_sets):
        effects.append(dict(cumn=(np.arange(1, n_markers + 1, dtype=np.uint32) * 2).astype(np.uint32),
                            allele=np.tile(sym, n_markers).astype(np.uint8),
                            eff=rng.standard_normal(2 * n_markers), centre=None))
    width = len(str(n_chr))
    mwidth = len(str(n_markers))
    mp = dict(lam=lam, chr_offset=chr_offset, dists=dists, marker_index=np.arange(n_markers, dtype=np.uint32),
              has_dists=np.ones(n_chr, dtype=np.uint8), pos_cm=pos,
              chr_names=["c%0*d" % (width, c + 1) for c in range(n_chr)])
    return dict(n_markers=n_markers, genotypes=g, names=["F%d" % (i + 1) for i in range(n_founders)],
                marker_names=["m%0*d" % (mwidth, m + 1) for m in range(n_markers)], maps=[mp], effects=effects)
