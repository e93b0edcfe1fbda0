"""oracle/datasets.py — TEST INFRASTRUCTURE ONLY.

Plain-numpy descriptions of simulation inputs ("datasets") that can be fed to
all three implementations: the reference (through text files, as its loaders
require), the C restatement (oracle/port.py) and the CUDA engine.

A dataset is a dict:
  n_markers   int
  genotypes   (n, 2M) uint8   alleles of marker m at [2m], [2m+1] (core.h:778-783)
  names       list[str]
  marker_names list[str]
  maps        list of dict(lam (C,) f64, chr_offset (C+1,) u32, dists (T,) f64,
                           marker_index (T,) u32, has_dists (C,) u8, chr_names, pos_cm (T,) f64)
  effects     list of dict(cumn (M,) u32, allele (nnz,) u8, eff (nnz,) f64, centre (M,) f64 | None)
"""
import os
import numpy as np


def chromosome_sizes(n_markers, n_chr):
    base = n_markers // n_chr
    sizes = np.full(n_chr, base, dtype=np.int64)
    sizes[n_chr - (n_markers - base * n_chr):] += 1   # remainder goes to the last chromosomes
    return sizes


def synthetic(n_founders, n_markers, n_chr, chr_len_cm=100.0, seed=1, spacing="even",
              alleles=(b"A", b"T"), n_effect_sets=1, missing_rate=0.0):
    """Synthetic founders of the BASELINE.json shapes: per-marker allele
    frequency ~ U(0.05, 0.95), phased; additive effects for every allele ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    sizes = chromosome_sizes(n_markers, n_chr)
    chr_offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    pos = np.zeros(n_markers, dtype=np.float64)
    for c in range(n_chr):
        n = int(sizes[c])
        if spacing == "even":
            p = np.arange(n, dtype=np.float64) * (chr_len_cm / max(n - 1, 1))
        else:
            p = np.sort(np.round(rng.uniform(0.0, chr_len_cm, n), 6))
        pos[chr_offset[c]:chr_offset[c + 1]] = p
    lam = np.zeros(n_chr, dtype=np.float64)
    dists = np.zeros(n_markers, dtype=np.float64)
    for c in range(n_chr):
        lo, hi = int(chr_offset[c]), int(chr_offset[c + 1])
        span = pos[hi - 1] - pos[lo]
        lam[c] = span / 100                       # core.c:5967
        with np.errstate(invalid="ignore", divide="ignore"):
            dists[lo:hi] = (pos[lo:hi] - pos[lo]) / span   # core.c:5929 (0/0 = NaN for one marker)
    width = len(str(n_chr))
    chr_names = ["c%0*d" % (width, c + 1) for c in range(n_chr)]
    mwidth = len(str(n_markers))
    marker_names = ["m%0*d" % (mwidth, m + 1) for m in range(n_markers)]

    sym = np.frombuffer(b"".join(alleles), dtype=np.uint8)
    na = len(sym)
    freq = rng.uniform(0.05, 0.95, n_markers)
    g = np.empty((n_founders, 2 * n_markers), dtype=np.uint8)
    u = rng.random((n_founders, 2 * n_markers))
    if na == 2:
        g[:] = np.where(u < np.repeat(freq, 2), sym[0], sym[1])
    else:
        g[:] = sym[np.minimum((u * na).astype(np.int64), na - 1)]
    if missing_rate > 0:
        g[rng.random(g.shape) < missing_rate] = 0
    names = ["F%d" % (i + 1) for i in range(n_founders)]

    effects = []
    for _ in range(n_effect_sets):
        cumn = (np.arange(1, n_markers + 1, dtype=np.uint32) * na).astype(np.uint32)
        allele = np.tile(sym, n_markers).astype(np.uint8)
        eff = rng.standard_normal(n_markers * na)
        effects.append(dict(cumn=cumn, allele=allele, eff=eff, centre=None))

    mp = dict(lam=lam, chr_offset=chr_offset, dists=dists,
              marker_index=np.arange(n_markers, dtype=np.uint32),
              has_dists=np.ones(n_chr, dtype=np.uint8), chr_names=chr_names, pos_cm=pos)
    return dict(n_markers=n_markers, genotypes=g, names=names, marker_names=marker_names,
                maps=[mp], effects=effects)


def write_files(ds, dirpath, tag="ds", effect_set=0):
    """Write genotype matrix / map / effect files the reference's loaders accept
    (formats: vignettes/gSvignette.Rmd "Input files"; loaders core.c:6229, 6333, 7085)."""
    os.makedirs(dirpath, exist_ok=True)
    M = ds["n_markers"]
    g = ds["genotypes"]
    geno = os.path.join(dirpath, tag + "-geno.txt")
    with open(geno, "wb") as f:
        f.write(b"name\t" + "\t".join(ds["marker_names"]).encode() + b"\n")
        pairs = g.reshape(g.shape[0], M, 2)
        for i, nm in enumerate(ds["names"]):
            row = np.empty((M, 3), dtype=np.uint8)
            row[:, 0] = ord("\t")
            row[:, 1:] = pairs[i]
            f.write(nm.encode() + row.tobytes() + b"\n")
    mp = ds["maps"][0]
    mapf = os.path.join(dirpath, tag + "-map.txt")
    with open(mapf, "w") as f:
        f.write("marker chr pos\n")
        for c in range(len(mp["lam"])):
            for p in range(int(mp["chr_offset"][c]), int(mp["chr_offset"][c + 1])):
                f.write("%s %s %s\n" % (ds["marker_names"][int(mp["marker_index"][p])],
                                        mp["chr_names"][c], repr(float(mp["pos_cm"][p]))))
    efff = None
    if ds["effects"]:
        e = ds["effects"][effect_set]
        efff = os.path.join(dirpath, tag + "-eff%d.txt" % effect_set)
        with open(efff, "w") as f:
            lo = 0
            for m in range(M):
                for k in range(lo, int(e["cumn"][m])):
                    f.write("%s %s %s\n" % (ds["marker_names"][m], chr(int(e["allele"][k])), repr(float(e["eff"][k]))))
                lo = int(e["cumn"][m])
    return geno, mapf, efff


def from_ref(refsim, group=0):
    """Export everything a RefSim holds into a dataset (maps/effects bit-exact)."""
    M = refsim.n_markers
    x = refsim.export_group(group)
    maps = []
    for mi in range(refsim.L.refb_n_maps(refsim.d)):
        chrs = refsim.export_map(mi)
        lam = np.array([c["lam"] for c in chrs], dtype=np.float64)
        sizes = np.array([c["n"] for c in chrs], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
        dists = np.concatenate([c["dists"] if c["dists"] is not None else np.full(c["n"], np.nan)
                                for c in chrs]) if chrs else np.zeros(0)
        mix = np.concatenate([c["marker_indexes"] for c in chrs]).astype(np.uint32) if chrs else np.zeros(0, np.uint32)
        has = np.array([c["dists"] is not None for c in chrs], dtype=np.uint8)
        maps.append(dict(lam=lam, chr_offset=off, dists=dists.astype(np.float64), marker_index=mix, has_dists=has))
    effects = [refsim.export_effects(i) for i in range(refsim.L.refb_n_eff_sets(refsim.d))]
    return dict(n_markers=M, genotypes=x["genotypes"], names=refsim.group_names(group),
                marker_names=refsim.marker_names(), maps=maps, effects=effects,
                ids=x["ids"], groups=x["groups"])


def to_jsonable(ds):
    out = {}
    for k, v in ds.items():
        if isinstance(v, np.ndarray):
            out[k] = v.tolist()
        elif isinstance(v, list) and v and isinstance(v[0], dict):
            out[k] = [{kk: (vv.tolist() if isinstance(vv, np.ndarray) else vv) for kk, vv in d.items()} for d in v]
        else:
            out[k] = v
    return out


def from_jsonable(js):
    ds = dict(js)
    ds["genotypes"] = np.array(js["genotypes"], dtype=np.uint8)
    ds["maps"] = [dict(lam=np.array(m["lam"], dtype=np.float64),
                       chr_offset=np.array(m["chr_offset"], dtype=np.uint32),
                       dists=np.array([np.nan if d is None else d for d in m["dists"]], dtype=np.float64),
                       marker_index=np.array(m["marker_index"], dtype=np.uint32),
                       has_dists=np.array(m["has_dists"], dtype=np.uint8)) for m in js["maps"]]
    ds["effects"] = [dict(cumn=np.array(e["cumn"], dtype=np.uint32),
                          allele=np.array(e["allele"], dtype=np.uint8),
                          eff=np.array(e["eff"], dtype=np.float64),
                          centre=None if e.get("centre") is None else np.array(e["centre"], dtype=np.float64))
                     for e in js["effects"]]
    return ds
