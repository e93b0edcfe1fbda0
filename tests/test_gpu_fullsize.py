"""BASELINE.json configs[1] at full size on the GPU (100 000 offspring x 50 000 SNPs x 21 chr, native
draws), checked through size-independent properties — the oracle would need minutes for this:
mosaic-of-parents, GEBV linearity against a host fp64 sum on a sample, allele-count checksum of
checksums, selection sortedness, DH homozygosity, selfing halving heterozygosity."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M, CHR, F, N = 50000, 21, 200, 100000


@pytest.fixture(scope="module")
def world():
    from genomicsimulation_b200 import SimData, synthetic
    ds = synthetic.founders(F, M, CHR, 150.0, seed=2)
    sim = SimData.from_dataset(ds)
    sim.use_native_draws(20261019)
    f1 = sim.make_random_crosses(1, N)
    yield ds, sim, f1
    sim.close()


def host_bv(ds, geno):
    eff = ds["effects"][0]["eff"].reshape(M, 2)          # columns: 'A', 'T'
    isT0 = geno[:, 0::2] == ord("T")
    isT1 = geno[:, 1::2] == ord("T")
    return (np.where(isT0, eff[:, 1], eff[:, 0]) + np.where(isT1, eff[:, 1], eff[:, 0])).sum(axis=1)


def test_generation_properties(world):
    ds, sim, f1 = world
    assert sim.group_size(f1) == N
    meta = sim.export_group(f1, genotypes=False)
    assert meta["ids"].tolist() == list(range(F + 1, F + N + 1))
    assert (meta["ped1"] != meta["ped2"]).all() and meta["ped1"].min() >= 1 and meta["ped1"].max() <= F
    bv = sim.calculate_bvs(f1, 1)
    assert bv.shape == (N,) and np.isfinite(bv).all()

    # a sample of offspring, downloaded through the chars boundary
    rng = np.random.default_rng(1)
    pick = np.sort(rng.choice(N, 160, replace=False))
    g = sim.make_group_from(meta["indexes"][pick])
    x = sim.export_group(g)
    geno = x["genotypes"]
    founders = ds["genotypes"]
    p1, p2 = x["ped1"].astype(int) - 1, x["ped2"].astype(int) - 1
    a, b = founders[p1], founders[p2]
    assert np.all((geno[:, 0::2] == a[:, 0::2]) | (geno[:, 0::2] == a[:, 1::2]))     # strand 0 from parent 1
    assert np.all((geno[:, 1::2] == b[:, 0::2]) | (geno[:, 1::2] == b[:, 1::2]))     # strand 1 from parent 2
    # strand switches seen where the parent is heterozygous: sum(lambda) = 31.5 crossovers per gamete
    # plus a coin flip at each of the 20 chromosome boundaries = 41.5, minus the few that hide in
    # homozygous stretches
    het = a[:, 0::2] != a[:, 1::2]
    from1 = geno[:, 0::2] == a[:, 1::2]
    switches = []
    for i in range(len(pick)):
        s = from1[i][het[i]]
        switches.append(int((s[1:] != s[:-1]).sum()))
    assert 36 < np.mean(switches) < 44, np.mean(switches)
    # GEBV linearity: device result == host fp64 sum over the downloaded alleles
    want = host_bv(ds, geno)
    got = sim.calculate_bvs(g, 1)
    assert np.all(np.abs(got - want) <= 1e-9 * np.maximum(np.abs(want), np.abs(want).mean()))
    np.testing.assert_array_equal(got, bv[pick])                 # and the same individual gives the same bits
    # checksum of checksums: total 'T' count from the count kernel == from the downloaded alleles
    cnt = sim.calculate_allele_counts(g, "T")
    assert cnt.sum() == (geno == ord("T")).sum()
    assert np.array_equal(cnt, (geno[:, 0::2] == ord("T")).astype(float) + (geno[:, 1::2] == ord("T")))


def test_selection_sortedness(world):
    ds, sim, f1 = world
    idx_all = sim.group_indexes(f1)
    n_before = len(idx_all)
    bv = sim.calculate_bvs(f1, 1)
    top = sim.split_by_bv(f1, 1, N // 100, False)                # select.by.gebv top 1 %
    sel = sim.group_indexes(top)
    assert len(sel) == N // 100
    pos = np.searchsorted(idx_all, sel)
    cut = bv[pos].min()
    rest = np.delete(bv, pos)
    assert rest.max() <= cut
    assert sim.group_size(f1) == n_before - N // 100


def test_chained_meiosis_properties(world):
    ds, sim, f1 = world
    lines = sim.make_group_from(sim.group_indexes(f1)[:3000])
    x0 = sim.export_group(lines)["genotypes"]
    het0 = (x0[:, 0::2] != x0[:, 1::2]).mean()
    s5 = sim.self_n_times(5, lines)                               # F1 -> F6
    x5 = sim.export_group(s5)["genotypes"]
    het5 = (x5[:, 0::2] != x5[:, 1::2]).mean()
    assert abs(het5 / het0 - 1 / 32) < 0.01
    dh = sim.make_doubled_haploids(s5)
    z = sim.export_group(dh)["genotypes"]
    assert np.array_equal(z[:, 0::2], z[:, 1::2])
    assert np.all((z[:, 0::2] == x5[:, 0::2]) | (z[:, 0::2] == x5[:, 1::2]))
    cl = sim.make_clones(dh)
    assert np.array_equal(sim.export_group(cl)["genotypes"], z)
