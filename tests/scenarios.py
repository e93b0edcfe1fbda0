"""Breeding-scheme scenarios that run unchanged against any implementation exposing the
shared method names (oracle.ref.RefSim, oracle.port.PortSim, the CUDA engine's SimData).

Each scenario only uses the hot-path API and returns
the list of group numbers it created, so callers can compare exported state group by group.
"""
import numpy as np


def crossing_scheme(sim, founders, small=False):
    """Every crossing entry point once, with caps, families, invalid targeted pairs
    (core.c:8751-8776 skips and counts them) and chained selfing (core.c:8900-8926)."""
    m0 = sim.map_arg(0)
    k = 1 if small else 3
    groups = []
    g2 = sim.make_random_crosses(founders, 10 * k, 0, m0, family_size=2)
    groups.append(g2)
    g3 = sim.make_random_crosses(founders, 4, 1, m0)
    groups.append(g3)
    g4 = sim.self_n_times(3, g2, m0)
    groups.append(g4)
    g5 = sim.self_n_times(2, g3, m0, family_size=2)
    groups.append(g5)
    g6 = sim.make_doubled_haploids(g4, m0)
    groups.append(g6)
    n_total = sim.n_genotypes
    g7 = sim.make_targeted_crosses([0, 3, 5, n_total + 50, 2, 0xFFFFFFFF], [1, 4, 6, 2, 7, 1],
                                   m0, m0, family_size=2)
    groups.append(g7)
    g8 = sim.make_random_crosses_between(g2, g4, 6 * k, 1, 2, m0, m0)
    groups.append(g8)
    g9 = sim.make_all_unidirectional_crosses(g3, m0)
    groups.append(g9)
    g10 = sim.make_clones(g6)
    groups.append(g10)
    g11 = sim.make_random_crosses(g8, 5, 0, m0, family_size=1, track_pedigree=False)
    groups.append(g11)
    return groups


def naming_scheme(sim, founders):
    """Names: prefix + id counter per 1000-genotype buffer (core.c:552-591, 8168-8180), clones that
    inherit their parent's name (core.c:9097-9108), unnamed offspring, ids switched off."""
    m0 = sim.map_arg(0)
    groups = []
    groups.append(sim.make_random_crosses(founders, 7, 0, m0, family_size=3, name_offspring=True, name_prefix="F1_"))
    groups.append(sim.make_clones(groups[0], inherit_names=True))
    groups.append(sim.make_clones(founders, inherit_names=False, name_offspring=True, name_prefix="cl"))
    groups.append(sim.self_n_times(1, groups[0], m0, allocate_ids=False, name_offspring=True, name_prefix="S"))
    groups.append(sim.make_doubled_haploids(groups[3], m0))
    groups.append(sim.make_random_crosses(groups[1], 1100, 0, m0, name_offspring=True, name_prefix="big"))
    return groups


def export_state(sim):
    x = sim.export_group(0)
    return dict(genotypes=x["genotypes"], ids=x["ids"], ped1=x["ped1"], ped2=x["ped2"], groups=x["groups"])


def assert_same_state(a, b):
    for key in ("groups", "ids", "ped1", "ped2"):
        np.testing.assert_array_equal(a[key], b[key], err_msg=key)
    np.testing.assert_array_equal(a["genotypes"], b["genotypes"], err_msg="genotype bytes")


def saver_scheme(sim, founders, effs, out, big=False):
    """The save/see boundary: every saver in both orientations, local GEBV
    tables, and the three save-as-you-go files of the crossing scaffold with and without names, ids
    and retained offspring (reference tests: test-printers.R:1-65, 67-111, 200-323;
    test-progression.R:54-109).  `out`: directory (str) the files go to; returns their names in order."""
    import os
    m0 = sim.map_arg(0)
    e0, e1 = effs
    files = []

    def f(name):
        files.append(name)
        return os.path.join(out, name)

    sim.save_genotypes(f("geno-group.txt"), founders, False)
    sim.save_genotypes(f("geno-all.txt"), 0, False)
    sim.save_genotypes(f("geno-group-t.txt"), founders, True)
    sim.save_genotypes(f("geno-all-t.txt"), 0, True)
    sim.save_allele_counts(f("count-T.txt"), founders, "T", False)
    sim.save_allele_counts(f("count-A-t.txt"), 0, "A", True)
    sim.save_bvs(f("bv.txt"), founders, e0)
    b = sim.blocks_evenlength(3, m0)
    sim.save_local_bvs(f("local-headers.txt"), founders, b, e0, True)
    sim.save_local_bvs(f("local-bare.txt"), 0, b, e1, False)
    b.free()
    # save-as-you-go, as in test-progression.R:54-109
    sim.make_clones(founders, True, file_prefix=os.path.join(out, "sayg1"), save_genotypes=True, retain=False)
    files.append("sayg1-genotype.txt")
    sim.make_clones(founders, True, file_prefix=os.path.join(out, "sayg2"), save_bvs=e0, retain=False)
    files.append("sayg2-bv.txt")
    sim.make_doubled_haploids(founders, m0, file_prefix=os.path.join(out, "sayg3"), save_pedigree=True, retain=False)
    files.append("sayg3-pedigree.txt")
    sim.make_clones(founders, False, file_prefix=os.path.join(out, "sayg4"), save_genotypes=True, save_bvs=e0, save_pedigree=True,
                    allocate_ids=False, retain=False)     # (retained id-less genotypes send the reference's id bisection, core.c:362-418, into a loop)
    files += ["sayg4-genotype.txt", "sayg4-bv.txt", "sayg4-pedigree.txt"]
    # crosses whose pedigrees recurse over named and unnamed ancestors
    n = 600 if big else 5
    f1 = sim.make_random_crosses(founders, n, 0, m0, family_size=2, name_offspring=True, name_prefix="x",
                                 file_prefix=os.path.join(out, "c1"), save_genotypes=True, save_bvs=e1, save_pedigree=True)
    files += ["c1-genotype.txt", "c1-bv.txt", "c1-pedigree.txt"]
    f2 = sim.self_n_times(2, f1, m0, file_prefix=os.path.join(out, "c2"), save_pedigree=True, save_genotypes=True)
    files += ["c2-pedigree.txt", "c2-genotype.txt"]
    f3 = sim.make_random_crosses_between(f1, f2, n, 0, 0, m0, m0, track_pedigree=True, allocate_ids=(not big),
                                         file_prefix=os.path.join(out, "c3"), save_pedigree=True, save_bvs=e0, retain=False)
    files += ["c3-pedigree.txt", "c3-bv.txt"]
    sim.save_pedigrees(f("ped-full.txt"), f2, True)
    sim.save_pedigrees(f("ped-parents.txt"), 0, False)
    sim.save_genotypes(f("geno-f2-t.txt"), f2, True)
    sim.save_bvs(f("bv-all.txt"), 0, e1)
    return files
