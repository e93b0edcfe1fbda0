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
