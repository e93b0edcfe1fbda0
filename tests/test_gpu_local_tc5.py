"""Local GEBVs for many effect sets in one pass: the tcgen05 kernel (csrc/local_tc5.cu, taken from 5
effect sets on) against the one-set-at-a-time mma.sync kernel and the oracle, on shapes that exercise
several marker segments, blocks that straddle words and segments, a last tile that is not full,
odd set counts (N padded to a multiple of 16) and centred effects (core.c:9979-10090)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import datasets, port   # noqa: E402

pytestmark = pytest.mark.gpu


def multi(sim, rows, off, mk, sets, device_out=False):
    import torch
    n, nb = len(rows), len(off) - 1
    effs = np.asarray(sets, dtype=np.uint32)
    if device_out:
        outs = [torch.full((2 * n, nb), 7.0, dtype=torch.float64, device="cuda") for _ in sets]
        torch.cuda.synchronize()
        ptrs = (C.c_void_p * len(sets))(*[o.data_ptr() for o in outs])
        st = sim.E.gsb200_local_gebv_multi_dev(sim.ctx, n, rows.ctypes.data, nb, off.ctypes.data, mk.ctypes.data, len(sets), effs.ctypes.data, ptrs)
        assert st == 0, sim.E.gsb200_last_error()
        return [o.cpu().numpy() for o in outs]
    outs = [np.full((2 * n, nb), 7.0) for _ in sets]
    ptrs = (C.c_void_p * len(sets))(*[o.ctypes.data for o in outs])
    st = sim.E.gsb200_local_gebv_multi(sim.ctx, n, rows.ctypes.data, nb, off.ctypes.data, mk.ctypes.data, len(sets), effs.ctypes.data, ptrs)
    assert st == 0, sim.E.gsb200_last_error()
    return outs


@pytest.mark.parametrize("shape", [(150, 9000, 5, 7, 10), (64, 2100, 3, 1, 5), (333, 20000, 21, 10, 7), (40, 700, 2, 30, 12)])
def test_many_sets_in_one_pass_match_one_set_at_a_time(shape):
    from genomicsimulation_b200 import SimData
    n, M, n_chr, per_chr, S = shape
    ds = datasets.synthetic(n, M, n_chr, 120.0, seed=M + S, spacing="random", n_effect_sets=S)
    rng = np.random.default_rng(S)
    ds["effects"][S - 1]["centre"] = rng.standard_normal(M)          # one centred set (core.c:10075-10076)
    sim = SimData.from_dataset(ds)
    blocks = sim.blocks_evenlength(per_chr, 1)
    off, mk = blocks.export()
    rows = np.arange(n, dtype=np.uint32)
    one_by_one = [multi(sim, rows, off, mk, [q])[0] for q in range(S)]          # < 5 sets: the mma.sync kernel
    together = multi(sim, rows, off, mk, list(range(S)))
    on_device = multi(sim, rows, off, mk, list(range(S)), device_out=True)
    for q in range(S):
        scale = np.abs(one_by_one[q]).mean() + 1e-300
        assert np.max(np.abs(together[q] - one_by_one[q])) <= 1e-12 * scale, q
        assert np.array_equal(on_device[q], together[q]), q
    # and against the oracle's sequential fp64 sums (1e-9 relative, BASELINE.json north_star)
    p = port.PortSim.from_dataset(ds)
    for q in (0, S - 1):
        want = p.calculate_local_bvs(1, off, mk, q)
        assert np.max(np.abs(together[q] - want)) <= 1e-9 * np.maximum(np.abs(want), np.abs(want).mean()).max()
    p.close()
    sim.close()
