"""Host grain generator (host/grains.cpp) vs the reference's GrainStructure::generate: bit-exact
flags against the golden vectors and, where oracle/_ref is present, against the reference."""
import json
import os

import numpy as np
import pytest

import helpers as H
from oracle import refapi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", autouse=True)
def _build_host():
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "pd_mg_pin_corrosion_b200", "csrc"), "-j8"],
                          stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(root, "host")], stdout=subprocess.DEVNULL)


@pytest.mark.parametrize("case", ["2d_default", "2d_poiseuille", "2d_offgrid", "3d_small", "3d_offgrid"])
def test_grains_match_golden(case):
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    z = np.load(os.path.join(GOLD, f"steps_{case}.npz"))
    N = int(z["dims"][3])
    dim, cfg, _ = H.load_cfg(case)
    g = GrainStructure().generate(z["node_type"], cfg, dim)
    assert np.array_equal(g.is_grain_boundary, np.unpackbits(z["is_gb"])[:N])
    assert np.array_equal(g.is_precipitate, np.unpackbits(z["is_precip"])[:N])


@pytest.mark.skipif(not refapi.have_ref(3), reason="oracle/_ref not built")
@pytest.mark.parametrize("case,extra", [("3d_default", None), ("2d_default", {"gb_width_cells": 2, "precip_cluster_cells": 2}),
                                        ("3d_small", {"gb_width_cells": 1, "precip_cluster_cells": 1, "precip_fraction": 0.2})])
def test_grains_match_reference(case, extra):
    from pd_mg_pin_corrosion_b200.grains import GrainStructure
    r = H.make_ref(case, extra)
    dim, cfg, _ = H.load_cfg(case, extra)
    g = GrainStructure().generate(r.get("node_type"), cfg, dim)
    assert g.n_grains == r.lib.ref_n_grains(r.h)
    assert np.array_equal(g.grain_id, r.get("grain_id"))
    assert np.array_equal(g.is_grain_boundary, r.get("is_gb"))
    assert np.array_equal(g.is_precipitate, r.get("is_precip"))
    if case == "3d_default":   # SURVEY.md section 4 known answers
        assert (g.n_grains, int(g.is_grain_boundary.sum()), int(g.is_precipitate.sum())) == (58, 11120, 224)
