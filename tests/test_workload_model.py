"""CPU test: the anchor-level workload model used by bench.py has the shape of real minimap2 seeding of the BASELINE.json
configurations.  tests/golden/workload_calibration.json was measured once with the reference CLI on sequence-level simulated
reads vs a 100 Mbp reference (tests/golden/make_calibration.py); the model must stay within tolerance of it on the statistics
the chaining cost depends on (anchors/read, reference cells per anchor, window size, chained fraction)."""
import json
import os

import pytest

from conftest import GOLDEN

TOL = {"anchors_per_read": 0.15, "cells_per_anchor": 0.10, "window_cells_per_anchor": 0.25, "chained_fraction": 0.08}


@pytest.mark.parametrize("name,n_reads", [("map-ont", 400), ("asm20", 120), ("ultralong", 40)])
def test_model_matches_real_seeding(oracle, pkg, name, n_reads):
    path = os.path.join(GOLDEN, "workload_calibration.json")
    if not os.path.exists(path):
        pytest.skip("calibration file not generated")
    real = json.load(open(path))[name]
    wl = pkg("workload")
    off, a = wl.preset_batch(name, n_reads, seed=3)
    st = oracle.replay(oracle.Params(), off, a, n_threads=4)["stats"]
    got = dict(anchors_per_read=len(a) / n_reads, cells_per_anchor=st.cells / len(a),
               window_cells_per_anchor=st.window_cells / len(a), chained_fraction=st.n_chained / len(a))
    for k, tol in TOL.items():
        if name == "ultralong" and k == "anchors_per_read":
            tol = 0.35          # 24 real reads with Gamma(4)-distributed lengths: the sample mean itself wanders by ~10-20 %
        assert abs(got[k] - real[k]) <= tol * real[k], (name, k, got[k], real[k])
    assert abs(st.n_chains / n_reads - real["chains_per_read"]) < 0.1
