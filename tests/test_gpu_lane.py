"""GPU: the lane-per-read dense kernels (dense_lane.cu, k <= 4) under every lane-group width the build
offers, and the round-1 kernels they replace, against the oracle.  The switches are read once per
process, so each variant runs tests/manual/lane_variants.py in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("env", [
    {},                                                    # defaults
    {"CFRK_LANE_SPLIT_K3": "1", "CFRK_LANE_SPLIT_K4": "1"},
    {"CFRK_LANE_SPLIT_K3": "4", "CFRK_LANE_SPLIT_K4": "2"},
    {"CFRK_DENSE_LANE": "0"},                              # the CTA / warp tile kernels of round 1
], ids=["default", "split1", "split4_2", "round1_kernels"])
def test_lane_variants_match_oracle(env):
    e = dict(os.environ)
    e.update(env)
    ks = "3,4" if "CFRK_LANE_SPLIT_K3" in env else "1,2,3,4"
    r = subprocess.run([sys.executable, os.path.join(HERE, "manual", "lane_variants.py"), "--ks", ks], env=e,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
