"""The kernels that only run when an environment switch (or a device without cooperative launch)
selects them: the single-CTA claim resolution of the projection searches, the round-robin dataflow
Cholesky and the multi-kernel Cholesky.  Each configuration runs tests/fallback_paths.py in its own
process (one of the switches is read once per process)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("env", [
    {"LORB_RESOLVE_COOP": "0", "LORB_CHOL_CHAIN": "0"},
    {"LORB_RESOLVE_COOP": "0", "LORB_CHOL_DATAFLOW": "0"},
    {"LORB_HAMMING_CSA": "0"},
], ids=["single-cta-resolve+dataflow-cholesky", "single-cta-resolve+multi-kernel-cholesky", "plain-popc-body"])
def test_fallback_kernels_match_the_oracle(env):
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, os.path.join(HERE, "fallback_paths.py")], env=e, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "FALLBACK_PATHS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
