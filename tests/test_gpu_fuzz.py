"""A fixed slice of the randomised parity campaigns (tools/fuzz_parity.py, tools/fuzz_match.py): random shapes, extractor
parameters, batch sizes, pathological image contents; random keypoint clouds, clustered descriptors, radii, level ranges,
thresholds and ratios for the matchers and ComputeStereoMatches — CUDA (C ABI) against the oracle, bit-exact.
The long campaigns (thousands of cases, `profiles/r2/fuzz_summary.txt`) are run with the tools themselves."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_fuzz_extractor_slice(oracle):
    import fuzz_parity
    lines = []
    fails = sum(fuzz_parity.one_case(np.random.default_rng([7, case]), case, lines.append) for case in range(30))
    assert fails == 0, [l for l in lines if l.startswith("MISMATCH")]
    assert sum(l.startswith("ok") for l in lines) >= 20      # the rest are geometries the library refuses like the reference would fail


@pytest.mark.parametrize("kind", ["bruteforce", "projection", "initialization", "stereo"])
def test_fuzz_matchers_slice(oracle, kind):
    import fuzz_match
    fn = getattr(fuzz_match, "case_" + kind)
    for case in range(12 if kind == "stereo" else 60):
        desc, ok, why = fn(np.random.default_rng([11, case]))
        assert ok, "%s: %s" % (desc, why)
