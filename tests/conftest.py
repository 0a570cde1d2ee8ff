import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `pytest -m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    """A GPU test takes seconds (the whole `-m gpu` suite under a minute).  With pytest-timeout present, one that is
    still running after 15 minutes is a hang (the kernels' barrier waits are bounded and trap, so this is a last
    resort): the run ends with the stacks of all threads instead of holding the GPU until an outer limit."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900, method="thread"))


@pytest.fixture(scope="session")
def golden_rows():
    z = np.load(os.path.join(GOLDEN, "chr22_subset50_rows.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_full():
    """The whole chr22_subset50 fixture (64 samples x 1,066,557 SNPs) with the keep mask that the reference's own QC
    ladder (tests/pca.py:86-105) gives for it -- tests/golden/make_golden.py."""
    z = np.load(os.path.join(GOLDEN, "chr22_subset50_full.npz"))
    d = {k: z[k] for k in z.files}
    m = d["payload"].shape[0]
    d["keep_ref"] = np.unpackbits(d["keep_ref_bits"])[:m].astype(bool)
    d["keep_rust"] = np.unpackbits(d["keep_rust_bits"])[:m].astype(bool)
    return d


@pytest.fixture(scope="session")
def gpu_ctx():
    import genomic_pca_b200 as gp
    ctx = gp.Context(0)
    yield ctx
    ctx.close()
