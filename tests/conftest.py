import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own viterbi.c/trellis.c (oracle/_ref); skipped when not built."""
    from oracle.oracle import Reference
    try:
        return Reference()
    except (FileNotFoundError, OSError) as e:
        pytest.skip(f"oracle/_ref unavailable: {e}")


@pytest.fixture(scope="session")
def golden_profiles():
    from deciphon_b200.synth import load_golden_profiles
    return load_golden_profiles()


@pytest.fixture(scope="session")
def golden_reads():
    return json.load(open(os.path.join(GOLDEN, "reads.json")))


@pytest.fixture(scope="session")
def ref_vectors():
    z = np.load(os.path.join(GOLDEN, "ref_vectors.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def node_pool(golden_profiles):
    from deciphon_b200.synth import NodePool
    return NodePool(golden_profiles)


@pytest.fixture(scope="session")
def device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deciphon_b200.device import Device
    d = Device(0)
    yield d
    d.close()
