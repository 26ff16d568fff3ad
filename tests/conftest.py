import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SIMPLE = os.path.join(GOLDEN, "simple")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def simple_key():
    from plonkit_b200 import reader
    return reader.load_key_monomial_form(os.path.join(SIMPLE, "setup_2^10.key"))


@pytest.fixture(scope="session")
def simple_circuit():
    from plonkit_b200 import circuit, reader
    r1cs = reader.load_r1cs(os.path.join(SIMPLE, "circuit.r1cs.json"))
    wit = reader.load_witness_from_file(os.path.join(SIMPLE, "witness.json"))
    return circuit.CircomCircuit(r1cs, wit)


@pytest.fixture(scope="session")
def ctx():
    from plonkit_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def vk_commitments(vk):
    return np.concatenate([vk.selector_commitments, vk.next_step_selector_commitments, vk.permutation_commitments])
