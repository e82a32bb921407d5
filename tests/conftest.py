import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tools')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
MODELS = os.path.join(GOLDEN, 'models')

# the reference's own test files kept as fixtures: run (unmodified) by test_gpu_parity.py, never collected directly
collect_ignore_glob = ['golden/ref_tests/*']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    with gzip.open(os.path.join(GOLDEN, 'reference_vectors.json.gz'), 'rb') as f:
        return json.loads(f.read().decode('utf-8'))


@pytest.fixture(scope='session')
def decode_fuzz():
    """id rows decoded / detokenized by the unmodified reference (tools/make_golden_decode.py)"""
    with gzip.open(os.path.join(GOLDEN, 'decode_fuzz.json.gz'), 'rb') as f:
        return json.loads(f.read().decode('utf-8'))


@pytest.fixture(scope='session')
def models_dir():
    return MODELS


@pytest.fixture(scope='session')
def bpe_rows(golden):
    """golden rows whose normalized text the BPE path accepts: every code point BPE-safe (HF's NFKC acts on it like NFC;
    U+09FE, newer than HF's Unicode tables, is the one member of normalize_text's alphabet that is not).  The other
    rows must make the encoder fail loudly (AKSHAR_ST_ALPHABET)."""
    import akshar_oracle as O
    T = O.tables()
    return [r for r in golden['rows'] if all(T.bpe_safe[ord(c)] for c in r['norm'])]


@pytest.fixture(scope='session')
def golden_raw():
    """vectors recorded from the unmodified reference with clean_hinglish=False (tools/make_golden_raw.py)"""
    with gzip.open(os.path.join(GOLDEN, 'reference_vectors_raw.json.gz'), 'rb') as f:
        return json.loads(f.read().decode('utf-8'))
