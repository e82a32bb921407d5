"""`akshar` -- the reference's import name (src/akshar/__init__.py:12-58), served by the B200 path.

Everything the hot path of the reference exports is re-exported from `akshar_b200` (hand-written sm_100a kernels behind
the C ABI of include/akshar_b200.h): the tokenizer class under both spellings the reference uses (`aksharTokenizer`,
tokenizer.py:18, and `AksharTokenizer`, tests/test_tokenizer.py:11), the normalize / segment entry points, the word
tokenizers.  The reference's other modules (phonetics, sandhi, morphology, features, CLI, app ...) are out of scope
 and are not provided.
"""
__version__ = "0.1.0"

from akshar_b200 import *  # noqa: F401,F403
from akshar_b200 import __all__ as _all
from akshar_b200 import aksharTokenizer, AksharTokenizer  # noqa: F401

__all__ = list(_all)
