"""reference src/akshar/normalize.py -> akshar_b200.normalize"""
from akshar_b200.normalize import (normalize_unicode, semantic_normalize, remove_elongations, roman_phonetic_signature,  # noqa: F401
                                   filter_garbage, normalize_hinglish, normalize_text, normalize_batch,
                                   roman_phonetic_signature_batch)
