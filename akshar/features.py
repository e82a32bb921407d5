"""reference src/akshar/features.py -> the two wrappers that are functions of the cluster boundaries (SURVEY.md section 8
row f4); the other eighteen lean on modules that are out of scope (sandhi, schwa, anusvara, vedic, transliteration ...)."""
from akshar_b200.segment import (akshara_level_tokenization, preserve_nukta, akshara_level_tokenization_batch,  # noqa: F401
                                 preserve_nukta_batch)
