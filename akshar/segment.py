"""reference src/akshar/segment.py -> akshar_b200.segment"""
from akshar_b200.segment import (segment_akshars, identify_script, detect_code_switches, segment_by_script,  # noqa: F401
                                 analyze_text_composition, is_matra, akshar_PAT, MATRA_RANGES, segment_akshars_batch,
                                 detect_code_switches_batch, analyze_text_composition_batch, word_tokenize,
                                 word_tokenize_hindi, word_tokenize_sanskrit, word_tokenize_batch,
                                 word_tokenize_hindi_batch, word_tokenize_sanskrit_batch)
