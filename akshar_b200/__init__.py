"""akshar_b200: B200-native batch path for Akshar's hot path (normalize -> akshars -> script runs -> subword ids).

Importable like the reference package for that path: `aksharTokenizer`, `normalize_text`, `segment_akshars`,
`detect_code_switches`, ... keep the reference's signatures and results, computed by hand-written sm_100a CUDA
kernels in `lib/libakshar_b200.so` (C ABI: include/akshar_b200.h).  Nothing here falls back to the CPU.
"""
__version__ = "0.1.0"

from .tokenizer import aksharTokenizer, AksharTokenizer
from .segment import (segment_akshars, detect_code_switches, segment_by_script, analyze_text_composition, identify_script,
                      is_matra, akshar_PAT, MATRA_RANGES, segment_akshars_batch, detect_code_switches_batch,
                      analyze_text_composition_batch, word_tokenize, word_tokenize_hindi, word_tokenize_sanskrit,
                      word_tokenize_batch, word_tokenize_hindi_batch, word_tokenize_sanskrit_batch, akshara_level_tokenization,
                      preserve_nukta, akshara_level_tokenization_batch, preserve_nukta_batch)
from .normalize import (normalize_text, normalize_hinglish, normalize_unicode, semantic_normalize, remove_elongations,
                        roman_phonetic_signature, filter_garbage, normalize_batch, roman_phonetic_signature_batch)
from .batch import Engine, engine, Ragged, TextBatch

__all__ = [
    'aksharTokenizer', 'AksharTokenizer', 'segment_akshars', 'detect_code_switches', 'segment_by_script',
    'analyze_text_composition', 'identify_script', 'is_matra', 'akshar_PAT', 'MATRA_RANGES', 'normalize_text',
    'normalize_hinglish', 'normalize_unicode', 'semantic_normalize', 'remove_elongations', 'roman_phonetic_signature',
    'filter_garbage', 'normalize_batch', 'segment_akshars_batch', 'detect_code_switches_batch',
    'analyze_text_composition_batch', 'roman_phonetic_signature_batch', 'word_tokenize', 'word_tokenize_hindi',
    'word_tokenize_sanskrit', 'word_tokenize_batch', 'word_tokenize_hindi_batch', 'word_tokenize_sanskrit_batch', 'akshara_level_tokenization', 'preserve_nukta',
    'akshara_level_tokenization_batch', 'preserve_nukta_batch', 'Engine', 'engine', 'Ragged', 'TextBatch',
]
