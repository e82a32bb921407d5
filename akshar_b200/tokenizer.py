"""Drop-in for the reference's `akshar.tokenizer.aksharTokenizer` (src/akshar/tokenizer.py) on the CUDA hot path.

Same constructor, attributes, methods, return types and exceptions as the reference; `tokenize_batch` /
`encode_batch` are the new batch entry points and return ragged token-id tensors.  The model files are the ones the
reference's scripts/train_bpe.py (HuggingFace tokenizers JSON) and scripts/train_spm.py (SentencePiece ModelProto)
write; they are parsed by libakshar_b200.so itself, `tokenizers` / `sentencepiece` are not needed at run time.
"""
import os
from typing import List, Optional, Union

from . import _lib as C
from .batch import engine
from .normalize import normalize_text, normalize_batch
from .segment import (segment_akshars, detect_code_switches, analyze_text_composition, segment_akshars_batch,
                      analyze_text_composition_batch, detect_code_switches_batch)

_SPM_NORMAL, _SPM_UNKNOWN, _SPM_CONTROL, _SPM_USER, _SPM_UNUSED, _SPM_BYTE = 1, 2, 3, 4, 5, 6


class _CudaModel:
    """what `self.model` holds: the vocabulary of the model loaded into the device context"""

    def __init__(self, eng, kind):
        self.kind = kind                       # 0 BPE, 1 Unigram
        self.size, self.vocab = eng.vocab(kind)

    def token(self, i):
        return self.vocab[i][0]


class aksharTokenizer:
    """reference tokenizer.py:18-288"""

    def __init__(self, model_path: Optional[str] = None, model_type: str = "sentencepiece", normalize_roman: bool = True,
                 clean_hinglish: bool = True, device: int = 0):
        self.model_path = model_path
        self.normalize_roman = normalize_roman
        self.clean_hinglish = clean_hinglish
        self.model = None
        self._configured_model_type = model_type
        self._device = device
        if model_path and os.path.exists(model_path):
            self._load_model()
        else:
            self.model_type = "akshar"        # reference tokenizer.py:69-71: silent fallback

    def _load_model(self):
        model_type = self._configured_model_type
        # one context per tokenizer so that two tokenizers with different models can coexist
        from .batch import Engine
        if model_type == "sentencepiece":
            self._eng = Engine(self._device)
            try:
                self._eng.load_spm(self.model_path)
            except C.AksharCudaError as e:
                raise RuntimeError(str(e))        # SentencePiece raises RuntimeError on an unparsable ModelProto
            self.model = _CudaModel(self._eng, 1)
            self.model_type = "sentencepiece"
        elif model_type == "bpe":
            self._eng = Engine(self._device)
            try:
                self._eng.load_bpe(self.model_path)
            except C.AksharCudaError as e:
                raise Exception(str(e))           # tokenizers raises a bare Exception on a bad JSON
            self.model = _CudaModel(self._eng, 0)
            self.model_type = "bpe"
        else:
            raise ValueError(f"unknown model_type: {model_type}")

    # ------------------------------------------------------------------ reference API
    def preprocess(self, text: str) -> str:
        return normalize_batch([text], self.normalize_roman, self.clean_hinglish, device=self._device)[0]

    def tokenize(self, text: str, return_metadata: bool = False) -> Union[List[str], dict]:
        norm = self.preprocess(text)
        if return_metadata:
            meta = analyze_text_composition_batch([norm], device=self._device)[0]
        if self.model is None:
            tokens = segment_akshars_batch([norm], device=self._device)[0]
        else:
            tokens = self._pieces(self._encode_normalized([norm])[0])
        if return_metadata:
            meta['tokens'] = tokens
            meta['token_count'] = len(tokens)
            meta['original_text'] = text
            meta['normalized_text'] = norm
            return meta
        return tokens

    def encode(self, text: str) -> List[int]:
        norm = self.preprocess(text)
        if self.model is None:
            raise ValueError("need model for IDs")
        return self._encode_normalized([norm])[0]

    def decode(self, ids: List[int]) -> str:
        if self.model is None:
            raise ValueError("need model to decode")
        if self.model_type == "sentencepiece":
            return self._decode_spm(ids)
        return ' '.join(self.model.token(i) for i in ids if not self.model.vocab[i][1])     # HF decode, decoder = null

    def detokenize(self, tokens: List[str]) -> str:
        if self.model_type == "sentencepiece":
            return ''.join(tokens).replace('▁', ' ').strip()
        elif self.model_type == "bpe":
            return ' '.join(tokens).replace(' ##', '').replace('Ġ', ' ').strip()
        return ''.join(tokens)

    def explain(self, text: str) -> dict:
        norm = self.preprocess(text)
        return {
            'original': text,
            'normalized': norm,
            'akshars': segment_akshars_batch([norm], device=self._device)[0],
            'code_switches': detect_code_switches_batch([norm], device=self._device)[0],
            'tokens': self.tokenize(text),
            'stats': analyze_text_composition_batch([norm], device=self._device)[0],
        }

    def vocab_size(self) -> int:
        return 0 if self.model is None else self.model.size

    # ------------------------------------------------------------------ batch API (new)
    def encode_batch(self, texts, as_device=False):
        """encode() over a batch: list[list[int]], or with as_device=True (Ragged int32 ids + int64 row_splits on
        the device, normalized TextBatch).  normalize and encode are enqueued back to back without a host sync."""
        if self.model is None:
            raise ValueError("need model for IDs")
        ids, norm = self._eng.tokenizer_encode_batch(texts, self.model.kind, self.normalize_roman, self.clean_hinglish)
        if as_device:
            return ids, norm
        return [r.tolist() for r in ids.rows()]

    def encode_batch_host(self, h_data, h_offsets, out_ids=None, out_splits=None, compact=False):
        """encode() over a batch given as pinned host tensors (uint8 text, int64 row offsets) -> pinned host tensors
        (int32 ids, int64 row_splits), or with compact=True a `CompactIds` (uint16 ids, int32 chunk-relative splits: a third
        of the bytes over the host link); copies and kernels are pipelined over three streams.  The results live in buffers
        the engine reuses: see Engine.encode_host_pipelined."""
        if self.model is None:
            raise ValueError("need model for IDs")
        return self._eng.encode_host_pipelined(h_data, h_offsets, self.model.kind, self.normalize_roman, self.clean_hinglish,
                                               out_ids=out_ids, out_splits=out_splits, compact=compact)

    def tokenize_batch(self, texts):
        """tokenize() over a batch -> list[list[str]]"""
        if self.model is None:
            norm = normalize_batch(texts, self.normalize_roman, self.clean_hinglish, device=self._device)
            return segment_akshars_batch(norm, device=self._device)
        return [self._pieces(ids) for ids in self.encode_batch(texts)]

    def explain_batch(self, texts):
        norm = normalize_batch(texts, self.normalize_roman, self.clean_hinglish, device=self._device)
        ak = segment_akshars_batch(norm, device=self._device)
        cs = detect_code_switches_batch(norm, device=self._device)
        st = analyze_text_composition_batch(norm, device=self._device)
        tk = self.tokenize_batch(texts)
        return [{'original': t, 'normalized': n, 'akshars': a, 'code_switches': c, 'tokens': k, 'stats': s}
                for t, n, a, c, k, s in zip(texts, norm, ak, cs, tk, st)]

    # ------------------------------------------------------------------ helpers
    def _encode_normalized(self, norms):
        if self.model.kind == 0:
            r = self._eng.encode_bpe_batch(norms)
        else:
            r = self._eng.encode_unigram_batch(norms)
        return [x.tolist() for x in r.rows()]

    def _pieces(self, ids):
        return [self.model.token(i) for i in ids]

    def _decode_spm(self, ids):
        """SentencePiece DecodeIds for a Unigram model with byte fallback: pieces are concatenated, runs of <0xNN>
        pieces are reassembled into UTF-8 (invalid bytes -> U+FFFD each), <unk> decodes to ' ⁇ ', control pieces
        are dropped, U+2581 becomes a space and the dummy-prefix space is removed."""
        out = []
        pend = bytearray()

        def flush():
            if pend:
                out.append(_decode_bytes_spm(bytes(pend)))
                pend.clear()

        first = True
        for i in ids:
            if i < 0 or i >= self.model.size:
                raise IndexError('piece id is out of range.')
            piece, ty = self.model.vocab[i]
            if ty == _SPM_BYTE:
                pend.append(int(piece[3:5], 16))
                continue
            flush()
            if ty == _SPM_CONTROL:
                continue
            if ty == _SPM_UNKNOWN:
                out.append(' ⁇ ')
                first = False
                continue
            if first and piece.startswith('▁'):
                piece = piece[1:]
            first = False
            out.append(piece.replace('▁', ' '))
        flush()
        return ''.join(out)


def _decode_bytes_spm(b):
    # maximal valid UTF-8 prefixes, each invalid byte becomes U+FFFD
    out = []
    i = 0
    while i < len(b):
        for n in (1, 2, 3, 4):
            try:
                ch = b[i:i + n].decode('utf-8')
                if len(ch) == 1:
                    out.append(ch)
                    i += n
                    break
            except UnicodeDecodeError:
                continue
        else:
            out.append('�')
            i += 1
    return ''.join(out)


AksharTokenizer = aksharTokenizer      # the name the reference's tests/test_tokenizer.py:11 imports
