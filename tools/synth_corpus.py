"""Seeded synthetic corpora for parity tests and bench.py.

Everything is generated from a seed with numpy, chunk by chunk, as the concatenated-UTF-8 + row_offsets
layout the CUDA path consumes (one sentence per row, no empty rows).  Kinds:

  hinglish   Zipf draws from a Devanagari + Roman lexicon, 4-18 words / sentence, script flips p~0.3, 8 % punctuation
  hindi      Devanagari-only sentences with danda / double danda and rare Vedic accents
  social     hinglish + elongations, random upper-casing, emoji / ZWJ sequences / flags, !!! / ..., precomposed
             nukta letters and decomposed Latin + combining marks (exercises the NFC slow lane)
  adversarial(n, seed)  random strings over the survey's fuzz alphabet (not timed; parity only)
"""
import numpy as np

_CONS = [chr(c) for c in range(0x0915, 0x093A)]
_NUKTA_OK = ['क', 'ख', 'ग', 'ज', 'ड', 'ढ', 'फ', 'य', 'न', 'र', 'ळ']
_MATRA = [chr(c) for c in (0x093E, 0x093F, 0x0940, 0x0941, 0x0942, 0x0947, 0x0948, 0x094B, 0x094C, 0x0943)]
_VOWEL = [chr(c) for c in range(0x0905, 0x0915)]
_FINAL = ['ं', 'ः', 'ँ']
_RC = list('bcdfghjklmnprstvwyz') + ['kh', 'gh', 'ch', 'th', 'ph', 'bh', 'dh', 'sh']
_RV = ['a', 'e', 'i', 'o', 'u', 'aa', 'ee', 'oo', 'ai', 'au']
_EMOJI = ['😀', '😂', '🙏', '❤️', '🔥', '👍🏽', '👨‍👩‍👧', '🇮🇳', '🇺🇸', '🏳️‍🌈', '✨', '🤣', '😭', '💯']


def _dev_word(rng, vedic=False, nukta_p=0.03):
    n = int(rng.integers(1, 5))
    out = []
    for s in range(n):
        if s == 0 and rng.random() < 0.15:
            out.append(_VOWEL[int(rng.integers(len(_VOWEL)))])
            continue
        c = _CONS[int(rng.integers(len(_CONS)))]
        out.append(c)
        if rng.random() < nukta_p:
            out.append('़')
        if rng.random() < 0.22:
            out.append('्')
            out.append(_CONS[int(rng.integers(len(_CONS)))])
            if rng.random() < 0.08:
                out.append('्')
                out.append(_CONS[int(rng.integers(len(_CONS)))])
        if rng.random() < 0.6:
            out.append(_MATRA[int(rng.integers(len(_MATRA)))])
        if rng.random() < 0.1:
            out.append(_FINAL[int(rng.integers(len(_FINAL)))])
        if vedic and rng.random() < 0.02:
            out.append(chr(0x0951 + int(rng.integers(2))))
    return ''.join(out)


def _rom_word(rng):
    n = int(rng.integers(1, 4))
    out = []
    for s in range(n):
        if not (s == 0 and rng.random() < 0.2):
            out.append(_RC[int(rng.integers(len(_RC)))])
        out.append(_RV[int(rng.integers(len(_RV)))])
    if rng.random() < 0.4:
        out.append(_RC[int(rng.integers(19))])
    return ''.join(out)


def _socialize(rng, w, roman):
    r = rng.random()
    if roman:
        if r < 0.10:       # elongation of a vowel or of the last letter
            pos = [i for i, ch in enumerate(w) if ch in 'aeiou'] or [len(w) - 1]
            p = pos[int(rng.integers(len(pos)))] if rng.random() < 0.7 else len(w) - 1
            w = w[:p] + w[p] * int(rng.integers(3, 8)) + w[p + 1:]
        elif r < 0.16:
            w = w.upper()
        elif r < 0.24:
            w = w.capitalize()
        elif r < 0.241:    # decomposed Latin + combining mark / precomposed Latin
            w = w + ('é' if rng.random() < 0.5 else 'É')
    else:
        if r < 0.03:       # matra elongation
            w = w + w[-1] * int(rng.integers(2, 5))
        elif r < 0.031:    # precomposed nukta letters (NFC decomposes them)
            w = w + chr(0x0958 + int(rng.integers(8)))
    return w


class Lexicon:
    def __init__(self, words):
        enc = [w.encode('utf-8') for w in words]
        self.n = len(enc)
        self.len = np.array([len(e) for e in enc], dtype=np.int64)
        self.start = np.zeros(self.n, dtype=np.int64)
        np.cumsum(self.len[:-1], out=self.start[1:])
        self.bytes = np.frombuffer(b''.join(enc), dtype=np.uint8)


# The LEXICON is a property of the language, not of a corpus draw: it is always built from these seeds -- the ones the
# 24k models under tests/golden/models were trained with (tools/make_golden.py: hinglish 101, hindi 102) -- so that every
# corpus, whatever its own seed, is in-vocabulary for them the way real text is for a model trained on the same language.
# (`social` shares the hinglish base words: the same generator calls in the same order, then the noisy variants.)
# A corpus seed only decides which words are drawn and how the sentences are put together.
LEXICON_SEED = {'hinglish': 101, 'social': 101, 'hindi': 102}


def _build(kind, seed=None):
    rng = np.random.default_rng(LEXICON_SEED[kind] if seed is None else seed)
    nd, nr = (60000, 40000)
    if kind == 'hindi':
        dev = [_dev_word(rng, vedic=True, nukta_p=0.03) for _ in range(nd)]
        rom = []
    else:
        dev = [_dev_word(rng, nukta_p=0.01) for _ in range(nd)]
        rom = [_rom_word(rng) for _ in range(nr)]
    if kind == 'social':
        # noise is realised at lexicon level: every base word gets 2 noisy variants at higher ranks
        dev = dev + [_socialize(rng, w, False) for w in dev] + [_socialize(rng, w, False) for w in dev]
        rom = rom + [_socialize(rng, w, True) for w in rom] + [_socialize(rng, w, True) for w in rom]
        extra = _EMOJI + ['!!!', '...', '!!!!!', '???', '....', 'hahahaha', 'lolll', '#tag', '@user', 'http://x.co/a_b']
    else:
        extra = []
    return Lexicon(dev), (Lexicon(rom) if rom else None), (Lexicon(extra) if extra else None)


_SUFFIX = [b' ', b', ', b'. ', b'! ', b'? ', b'', b' \xe0\xa5\xa4 ', b' \xe0\xa5\xa5', b' \xe0\xa5\xa4']


class Corpus:
    """chunked generator: `chunk(i, nbytes)` -> (uint8 array, int64 row_offsets) deterministic in (kind, seed, i)"""

    def __init__(self, kind='hinglish', seed=1234):
        assert kind in ('hinglish', 'hindi', 'social')
        self.kind = kind
        self.seed = seed
        self.dev, self.rom, self.extra = _build(kind)
        self.suf = Lexicon.__new__(Lexicon)
        self.suf.n = len(_SUFFIX)
        self.suf.len = np.array([len(s) for s in _SUFFIX], dtype=np.int64)
        self.suf.start = np.zeros(self.suf.n, dtype=np.int64)
        np.cumsum(self.suf.len[:-1], out=self.suf.start[1:])
        self.suf.bytes = np.frombuffer(b''.join(_SUFFIX), dtype=np.uint8)

    def _zipf(self, rng, n, size):
        # Zipf(~1.0) over ranks 0..n-1 by inverse CDF of 1/(r+1)
        u = rng.random(size)
        r = np.exp(u * np.log(n + 1.0)) - 1.0
        return np.minimum(r.astype(np.int64), n - 1)

    def chunk(self, index, nbytes):
        rng = np.random.default_rng([self.seed, index, 77])
        avg_word = 9.0 if self.kind != 'hindi' else 12.0
        nwords = int(nbytes / avg_word * 1.15) + 64
        # sentence structure
        slen = rng.integers(4, 19, size=nwords // 4 + 2)
        ends = np.cumsum(slen) - 1
        ends = ends[ends < nwords]
        is_end = np.zeros(nwords, dtype=bool)
        is_end[ends] = True
        is_start = np.zeros(nwords, dtype=bool)
        is_start[0] = True
        is_start[ends[:-1] + 1] = True
        # script per word
        if self.rom is not None:
            flip = rng.random(nwords) < 0.3
            init = rng.random(nwords) < 0.4          # True = roman
            # state = init at sentence start, then toggled by flips
            sid = np.cumsum(is_start) - 1
            tog = np.cumsum(flip)
            base_tog = tog[np.flatnonzero(is_start)][sid]
            roman = (init[np.flatnonzero(is_start)][sid].astype(np.int64) + (tog - base_tog)) % 2 == 1
        else:
            roman = np.zeros(nwords, dtype=bool)
        dev_id = self._zipf(rng, self.dev.n, nwords)
        wstart = self.dev.start[dev_id]
        wlen = self.dev.len[dev_id]
        src = np.zeros(nwords, dtype=np.int8)      # 0 dev 1 rom 2 extra
        if self.rom is not None:
            rom_id = self._zipf(rng, self.rom.n, nwords)
            wstart = np.where(roman, self.rom.start[rom_id], wstart)
            wlen = np.where(roman, self.rom.len[rom_id], wlen)
            src[roman] = 1
        if self.extra is not None:
            ex = rng.random(nwords) < 0.04
            ex_id = rng.integers(0, self.extra.n, size=nwords)
            wstart = np.where(ex, self.extra.start[ex_id], wstart)
            wlen = np.where(ex, self.extra.len[ex_id], wlen)
            src[ex] = 2
        # suffix after every word
        r = rng.random(nwords)
        suf = np.zeros(nwords, dtype=np.int64)
        suf[r < 0.08] = 1 + (rng.integers(0, 4, size=nwords)[r < 0.08])
        if self.kind == 'hindi':
            suf[is_end] = np.where(rng.random(int(is_end.sum())) < 0.7, 8, 7)
            mid = (~is_end) & (r > 0.97)
            suf[mid] = 6
        else:
            suf[is_end] = 5
            pe = is_end & (r < 0.3)
            suf[pe] = 5
        slen_b = self.suf.len[suf]
        sstart = self.suf.start[suf]
        # interleave word / suffix segments and gather
        seg_len = np.empty(2 * nwords, dtype=np.int64)
        seg_len[0::2] = wlen
        seg_len[1::2] = slen_b
        seg_src = np.empty(2 * nwords, dtype=np.int64)
        off_dev = 0
        off_rom = self.dev.bytes.size
        off_ex = off_rom + (self.rom.bytes.size if self.rom is not None else 0)
        off_suf = off_ex + (self.extra.bytes.size if self.extra is not None else 0)
        pool = np.concatenate([self.dev.bytes] + ([self.rom.bytes] if self.rom is not None else []) +
                              ([self.extra.bytes] if self.extra is not None else []) + [self.suf.bytes])
        seg_src[0::2] = wstart + np.where(src == 0, off_dev, np.where(src == 1, off_rom, off_ex))
        seg_src[1::2] = sstart + off_suf
        out_start = np.zeros(2 * nwords + 1, dtype=np.int64)
        np.cumsum(seg_len, out=out_start[1:])
        total = int(out_start[-1])
        idx = np.repeat(seg_src - out_start[:-1], seg_len) + np.arange(total, dtype=np.int64)
        data = pool[idx]
        # rows: a sentence ends after its last word's suffix
        row_end = out_start[1:][1::2][is_end]
        row_end = row_end[row_end <= nbytes]
        if row_end.size == 0:
            row_end = out_start[1:][1::2][is_end][:1]
        last = int(row_end[-1])
        offs = np.concatenate([np.zeros(1, dtype=np.int64), row_end.astype(np.int64)])
        return np.ascontiguousarray(data[:last]), offs

    def generate(self, nbytes, chunk_bytes=32 << 20):
        """-> (uint8 array of ~nbytes, int64 row_offsets)"""
        parts, offs, base, i = [], [np.zeros(1, dtype=np.int64)], 0, 0
        while base < nbytes:
            d, o = self.chunk(i, min(chunk_bytes, nbytes - base))
            parts.append(d)
            offs.append(o[1:] + base)
            base += d.size
            i += 1
        return np.concatenate(parts), np.concatenate(offs)

    def lines(self, nbytes, index=0):
        d, o = self.chunk(index, nbytes)
        b = d.tobytes()
        return [b[o[i]:o[i + 1]].decode('utf-8') for i in range(len(o) - 1)]


# ---- adversarial alphabet ---------------------------------------------------
_ADV = (
    [chr(c) for c in range(0x0900, 0x0980)] + [chr(c) for c in range(0x0980, 0x0A00)] +
    list('aAbBzZeEoOkhgcstpd019 .,!?;:\'"-_@#()[]{}<>/s') +
    ['\r', '\n', '\t', '\x0b', '\x0c', '\x1c', '\x1f', '\x85', '\xa0', ' ', ' ', ' ', ' ', ' ',
     ' ', ' ', '　', '​', '‌', '‍', '­', '̀', '́', '̇', '̣',
     '̧', '़', '्', '়', '্', 'া', 'ৗ', 'ে', '؀', 'ൎ', '\U000e0020',
     '\U000e0001', 'ᄀ', 'ᅡ', 'ᆨ', '가', '각', 'ᅠ', 'ᅟ', 'ힰ', 'ퟋ',
     '😀', '👍', '🏽', '👨', '👩', '👧', '❤', '️', '⃣', '©', '🇮', '🇳', '🇺', '🇸', '🏳', '🌈',
     'É', 'é', 'İ', 'ı', 'K', 'Å', 'ß', 'Σ', 'σ', 'Ａ', 'ａ', 'ǅ', 'ẞ', 'Ω', 'Ω', 'ǅ',
     'क़', 'य़', 'ড়', 'য়', 'ऩ', 'ऱ', 'ऴ', 'ো', 'ৌ', 'ೀ', 'ೕ',
     'ཱི', 'ཱ', 'ི', 'ḋ', 'ḍ', '̈́', '͸', '\U0001f1e5', '\U0001f1e6', '\U0010ffff',
     'ഀ', '్', 'క', 'क', 'ष', '▁', '▁', 'ﬁ', '①']
)


def adversarial(n, seed, max_len=64, alphabet=None):
    """n random strings; heavy on the characters that trigger look-back rules"""
    rng = np.random.default_rng(seed)
    alpha = alphabet or _ADV
    hot = ['्', '़', '‍', '‌', 'क', 'ष', 'a', 'a', ' ', '🇮', '🇳', '👨', '̀', '́', 'e', 'E']
    out = []
    for _ in range(n):
        L = int(rng.integers(0, max_len + 1))
        mode = rng.random()
        if mode < 0.5:
            s = ''.join(alpha[int(i)] for i in rng.integers(0, len(alpha), size=L))
        elif mode < 0.8:
            s = ''.join(hot[int(i)] for i in rng.integers(0, len(hot), size=L))
        else:
            # runs: repeated characters to exercise elongation collapse and RI parity
            s = ''
            while len(s) < L:
                ch = alpha[int(rng.integers(len(alpha)))] if rng.random() < 0.5 else hot[int(rng.integers(len(hot)))]
                s += ch * int(rng.integers(1, 6))
        out.append(s)
    return out


def pack(lines):
    """list[str] -> (uint8 array, int64 row_offsets)"""
    enc = [s.encode('utf-8') for s in lines]
    offs = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in enc], out=offs[1:])
    data = np.frombuffer(b''.join(enc), dtype=np.uint8) if enc else np.zeros(0, dtype=np.uint8)
    return data.copy(), offs
