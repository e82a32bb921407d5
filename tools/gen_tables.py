#!/usr/bin/env python
"""Generate the Unicode property tables used by the CUDA kernels and by the oracle.

Every table is produced by PROBING the module the reference actually calls for that step
 -- never from a UCD download and never from a different module:

  * grapheme-cluster classes (GCB, InCB, ExtPict) and `\\s`  <- `regex`   (reference: segment.py:14, normalize.py:97-103)
  * ccc, canonical decompositions, composition pairs, NFC_QC,
    `'LATIN' in name` + `lower()`, `isdigit()`               <- CPython `unicodedata`/`str` (normalize.py:18,37-39; segment.py:139)
  * HF `Whitespace` pre-tokenizer classes (\\w / \\s / other)   <- `tokenizers` (scripts/train_bpe.py:74)

Outputs (both committed, so neither the GPU box nor the tests need to re-probe):
  akshar_b200/csrc/unicode_tables.inc   C arrays compiled into libakshar_b200.so
  oracle/ucd_tables.json                range lists read by oracle/akshar_oracle.py

Run here (dev container):  python tools/gen_tables.py
"""
import json
import os
import sys
import unicodedata

import regex

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NCP = 0x110000

GCB_NAMES = ["Other", "CR", "LF", "Control", "Extend", "ZWJ", "Regional_Indicator", "Prepend",
             "SpacingMark", "L", "V", "T", "LV", "LVT"]
INCB_NAMES = ["None", "Consonant", "Linker", "Extend"]
TAG_NAMES = ["devanagari", "roman", "digit", "punct", "other"]


def probe_regex(prop):
    allc = ''.join(chr(c) for c in range(NCP))
    out = bytearray(NCP)
    for m in regex.finditer(r'\p{%s}' % prop, allc):
        out[m.start()] = 1
    return out


def probe_regex_class(pattern):
    allc = ''.join(chr(c) for c in range(NCP))
    out = bytearray(NCP)
    for m in regex.finditer(pattern, allc):
        out[m.start()] = 1
    return out


def to_ranges(flags):
    """list of [lo, hi] inclusive where flags[cp] is truthy"""
    r = []
    start = None
    for cp in range(NCP):
        if flags[cp]:
            if start is None:
                start = cp
        elif start is not None:
            r.append([start, cp - 1])
            start = None
    if start is not None:
        r.append([start, NCP - 1])
    return r


def value_ranges(vals):
    """list of [lo, hi, v] for maximal runs of equal non-zero value"""
    r = []
    cp = 0
    while cp < NCP:
        v = vals[cp]
        if v:
            lo = cp
            while cp + 1 < NCP and vals[cp + 1] == v:
                cp += 1
            r.append([lo, cp, int(v)])
        cp += 1
    return r


def main():
    ver = {
        "unicodedata": unicodedata.unidata_version,
        "regex": regex.__version__,
        "python": sys.version.split()[0],
    }
    import tokenizers
    ver["tokenizers"] = tokenizers.__version__
    print("versions", ver)

    # ---- grapheme properties from `regex` (Unicode 17.0 in regex 2026.3.32)
    gcb = bytearray(NCP)
    for i, name in enumerate(GCB_NAMES):
        if i == 0:
            continue
        fl = probe_regex("GCB=" + name)
        for cp in range(NCP):
            if fl[cp]:
                assert gcb[cp] == 0
                gcb[cp] = i
    incb = bytearray(NCP)
    for i, name in enumerate(INCB_NAMES):
        if i == 0:
            continue
        fl = probe_regex("InCB=" + name)
        for cp in range(NCP):
            if fl[cp]:
                assert incb[cp] == 0
                incb[cp] = i
    extpict = probe_regex("ExtPict")
    rx_space = probe_regex_class(r'\s')
    # the reference's allow-list, probed through the very pattern it compiles (normalize.py:97-103)
    allow = probe_regex_class(r'[ऀ-ॿঀ-৿a-zA-Z0-9\s.,!?;:\'\"\-]')
    # `.` in remove_elongations excludes only \n (normalize.py:56)
    dot = probe_regex_class(r'.')
    not_dot = [cp for cp in range(NCP) if not dot[cp]]
    assert not_dot == [0x0A], not_dot

    # ---- CPython unicodedata / str (Unicode 15.0)
    ccc = bytearray(NCP)
    nfc_qc = bytearray(NCP)  # 0 yes 1 no 2 maybe
    decomp = {}              # cp -> full canonical decomposition (list), non-Hangul only
    pairs = {}               # (a,b) -> composite
    latin_lower = {}         # cp -> list of cps, only where 'LATIN' in name and lower() changes
    full_lower = {}          # cp -> list of cps where str.lower() changes (used by roman_phonetic_signature)
    isdigit = bytearray(NCP)
    for cp in range(NCP):
        ch = chr(cp)
        ccc[cp] = unicodedata.combining(ch)
        if ch.isdigit():
            isdigit[cp] = 1
        lo = ch.lower()
        if lo != ch:
            full_lower[cp] = [ord(x) for x in lo]
            if 'LATIN' in unicodedata.name(ch, ''):
                latin_lower[cp] = [ord(x) for x in lo]
        if 0xD800 <= cp <= 0xDFFF:
            continue
        nfc = unicodedata.normalize('NFC', ch)
        if nfc != ch:
            nfc_qc[cp] = 1
        if 0xAC00 <= cp <= 0xD7A3:
            continue  # Hangul syllables are algorithmic
        d = unicodedata.decomposition(ch)
        if d and not d.startswith('<'):
            nfd = unicodedata.normalize('NFD', ch)
            decomp[cp] = [ord(x) for x in nfd]
            one = [int(x, 16) for x in d.split()]
            if len(one) == 2 and nfc == ch:
                pairs[(one[0], one[1])] = cp
    for (a, b), c in pairs.items():
        nfc_qc[b] = 2 if nfc_qc[b] == 0 else nfc_qc[b]
    for cp in list(range(0x1161, 0x1176)) + list(range(0x11A8, 0x11C3)):
        nfc_qc[cp] = 2  # Hangul V / T jamo compose backwards
    # every multi-cp lowercase in CPython: only U+0130
    multi = {cp: v for cp, v in full_lower.items() if len(v) > 1}
    assert multi == {0x130: [0x69, 0x307]}, multi
    max_dec = max(len(v) for v in decomp.values())
    print("decomp", len(decomp), "max len", max_dec, "pairs", len(pairs), "latin_lower", len(latin_lower),
          "full_lower", len(full_lower), "qc_no", sum(1 for x in nfc_qc if x == 1), "qc_maybe",
          sum(1 for x in nfc_qc if x == 2), "ccc!=0", sum(1 for x in ccc if x))

    # ---- Final_Sigma context classes of str.lower() (normalize.py:72 `word.lower()`), probed through lower() itself
    case_ign = bytearray(NCP)
    cased = bytearray(NCP)          # cased AND not case-ignorable (the only place CPython consults `cased`)
    for cp in range(NCP):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        ch = chr(cp)
        a = ('a\u03a3' + ch).lower()[1]
        b = ('a\u03a3' + ch + 'a').lower()[1]
        if a == '\u03c2' and b == '\u03c3':
            case_ign[cp] = 1
        elif a == '\u03c3':
            cased[cp] = 1
    print("case_ignorable", sum(case_ign), "cased", sum(cased))

    # ---- script tag, exactly the order of segment.py:128-147
    tag = bytearray(NCP)
    punct = set(' .,!?;:\'"()-[]{}')
    for cp in range(NCP):
        if 0x900 <= cp <= 0x97F:
            t = 0
        elif 0x41 <= cp <= 0x5A or 0x61 <= cp <= 0x7A:
            t = 1
        elif isdigit[cp]:
            t = 2
        elif chr(cp) in punct:
            t = 3
        else:
            t = 4
        tag[cp] = t

    # ---- HF Whitespace pre-tokenizer classes: 0 = punct-like, 1 = \w, 2 = \s
    from tokenizers import pre_tokenizers
    pt = pre_tokenizers.Whitespace()
    hf = bytearray(NCP)
    for cp in range(NCP):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        pieces = [p for p, _ in pt.pre_tokenize_str('a' + chr(cp) + 'a')]
        if len(pieces) == 1:
            hf[cp] = 1
        elif len(pieces) == 2:
            hf[cp] = 2
        else:
            assert len(pieces) == 3, (hex(cp), pieces)
            hf[cp] = 0
    print("hf w", sum(1 for x in hf if x == 1), "s", sum(1 for x in hf if x == 2))
    # HF NFKC on the closed alphabet: identity except the exotic spaces -> U+0020
    from tokenizers import normalizers
    nfkc = normalizers.NFKC()
    nfkc_changes = {}
    for cp in range(NCP):
        if allow[cp] and nfc_qc[cp] != 1 and not (0x41 <= cp <= 0x5A):
            s = nfkc.normalize_str(chr(cp))
            if s != chr(cp):
                nfkc_changes[cp] = [ord(x) for x in s]
    print("HF NFKC changes inside closed alphabet:", {hex(k): [hex(x) for x in v] for k, v in nfkc_changes.items()})
    for cp, v in nfkc_changes.items():
        assert v == [0x20] and hf[cp] == 2, hex(cp)

    # ---- code points the BPE encoder may meet without a full NFKC: HF's NFKC (Unicode <= 12) must act on them exactly
    # like the NFC the kernels implement (Unicode 15), alone and next to marks; whitespace that NFKC folds to U+0020
    # is fine too (never inside a word).  '<' stays out: it starts every added token ("<s>", "</s>", ...), which HF
    # matches in the raw text before normalizing.
    comp_first0 = set(a for (a, b) in pairs) | set(range(0x1100, 0x1113))
    ctxs = [lambda c: c, lambda c: 'a' + c, lambda c: c + '\u0301', lambda c: 'a' + c + '\u0301', lambda c: 'a\u0301' + c,
            lambda c: '\u0915' + c, lambda c: c + '\u093c', lambda c: '\u0323' + c]
    bpe_safe = bytearray(NCP)
    for cp in range(NCP):
        if 0xD800 <= cp <= 0xDFFF or cp == 0x3C:
            continue
        ch = chr(cp)
        k = nfkc.normalize_str(ch)
        ok = k == unicodedata.normalize('NFC', ch)
        if not ok and hf[cp] == 2 and k == ' ':
            bpe_safe[cp] = 1
            continue
        if ok and (ccc[cp] or nfc_qc[cp] or cp in decomp or cp in comp_first0 or 0x1100 <= cp <= 0x11FF or 0xAC00 <= cp <= 0xD7A3):
            for f in ctxs:
                t = f(ch)
                if nfkc.normalize_str(t) != unicodedata.normalize('NFC', t):
                    ok = False
                    break
        if ok:
            bpe_safe[cp] = 1
    print("bpe_safe", sum(bpe_safe), "allow-list members not safe:", [hex(c) for c in range(NCP) if allow[c] and not bpe_safe[c] and nfc_qc[c] != 1])

    # ---- HF's NFKC, for the rows the BPE encoder normalizes itself (scripts/train_bpe.py:71): probed from the `tokenizers`
    # normalizers, whose Unicode data is older than CPython's.
    #   hf_kmap[cp]   = HF NFKD(cp) wherever it differs from the canonical decomposition above (compatibility mappings)
    #   hf_unknown    = code points with normalization properties here (ccc / canonical decomposition) that HF does not
    #                   know: it passes them through as inert starters
    nfkd, nfd = normalizers.NFKD(), normalizers.NFD()
    hf_kmap = {}
    hf_unknown = set()
    for cp in range(NCP):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        ch = chr(cp)
        ours = unicodedata.normalize('NFD', ch)
        if nfd.normalize_str(ch) != ours:
            hf_unknown.add(cp)
        else:
            h = nfkd.normalize_str(ch)
            if h != ours:
                hf_kmap[cp] = [ord(c) for c in h]
        cc = unicodedata.combining(ch)
        if cc:
            t = ('a' + ch + '\u0334') if cc > 1 else ('a\u0301' + ch)
            if nfd.normalize_str(t) != unicodedata.normalize('NFD', t):
                hf_unknown.add(cp)
    for cp in hf_unknown:
        assert nfkc.normalize_str(chr(cp)) == chr(cp), hex(cp)
    for cp, v in hf_kmap.items():
        assert not any(x in hf_unknown for x in v), hex(cp)
        assert not bpe_safe[cp] or v == [0x20], hex(cp)
    # rows with such a code point, or with the pieces one of them decomposes into (HF does not compose those), are
    # normalized by the exact row path
    for cp in hf_unknown:
        bpe_safe[cp] = 0
        for x in decomp.get(cp, []):
            bpe_safe[x] = 0
    print("hf_kmap", len(hf_kmap), "data", sum(len(v) for v in hf_kmap.values()), "hf_unknown", len(hf_unknown))

    # ---- pack the 32-bit property word
    # first element of some canonical composition pair (incl. the algorithmic Hangul L + V; LV + T is covered by
    # the decomposition bit): a mark after such a starter may compose, after any other atomic starter it cannot
    comp_first = set(a for (a, b) in pairs) | set(range(0x1100, 0x1113))
    props = [0] * NCP
    for cp in range(NCP):
        w = gcb[cp] | (incb[cp] << 4) | (extpict[cp] << 6) | (tag[cp] << 7) | (allow[cp] << 10)
        w |= (nfc_qc[cp] << 11) | ((1 if cp in latin_lower else 0) << 13) | (hf[cp] << 14) | (ccc[cp] << 16)
        w |= ((1 if cp in full_lower else 0) << 24) | ((1 if (cp in decomp or 0xAC00 <= cp <= 0xD7A3) else 0) << 25)
        w |= (case_ign[cp] << 26) | (cased[cp] << 27) | ((1 if cp in comp_first else 0) << 28) | (bpe_safe[cp] << 29)
        props[cp] = w
    pages = {}
    page_index = []
    leaves = []
    for p in range(NCP >> 8):
        key = tuple(props[p << 8:(p + 1) << 8])
        if key not in pages:
            pages[key] = len(pages)
            leaves.extend(key)
        page_index.append(pages[key])
    print("unique pages", len(pages), "leaf bytes", len(leaves) * 4)

    # ---- emit C
    def arr(name, ctype, vals, per_line=16, fmt="%d"):
        lines = ["const %s %s[%d] = {" % (ctype, name, len(vals))]
        for i in range(0, len(vals), per_line):
            lines.append("  " + ",".join(fmt % v for v in vals[i:i + per_line]) + ",")
        lines.append("};")
        return "\n".join(lines)

    dec_keys = sorted(decomp)
    dec_off = []
    dec_data = []
    for k in dec_keys:
        dec_off.append(len(dec_data))
        dec_data.extend(decomp[k])
    dec_off.append(len(dec_data))
    pair_items = sorted(pairs.items())
    pair_keys = [(a << 21) | b for (a, b), _ in pair_items]
    pair_vals = [c for _, c in pair_items]
    ll_keys = sorted(latin_lower)
    ll_vals = [latin_lower[k][0] for k in ll_keys]
    fl_keys = sorted(full_lower)
    fl_vals = [full_lower[k][0] for k in fl_keys]
    km_keys = sorted(hf_kmap)
    km_off = []
    km_data = []
    for k in km_keys:
        km_off.append(len(km_data))
        km_data.extend(hf_kmap[k])
    km_off.append(len(km_data))
    unk_keys = sorted(hf_unknown)

    out = []
    out.append("// GENERATED by tools/gen_tables.py -- do not edit.")
    out.append("// sources: regex %s (grapheme props), CPython %s unicodedata %s (NFC/lower/isdigit), tokenizers %s (pre-tokenizer classes)"
               % (ver["regex"], ver["python"], ver["unicodedata"], ver["tokenizers"]))
    out.append("// property word: [0:4) GCB  [4:6) InCB  [6] ExtPict  [7:10) script tag  [10] allow-list  [11:13) NFC_QC (0 yes,1 no,2 maybe)")
    out.append("//   [13] latin-lower changes  [14:16) HF pretok class (0 other,1 \\w,2 \\s)  [16:24) ccc  [24] str.lower changes  [25] has canonical decomposition  [26] case-ignorable  [27] cased (and not case-ignorable)  [28] first of a composition pair  [29] BPE-safe (HF NFKC acts like NFC)")
    out.append("#define AK_N_PAGES %d" % len(page_index))
    out.append("#define AK_N_LEAF_PAGES %d" % len(pages))
    out.append("#define AK_N_DECOMP %d" % len(dec_keys))
    out.append("#define AK_N_DECOMP_DATA %d" % len(dec_data))
    out.append("#define AK_N_PAIRS %d" % len(pair_keys))
    out.append("#define AK_N_LATIN_LOWER %d" % len(ll_keys))
    out.append("#define AK_N_FULL_LOWER %d" % len(fl_keys))
    out.append("#define AK_N_KMAP %d" % len(km_keys))
    out.append("#define AK_N_KMAP_DATA %d" % len(km_data))
    out.append("#define AK_N_HF_UNKNOWN %d" % len(unk_keys))
    out.append(arr("ak_tbl_page_index", "unsigned short", page_index))
    out.append(arr("ak_tbl_leaves", "unsigned int", leaves, 8, "0x%xu"))
    out.append(arr("ak_tbl_decomp_keys", "unsigned int", dec_keys, 12, "0x%x"))
    out.append(arr("ak_tbl_decomp_off", "unsigned short", dec_off))
    out.append(arr("ak_tbl_decomp_data", "unsigned int", dec_data, 12, "0x%x"))
    out.append(arr("ak_tbl_pair_keys", "unsigned long long", pair_keys, 6, "0x%xull"))
    out.append(arr("ak_tbl_pair_vals", "unsigned int", pair_vals, 12, "0x%x"))
    out.append(arr("ak_tbl_latin_lower_keys", "unsigned int", ll_keys, 12, "0x%x"))
    out.append(arr("ak_tbl_latin_lower_vals", "unsigned int", ll_vals, 12, "0x%x"))
    out.append(arr("ak_tbl_full_lower_keys", "unsigned int", fl_keys, 12, "0x%x"))
    out.append(arr("ak_tbl_full_lower_vals", "unsigned int", fl_vals, 12, "0x%x"))
    out.append(arr("ak_tbl_kmap_keys", "unsigned int", km_keys, 12, "0x%x"))
    out.append(arr("ak_tbl_kmap_off", "unsigned short", km_off))
    out.append(arr("ak_tbl_kmap_data", "unsigned int", km_data, 12, "0x%x"))
    out.append(arr("ak_tbl_hf_unknown", "unsigned int", unk_keys, 12, "0x%x"))
    inc = os.path.join(ROOT, "akshar_b200", "csrc", "unicode_tables.inc")
    with open(inc, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", inc, os.path.getsize(inc))

    # ---- emit oracle JSON (range lists; a different representation of the same probes)
    oj = {
        "versions": ver,
        "gcb_names": GCB_NAMES,
        "incb_names": INCB_NAMES,
        "tag_names": TAG_NAMES,
        "gcb": value_ranges(gcb),
        "incb": value_ranges(incb),
        "extpict": to_ranges(extpict),
        "regex_space": to_ranges(rx_space),
        "allow": to_ranges(allow),
        "isdigit": to_ranges(isdigit),
        "ccc": value_ranges(ccc),
        "nfc_qc": value_ranges(nfc_qc),
        "decomp": {"%x" % k: decomp[k] for k in dec_keys},
        "pairs": [[a, b, c] for (a, b), c in pair_items],
        "latin_lower": {"%x" % k: latin_lower[k] for k in ll_keys},
        "full_lower": {"%x" % k: full_lower[k] for k in fl_keys},
        "hf_class": value_ranges(hf),
        "case_ignorable": to_ranges(case_ign),
        "cased": to_ranges(cased),
        "bpe_safe": to_ranges(bpe_safe),
        "hf_kmap": {"%x" % k: hf_kmap[k] for k in km_keys},
        "hf_unknown": unk_keys,
    }
    oj_path = os.path.join(ROOT, "oracle", "ucd_tables.json")
    with open(oj_path, "w") as f:
        json.dump(oj, f, separators=(",", ":"))
    print("wrote", oj_path, os.path.getsize(oj_path))


if __name__ == "__main__":
    main()
