"""Oracle results in the ragged layouts the CUDA path returns (test helper)."""
import numpy as np

import akshar_oracle as O


def normalize_batch(lines, normalize_roman=True, clean_hinglish=True):
    outs = [O.normalize_text(s, normalize_roman, clean_hinglish).encode('utf-8') for s in lines]
    off = np.zeros(len(outs) + 1, dtype=np.int64)
    np.cumsum([len(o) for o in outs], out=off[1:])
    return np.frombuffer(b''.join(outs), dtype=np.uint8), off


def segment_batch(lines, matras=False):
    ends, splits = [], [0]
    for s in lines:
        cps = [ord(c) for c in s]
        e = O.cp_ends_to_byte_ends(cps, O.segment_breaks(cps, matras))
        ends.extend(e)
        splits.append(len(ends))
    return np.array(ends, dtype=np.int32), np.array(splits, dtype=np.int64)


def runs_batch(lines):
    ends, tags, splits = [], [], [0]
    for s in lines:
        cps = [ord(c) for c in s]
        r = O.script_runs(cps)
        ends.extend(O.cp_ends_to_byte_ends(cps, [e for e, _ in r]))
        tags.extend(255 if t is None else t for _, t in r)
        splits.append(len(ends))
    return np.array(ends, dtype=np.int32), np.array(tags, dtype=np.uint8), np.array(splits, dtype=np.int64)
