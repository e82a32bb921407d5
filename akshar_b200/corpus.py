"""File / corpus front end of the hot path: the callers on either side of the batch kernels.

Mirrors, on the batch path, what the reference does one string at a time:
  * `preprocess_corpus`  = reference cli.py:165-190 (= scripts/train_bpe.py:16-35, train_spm.py:18-44): read lines,
    `strip()`, skip the empty ones, `normalize_text` each, write them back one per line;
  * `tokenize_file`      = reference cli.py:46-84 `akshar tokenize -i FILE --format text|json|id`: the WHOLE file is one
    string (one very long row for the kernels);
  * `encode_lines`       = new: one row per non-empty line, ragged ids back.
Line splitting / stripping stay Python's own (`readlines`, `str.strip`): they define what a row is.
"""
import json

from .normalize import normalize_batch


def read_rows(input_file):
    with open(input_file, 'r', encoding='utf-8') as f:
        lines = f.readlines()
    rows = []
    for line in lines:
        line = line.strip()
        if line:
            rows.append(line)
    return rows


def preprocess_corpus(input_file, output_file, normalize_roman=True, clean_hinglish=True):
    rows = read_rows(input_file)
    print(f"Preprocessing {input_file}...")
    processed = normalize_batch(rows, normalize_roman, clean_hinglish) if rows else []
    with open(output_file, 'w', encoding='utf-8') as f:
        for line in processed:
            f.write(line + '\n')
    print(f"Wrote {len(processed)} lines to {output_file}")
    return str(output_file)


def tokenize_file(tokenizer, input_file, fmt='text', output_file=None):
    """-> the string the reference CLI would print / write"""
    with open(input_file, 'r', encoding='utf-8') as f:
        text = f.read()
    if fmt == 'id':
        if tokenizer.model is None:
            raise ValueError("need model for IDs")
        output = ' '.join(map(str, tokenizer.encode_batch([text])[0]))
    else:
        tokens = tokenizer.tokenize_batch([text])[0]
        output = json.dumps(tokens, ensure_ascii=False, indent=2) if fmt == 'json' else ' '.join(tokens)
    if output_file:
        with open(output_file, 'w', encoding='utf-8') as f:
            f.write(output)
    return output


def encode_lines(tokenizer, input_file, as_device=False):
    """one row per non-empty stripped line -> list[list[int]] (or the device Ragged with as_device=True)"""
    return tokenizer.encode_batch(read_rows(input_file), as_device=as_device)
