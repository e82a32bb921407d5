#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/sharded_check.py: aksharTokenizer.encode_batch_sharded over N GPUs == the same batch
encoded by one GPU (rank 0 checks and prints)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tools')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import akshar_b200 as A  # noqa: E402
import synth_corpus as sc  # noqa: E402


def main():
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl')
    data, off = sc.Corpus('hinglish', 5).generate(96 << 20)
    for name, kind in (('bpe24k.json', 'bpe'), ('spm24k.model', 'sentencepiece')):
        tk = A.aksharTokenizer(os.path.join(ROOT, 'tests', 'golden', 'models', name), kind, device=local)
        ids, sp = tk.encode_batch_sharded(data, off)
        if rank == 0:
            ref, _ = tk._eng.tokenizer_encode_batch((torch.from_numpy(data), torch.from_numpy(off)), tk.model.kind)
            ok = np.array_equal(ids, ref.values.cpu().numpy()) and np.array_equal(sp, ref.splits.cpu().numpy())
            print('sharded == single GPU (%s, %d ranks, %d rows, %d ids): %s' % (kind, dist.get_world_size(), off.size - 1, ids.size, ok))
            assert ok
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
