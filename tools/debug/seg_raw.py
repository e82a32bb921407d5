import sys, os, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import numpy as np, synth_corpus as sc, akshar_b200 as A
from akshar_b200 import batch as Bm
eng = Bm.engine(0)
for kind in ('social', 'hinglish'):
    data, off = sc.Corpus(kind, 3).generate(128 << 20)
    tb = eng.put((torch.from_numpy(data), torch.from_numpy(off)))
    for name, fn in (('segment raw', lambda: eng.segment_batch(tb, clusters=True, runs=True)),):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3): fn()
        torch.cuda.synchronize()
        print(kind, name, '%.1f GB/s' % (3 * data.size / (time.perf_counter() - t) / 1e9))
    norm = eng.normalize_batch(tb)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3): eng.segment_batch(norm, clusters=True, runs=True)
    torch.cuda.synchronize()
    print(kind, 'segment normalized', '%.1f GB/s' % (3 * data.size / (time.perf_counter() - t) / 1e9))
