import sys, os, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import synth_corpus as sc
from akshar_b200 import batch as Bm
eng = Bm.engine(0)
data, off = sc.Corpus('social', 3).generate(128 << 20)
tb = eng.put((torch.from_numpy(data), torch.from_numpy(off)))
for name, kw in (('default', {}), ('clean_hinglish=False', {'clean_hinglish': False})):
    eng.normalize_batch(tb, **kw); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3): eng.normalize_batch(tb, **kw)
    torch.cuda.synchronize()
    print('normalize', name, '%.1f GB/s' % (3 * data.size / (time.perf_counter() - t) / 1e9))
