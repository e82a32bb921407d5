python - <<'PY'
import torch, time
x = torch.empty(1<<30, dtype=torch.uint8).pin_memory(); d = torch.empty(1<<30, dtype=torch.uint8, device='cuda')
y = torch.empty(1<<30, dtype=torch.uint8).pin_memory(); d2 = torch.empty(1<<30, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in (('h2d', lambda: d.copy_(x, non_blocking=True)), ('d2h', lambda: y.copy_(d2, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); print(name, 3*(1<<30)/(time.perf_counter()-t)/1e9, 'GB/s')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print('both', 3*(1<<30)/(time.perf_counter()-t)/1e9, 'GB/s each direction')
PY
for mb in 32 64 128 256; do
  echo "== chunk $mb"
  AKSHAR_CHUNK_MB=$mb python bench.py --steps 3 --warmup 3 --cpu-sample-mb 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bpe', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['e2e']['ms_per_step'])"
done
