set -e
cd akshar_b200/csrc
for cfg in "-DAKB3_MINB=8" "-DAKB3_MINB=7" "-DAKB3_MINB=6 -DAKB3_EVCAP=512"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $cfg -o ../lib/libakshar_b200.so ak_kernels.cu ak_models.cpp 2>/dev/null
  cd ../..
  echo "== $cfg"
  python bench.py --steps 5 --warmup 3 --cpu-sample-mb 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bpe', round(d['value'],2), d['ms_per_step'], d['roofline']['kernels_ms'])"
  cd akshar_b200/csrc
done
