# A/B of two library builds on the configs that matter: bash tools/ab.sh  (head = libmzb200_head.so)
for i in 1 2; do
for l in head new; do
  if [ $l = head ]; then export MZ_B200_LIB=$PWD/simd-minimizers_b200/libmzb200_head.so; else unset MZ_B200_LIB; fi
  echo "== $l"
  python tools/quick.py
done; done
for l in head new; do
  if [ $l = head ]; then export MZ_B200_LIB=$PWD/simd-minimizers_b200/libmzb200_head.so; else unset MZ_B200_LIB; fi
  echo "== $l"
  for c in c2 c4 c3; do python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['value'],1), round(d['ms_per_step'],3), d['result']['checksum_device_shards'])"; done
  python tools/perf_matrix.py 2>&1 | head -12
done
