for s in 224 256 288; do for b in 4 5; do
echo -n "S=$s BPS=$b: "; MZ_FAST_S=$s MZ_FAST_BPS=$b python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --n-bases 800000000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done; done
