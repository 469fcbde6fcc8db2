# S sweep (windows per thread) of the fast kernel, C2 3.1 Gbp device-resident
run() { echo -n "$*: "; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
run MZ_FAST_S=285
run MZ_FAST_S=304
run MZ_FAST_S=323
run MZ_FAST_S=342
run MZ_FAST_S=361
run MZ_FAST_S=380
