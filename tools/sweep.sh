# S / staging-list sweep of the fast kernel, 3.1 Gbp C2, device-resident
run() { echo -n "$*: "; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
run MZ_FAST_S=304
run MZ_FAST_S=304 MZ_FAST_LISTF=1.0
run MZ_FAST_S=304 MZ_FAST_LISTF=0.9
run MZ_FAST_S=304 MZ_FAST_LISTF=0.85
run MZ_FAST_S=266 MZ_FAST_LISTF=1.0
run MZ_FAST_S=228 MZ_FAST_LISTF=1.1
