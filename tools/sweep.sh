# e2e A/B: scalars through mapped host memory (default) vs D2H copy node (MZ_HS_COPY=1)
run() { echo -n "$*: "; env "$@" python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['e2e']['ms_per_step'],2),'ms', round(d['e2e']['value'],1),'Gbp/s')"; }
run A=1
run MZ_HS_COPY=1
run A=1
run MZ_HS_COPY=1
