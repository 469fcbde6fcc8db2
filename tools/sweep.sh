# e2e (pinned host in/out through mz_run) vs chunk size of the pipelined host path, C2
run() { echo -n "$*: "; env "$@" python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['e2e']['ms_per_step'],2),'ms', round(d['e2e']['value'],1),'Gbp/s')"; }
run MZ_CHUNK_WINDOWS=129166666
run MZ_CHUNK_WINDOWS=64583333
run MZ_CHUNK_WINDOWS=33554432
run MZ_CHUNK_WINDOWS=258333333
