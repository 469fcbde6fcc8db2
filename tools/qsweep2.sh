run() { python bench.py --config $1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['result']['checksum_device_shards'][-6:])"; }
for S in 418 380 342 304; do for s in 3 4; do echo "c2 S $S sigma $s: $(MZ_FAST_S=$S MZ_FAST_QSIGMA=$s run c2)"; done; done
for s in 1.5 1 0.5 0; do echo "c4 sigma $s: $(MZ_FAST_QSIGMA=$s run c4)"; done
for s in 6 3; do echo "== perf matrix sigma $s"; MZ_FAST_QSIGMA=$s python tools/perf_matrix.py; done
