run() { python bench.py --config $1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['result']['checksum_device_shards'][-6:])"; }
echo "defaults: c2 $(run c2) c3 $(run c3) c4 $(run c4) c1 $(run c1)"
for s in 3 2 1.5; do echo "sigma $s: c4 $(MZ_FAST_QSIGMA=$s run c4)"; done
python tools/perf_matrix.py
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
