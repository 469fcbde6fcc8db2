run() { python bench.py --config $1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['result']['checksum_device_shards'][-6:])"; }
echo "default: c2 $(run c2)"
for smax in 420 520 640 800; do for kb in 130 162 194; do echo "nbuf 1 smax $smax smem $kb: c2 $(MZ_FAST_NBUF=1 MZ_FAST_SMAX=$smax MZ_FAST_SMEM_KB=$kb run c2)"; done; done
