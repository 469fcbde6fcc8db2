run() { python bench.py --config $1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['result']['checksum_device_shards'][-6:])"; }
echo "tc8 default: c2 $(run c2) c4 $(run c4)"
export MZ_B200_LIB=$PWD/simd-minimizers_b200/libmzb200_tc4.so
for kb in 194 178 162; do echo "tc4 smem $kb: c2 $(MZ_FAST_SMEM_KB=$kb run c2)"; done
echo "tc4 default: c4 $(run c4)"
