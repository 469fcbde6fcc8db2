# Sweep of the queue slack (MZ_FAST_QSIGMA) and the segment cap (MZ_FAST_SMAX) on C2 / C4 / C3 (device-resident).
for smax in 420 520 640; do for s in 3.5 3 2.5 2 1.5; do
  line="smax $smax sigma $s:"
  for c in c2 c4; do
    r=$(MZ_FAST_SMAX=$smax MZ_FAST_QSIGMA=$s python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['result']['checksum_device_shards'][-6:])")
    line="$line  $c $r"
  done
  echo "$line"
done; done
