# small inputs / shards: time per launch for a range of sizes (forward k=21 w=11 and C2 parameters)
for n in 10000000 100000000 387500000 3100000000; do
  for c in c1 c2; do
    echo -n "config $c n=$n: "; python bench.py --config $c --n-bases $n --steps 10 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1000,1),'us', round(d['value'],1),'Gbp/s')"
  done
done
