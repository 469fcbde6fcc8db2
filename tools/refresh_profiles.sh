# Re-measure everything profiles/ holds for the current build (one B200).  Outputs land in
# gpurun_out/ under the names profiles/ uses (tag = $1, default r2); copy them over afterwards.
T=${1:-r2}
O=gpurun_out
set -x
mkdir -p $O
python bench.py --steps 5 --warmup 3 2> $O/${T}_bench_c2_n1.err | tail -1 > $O/${T}_bench_c2_n1.json
for c in c1 c3 c4 c5; do python bench.py --config $c --steps 5 --warmup 3 2>/dev/null | tail -1 > $O/${T}_bench_${c}_n1.json; done
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/${T}_bench_reference_arm_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/${T}_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none -c 40 --csv --log-file $O/${T}_launches_bench_c1.csv python bench.py --config c1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
bash tools/traffic.sh > $O/${T}_traffic_c2.txt 2>&1; cp $O/traffic.csv $O/${T}_traffic_c2_full.csv
ncu --set full --clock-control none --import-source on -k regex:mz_fast -s 2 -c 1 -o $O/${T}_prof_fast -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --n-bases 800000000 > /dev/null 2>&1
python tools/amb_bench.py > $O/${T}_skip_ambiguous.txt 2>&1
python tools/perf_matrix.py > $O/${T}_perf_matrix.txt 2>&1
ls -la $O | tail -15
