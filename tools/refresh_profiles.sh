# Re-measure everything profiles/ holds for the current build (one B200).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_c2.err | tail -1 > gpurun_out/bench_c2.json
for c in c1 c3 c4; do python bench.py --config $c --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_$c.json; done
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
bash tools/traffic.sh > gpurun_out/traffic.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:mz_fast -s 2 -c 1 -o gpurun_out/prof_fast_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --n-bases 800000000 > /dev/null 2>&1
python tools/amb_bench.py > gpurun_out/amb_bench.txt 2>&1
python tools/bench_c5.py > gpurun_out/bench_c5.txt 2>&1
python tools/perf_matrix.py > gpurun_out/perf_matrix.txt 2>&1
ls -la gpurun_out | tail -15
