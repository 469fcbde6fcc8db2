# dram traffic + duration of the dominant kernel for the default bench command
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mz_fast -s 3 -c 1 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" > /dev/null 2>&1
grep -E "dram__bytes|gpu__time" gpurun_out/traffic.csv | awk -F'","' '{print $(NF-2), $NF}'
