# Sweeps behind the fast kernel's queue / shared-memory geometry (device-resident bench lines).
#   bash tools/qsweep.sh slack      queue slack (MZ_FAST_QSIGMA) x shared-memory budget (MZ_FAST_SMEM_KB) on C2 / C4
#   bash tools/qsweep.sh cap        segment cap (MZ_FAST_SMAX) x slack
#   bash tools/qsweep.sh onebuf     one queue buffer (no deferred look-back) with more L1
# Variants that need another build (MZ_LIB_OUT=... MZ_NVCC_EXTRA="-DMZ_FAST_TC=4" / "-DMZ_FAST_NT=448"
# python __graft_entry__.py) are selected with MZ_B200_LIB=<that library>.
# Results of the round-2 runs: profiles/r2_queue_slack_sweep.txt.
run() { python bench.py --config $1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['result']['checksum_device_shards'][-6:])"; }
case "${1:-slack}" in
slack)
  for kb in 226 194 162 130; do for s in 6 4 3 2.5 2 1.5; do
    echo "smem $kb sigma $s: c2 $(MZ_FAST_SMEM_KB=$kb MZ_FAST_QSIGMA=$s run c2)  c4 $(MZ_FAST_SMEM_KB=$kb MZ_FAST_QSIGMA=$s run c4)"; done; done ;;
cap)
  for smax in 420 520 640; do for s in 3.5 3 2.5 2; do
    echo "smax $smax sigma $s: c2 $(MZ_FAST_SMAX=$smax MZ_FAST_QSIGMA=$s run c2)  c4 $(MZ_FAST_SMAX=$smax MZ_FAST_QSIGMA=$s run c4)"; done; done ;;
onebuf)
  echo "default: c2 $(run c2)"
  for smax in 420 520 640 800; do for kb in 130 162 194; do
    echo "nbuf 1 smax $smax smem $kb: c2 $(MZ_FAST_NBUF=1 MZ_FAST_SMAX=$smax MZ_FAST_SMEM_KB=$kb run c2)"; done; done ;;
esac
