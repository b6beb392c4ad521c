#!/bin/sh
# Round profile recipe (run under gpurun): plain runs first, then the ncu launch list and one --set full capture per hot kernel.
#   sh tools/profile_round.sh TAG
tag=${1:-rXX}
out=gpurun_out
set -x
python bench.py --steps 3 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
python bench.py --games 16384 --steps 2 --warmup 1 --mccfr-roots 512 --no-cpu-baseline > $out/${tag}_bench_small.json 2>> $out/${tag}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --games 16384 --steps 2 --warmup 1 --mccfr-roots 512 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
python tools/playout_perf.py 16384 > $out/${tag}_playout_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:ctd_k_playout -s 1 -c 1 -f -o $out/${tag}_playout \
  python tools/playout_perf.py 16384 > $out/${tag}_ncu_playout.log 2>&1
python tools/playout_perf.py 1048576 > $out/${tag}_playout_1M_plain.log 2>&1 || exit 1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ctd_k_playout -s 1 -c 1 --csv \
  --log-file $out/${tag}_traffic_1M.csv python tools/playout_perf.py 1048576 > $out/${tag}_ncu_traffic.log 2>&1
REPS=1 python tools/mccfr_perf.py 2048 200 both > $out/${tag}_mccfr_plain.log 2>&1 || exit 1
REPS=1 ncu --set full --clock-control none --import-source on -k regex:'ctd_k_mccfr(_preset)?$' -c 1 -f -o $out/${tag}_mccfr \
  python tools/mccfr_perf.py 2048 200 pure > $out/${tag}_ncu_mccfr.log 2>&1
REPS=1 ncu --set full --clock-control none --import-source on -k regex:ctd_k_mccfr_pred -c 1 -f -o $out/${tag}_mccfr_pred \
  python tools/mccfr_perf.py 2048 200 deep > $out/${tag}_ncu_mccfr_pred.log 2>&1
REPS=1 CTD_BACKEND=tcgen05 python tools/mccfr_perf.py 2048 200 deep > $out/${tag}_tc_plain.log 2>&1 || exit 1
REPS=1 CTD_BACKEND=tcgen05 ncu --set full --clock-control none --import-source on -k regex:ctd_k_linear_tc_tma -s 6 -c 1 -f -o $out/${tag}_linear_tc \
  python tools/mccfr_perf.py 2048 200 deep > $out/${tag}_ncu_tc.log 2>&1
ls -la $out | tail -20
