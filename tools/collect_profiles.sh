#!/bin/sh
# developer tool: turn what tools/profile_round.sh TAG left in gpurun_out/ into the tracked summaries under profiles/
#   sh tools/collect_profiles.sh r02
tag=${1:-r02}
out=gpurun_out
cp $out/${tag}_bench.json profiles/${tag}_bench_1Mgames.json
cp $out/${tag}_bench_small.json profiles/${tag}_bench_16384games.json
grep -v '^==' $out/${tag}_launches.csv > profiles/${tag}_launches.csv
grep -v '^==' $out/${tag}_traffic_1M.csv > profiles/${tag}_playout_traffic_1Mgames.csv
for k in playout mccfr mccfr_pred linear_tc; do
  ncu -i $out/${tag}_$k.ncu-rep --page details > profiles/${tag}_ncu_details_$k.txt 2>/dev/null
done
python - "$tag" <<'PY'
import csv, json, sys
tag = sys.argv[1]
rows = list(csv.DictReader(open(f"profiles/{tag}_playout_traffic_1Mgames.csv")))
m = {r["Metric Name"]: float(r["Metric Value"].replace(",", "")) for r in rows}
json.dump({"games": 1048576, "dram_bytes_read": int(m["dram__bytes_read.sum"]), "dram_bytes_write": int(m["dram__bytes_write.sum"]),
           "kernel_ns_under_ncu": int(m["gpu__time_duration.sum"]),
           "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on ctd_k_playout_preset, 2^20 games in one launch (profiles/{tag}_playout_traffic_1Mgames.csv)"},
          open(f"profiles/{tag}_playout_traffic.json", "w"))
print(open(f"profiles/{tag}_playout_traffic.json").read())
PY
