#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box, ONE GPU):  tools/gpu_profile.sh TAG
#   gpurun_out/TAG_launches_bench.csv      per-launch gpu__time_duration.sum of bench.py --steps 2 --warmup 3 --fast
#   gpurun_out/TAG_all_kernels.ncu-rep     ncu --set full of every kernel of the LAST a3_detect_batch call of tools/profile_target.py
# Both commands are first run without ncu and must exit 0.  Summaries: tools/ncu_summary.py (here, no GPU needed).
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 2 --warmup 3 --fast --no-cpu-baseline > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --fast --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/profile_target.py 256 3 > $OUT/${TAG}_target_plain.log 2>&1 || { echo "profile_target failed"; exit 1; }
# launch list of the target: where the third call's K1 starts
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_target.csv \
    python tools/profile_target.py 256 3 > $OUT/${TAG}_ncu_target.log 2>&1
SKIP=$(python - "$OUT/${TAG}_launches_target.csv" <<'EOF'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 4 and r[0].isdigit()]
k1 = [i for i, r in enumerate(rows) if "k1_strips_kernel" in ",".join(r)]
print(k1[-1] if k1 else 0, len(rows) - (k1[-1] if k1 else 0))
EOF
)
set -- $SKIP
echo "full capture: skip $1 launches, take $2"
ncu --set full --clock-control none --import-source on --launch-skip $1 -c $2 -f -o $OUT/${TAG}_all_kernels \
    python tools/profile_target.py 256 3 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_*
