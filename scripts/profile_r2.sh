#!/bin/bash
# ncu evidence of round 2 (run on the GPU box AFTER the same commands exited 0 without ncu):
#   1. launch list of the bench command (gpu__time_duration per launch, first 400 launches)
#   2. `--set full` of every hot-path kernel at BASELINE size (scripts/profile_driver.py), raw + source pages as CSV
out=${1:-gpurun_out}
mkdir -p "$out"
NCU=/usr/local/cuda/bin/ncu
python scripts/profile_driver.py > "$out/profile_driver.log" 2>&1 || { echo "driver failed"; tail -5 "$out/profile_driver.log"; exit 1; }
python bench.py --steps 8 --warmup 3 --no-cpu --skip-stages > "$out/r02_bench_short.json" 2> "$out/r02_bench_short.err" || { echo "bench failed"; exit 1; }
$NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/r02_launches_bench.csv" \
    python bench.py --steps 8 --warmup 3 --no-cpu --skip-stages > "$out/r02_ncu_launches.log" 2>&1
# the second pass of the driver: skip the launches of the first (warm-up) pass
n=$($NCU --metrics gpu__time_duration.sum --clock-control none --csv python scripts/profile_driver.py 2>/dev/null | grep -c '"gpu__time_duration.sum"')
half=$((n / 2))
$NCU --set full --clock-control none --import-source on --launch-skip "$half" -o "$out/prof_r2_all" -f \
    python scripts/profile_driver.py > "$out/r02_ncu_full.log" 2>&1
$NCU -i "$out/prof_r2_all.ncu-rep" --page raw --csv > "$out/r02_ncu_full_raw.csv" 2>/dev/null
$NCU -i "$out/prof_r2_all.ncu-rep" --page details --csv > "$out/r02_ncu_full_details.csv" 2>/dev/null
# the report itself (~150 MB) exceeds what gpurun copies back: keep the CSV exports only
rm -f "$out/prof_r2_all.ncu-rep"
echo "launches: $n (profiled the last $((n - half)))"
