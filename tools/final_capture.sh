#!/bin/bash
# Round-end evidence of the bench command on ONE GPU (run under gpurun): the driver-shaped bench line of both arms first
# (no profiler), then the ncu launch list of the same command and one `--set full` capture of its merged BULK launches.
#   tools/final_capture.sh <tag>      -> gpurun_out/<tag>_*.{json,csv,ncu-rep,txt}
tag=${1:-final}
out=gpurun_out
mkdir -p $out
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> /dev/null
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
python bench.py --steps 48 --warmup 5 --no-extras --no-cpu-baseline > $out/${tag}_bench_steps48.json 2> /dev/null
ARGS="--steps 20 --warmup 5 --no-extras --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/${tag}_launches.csv \
    python bench.py $ARGS > $out/${tag}_ncu_launches.log 2>&1
python tools/ncu_summary.py launches $out/${tag}_launches.csv > $out/${tag}_launches.txt 2>&1
# the merged BULK launches of the timed region: kernel <..., 4, 1, 1> (SEG = true); skip the warm-up group
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"SpecCassieFeetPelvisArrow, double, .int.4, .int.1, .bool.1" \
    --launch-skip 1 -c 2 -f -o $out/${tag}_merged python bench.py $ARGS > $out/${tag}_ncu_merged.log 2>&1
python tools/ncu_summary.py report $out/${tag}_merged.ncu-rep > $out/${tag}_merged_full.txt 2>&1
tail -2 $out/${tag}_ncu_merged.log
head -12 $out/${tag}_launches.txt
grep -E "kernel:|time_duration|fp64_cycles_active.avg.pct_of_peak_sustained_active|dram__bytes" $out/${tag}_merged_full.txt | head -12
