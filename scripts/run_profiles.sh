#!/bin/bash
# Profiles of the final kernels, run on the GPU box:  gpurun --timeout 1500 -- 'bash scripts/run_profiles.sh r02d'
# Every program runs once WITHOUT ncu first (and must exit 0); numbers printed under ncu are never bench values.
# Outputs land in gpurun_out/<tag>_*; scripts/summarize_ncu.py turns the .ncu-rep files into profiles/<tag>_*.csv/.json.
set -u
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
run_ok() { "$@" > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed: $*"; tail -5 $OUT/${TAG}_plain.log; exit 1; }; }

BENCH="python bench.py --steps 2 --warmup 1 --no-cpu"
run_ok $BENCH
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/${TAG}_launches_bench_step.csv $BENCH > $OUT/ncu1.log 2>&1
ENVB="python bench.py --workload env_step --steps 2 --warmup 1 --no-cpu"
run_ok $ENVB
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/${TAG}_launches_env_step.csv $ENVB > $OUT/ncu2.log 2>&1
ENVB256="python bench.py --workload env_step --size 256 --steps 2 --warmup 1 --no-cpu"
run_ok $ENVB256
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/${TAG}_launches_env_step_256.csv $ENVB256 > $OUT/ncu3.log 2>&1

SEQ="python bench.py --steps 1 --warmup 1 --no-cpu --sequential"
run_ok $SEQ
$NCU --set full --import-source on -k regex:'gather_kernel|detect_kernel|make_effective' -c 12 -f -o $OUT/${TAG}_gather $SEQ > $OUT/ncu4.log 2>&1
$NCU --set full --import-source on -k regex:'sweep2|presample|scatter_events|bucket_scan' -c 8 -f -o $OUT/${TAG}_scan $SEQ > $OUT/ncu5.log 2>&1
ENV1="python bench.py --workload env_step --size 256 --steps 1 --warmup 1 --no-cpu"
$NCU --set full --import-source on -k regex:'gaussfit|nanodomain|decorr|objectives|synapse_cell|f1_match|apodize|spectrum_gather|pack_observation' -c 18 -f -o $OUT/${TAG}_env256 $ENV1 > $OUT/ncu6.log 2>&1
ls -la $OUT | tail -20
# summaries are extracted on the box (the reports together can exceed what gpurun merges back)
for n in gather scan env256; do
  python scripts/summarize_ncu.py $OUT/${TAG}_$n.ncu-rep $OUT/${TAG}_ncu_$n > /dev/null 2>> $OUT/ncu_sum.log
done
du -sh $OUT/*.ncu-rep
rm -f $OUT/*.ncu-rep
