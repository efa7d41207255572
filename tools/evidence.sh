# Evidence run of a round: `gpurun --timeout 1700 -- bash tools/evidence.sh r02x`
# Writes everything under gpurun_out/<tag>_*; profiles/ncu_summary.py turns the .ncu-rep files into
# the text summaries that are committed.  Bench numbers never come from a run under ncu.
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || { tail -5 $O/${TAG}_bench.err; exit 1; }
python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
# launch list of the bench command's main workload (cold-cache, serialised: shares, not absolutes)
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_launches.log 2>&1
# one full capture of the dominant kernel: textured instantiation on the headline workload (64 views keep the replays short),
# untextured on configs[1]
CMD="python bench.py --views 64 --steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
$CMD > $O/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:raster_tex_kernel -s 3 -c 1 -f -o $O/${TAG}_tex_raster $CMD > $O/${TAG}_tex_raster.log 2>&1
CMD="python bench.py --workload mesh1080 --steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
$CMD > $O/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -f -o $O/${TAG}_raster $CMD > $O/${TAG}_raster.log 2>&1
ls -la $O | grep ${TAG}
