# Evidence run of a round: `gpurun --timeout 1500 -- bash tools/evidence.sh r01q`
# Writes everything under gpurun_out/<tag>_*; profiles/ncu_summary.py turns the .ncu-rep files into
# the text summaries that are committed.  Bench numbers never come from a run under ncu.
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python bench.py --steps 200 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
for w in mesh4k_tex views1080_tex fill4k; do
	python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > $O/${TAG}_$w.json 2> $O/${TAG}_$w.err
done
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
	python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
# one full capture of the dominant kernel, untextured and textured instantiation
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -f -o $O/${TAG}_raster \
	python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $O/${TAG}_raster.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_tex_kernel -s 3 -c 1 -f -o $O/${TAG}_tex_raster \
	python bench.py --workload mesh4k_tex --steps 3 --warmup 2 --no-cpu-baseline > $O/${TAG}_tex_raster.log 2>&1
ls -la $O | grep ${TAG}
