# Evidence runs of a round, ONE ncu invocation per gpurun call:
#   gpurun --timeout 1700 -- bash tools/evidence.sh r02x bench     bench + reference arm + launch list (the ncu pass)
#   gpurun --timeout 900  -- bash tools/evidence.sh r02x tex       one `ncu --set full` pass: the raster stage's kernels, 64 textured 1080p views
#   gpurun --timeout 900  -- bash tools/evidence.sh r02x mesh      ... configs[1]
# Writes everything under gpurun_out/<tag>_*; profiles/ncu_summary.py turns the .ncu-rep files into
# the text summaries that are committed.  Bench numbers never come from a run under ncu.
TAG=${1:-rXX}
MODE=${2:-bench}
O=gpurun_out
mkdir -p $O
# the raster stage is raster_opaque_kernel<true> for these workloads (every primitive an opaque triangle; with
# DTR_B200_FUSED=0 it is raster_opaque_kernel<false> + resolve_kernel: COUNT=2 captures both)
STAGE='regex:raster_opaque_kernel|resolve_kernel|raster_kernel|raster_tex_kernel'
COUNT=${COUNT:-1}
case $MODE in
bench)
  python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || { tail -5 $O/${TAG}_bench.err; exit 1; }
  python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
  # launch list of the bench command's main workload (cold-cache, serialised: shares, not absolutes)
  CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
  $CMD > $O/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_launches.log 2>&1
  ;;
tex)
  CMD="python bench.py --views 64 --steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
  $CMD > $O/${TAG}_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "$STAGE" -s 6 -c $COUNT -f -o $O/${TAG}_tex_stage $CMD > $O/${TAG}_tex_stage.log 2>&1
  ;;
mesh)
  CMD="python bench.py --workload mesh1080 --steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
  $CMD > $O/${TAG}_plain3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "$STAGE" -s 6 -c $COUNT -f -o $O/${TAG}_mesh_stage $CMD > $O/${TAG}_mesh_stage.log 2>&1
  ;;
esac
ls -la $O | grep ${TAG}
