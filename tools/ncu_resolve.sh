O=gpurun_out
CMD="python bench.py --views 64 --steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
$CMD > $O/res_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:resolve_kernel -s 3 -c 1 -f -o $O/r02j_resolve $CMD > $O/r02j_resolve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_vis_kernel -s 3 -c 1 -f -o $O/r02j_vis $CMD > $O/r02j_vis.log 2>&1
ls -la $O | grep r02j
