O=gpurun_out; TAG=r02d
CMD="python bench.py --workload mesh1080 --steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
$CMD > $O/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -f -o $O/${TAG}_raster $CMD > $O/${TAG}_raster.log 2>&1
CMD="python bench.py --views 64 --steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 2"
$CMD > $O/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:raster_tex_kernel -s 3 -c 1 -f -o $O/${TAG}_tex_raster $CMD > $O/${TAG}_tex_raster.log 2>&1
ls -la $O | grep ${TAG}
