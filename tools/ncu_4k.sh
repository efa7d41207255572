# one ncu pass: raster_tex_kernel of configs[2] (textured mesh at 4K + 16 blended overlay quads per frame)
O=gpurun_out
CMD="python bench.py --workload mesh4k_tex --steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 1"
$CMD > $O/4k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:raster_tex_kernel -s 3 -c 1 -f -o $O/r02n_4k $CMD > $O/r02n_4k.log 2>&1
ls -la $O | grep r02n_4k
