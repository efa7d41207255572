# one ncu pass: the opaque stage on the fill-rate workload (after a plain run; DTR_B200_FUSED=0 and -c 2 for the two-kernel form)
O=gpurun_out
CMD="python bench.py --workload fill4k --steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 1"
$CMD > $O/fill_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"raster_opaque_kernel|resolve_kernel" -s 6 -c 2 -f -o $O/r02j_fill $CMD > $O/r02j_fill.log 2>&1
tail -3 $O/r02j_fill.log; ls -la $O | grep r02j_fill
