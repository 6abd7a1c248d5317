mkdir -p gpurun_out
CMD="python bench.py --workload C4_plate3d_CG2 --steps 1 --warmup 1 --no-other-configs --no-cpu-baseline --no-parity"
for t in 1 0; do
SG_STENCIL_TILED=$t ncu --set full --clock-control none --import-source on -k "regex:k_stencil_apply" -s 20 -c 2 -o gpurun_out/r2y_stencil_tiled$t -f $CMD > gpurun_out/r2y_ncu$t.log 2>&1; echo "ncu $t rc=$?"
done
ls -la gpurun_out/*.ncu-rep
