mkdir -p gpurun_out
for n in 2097152 4194304; do
  VRT_WAVE_DENSE_RAYS_PER_BRICK=0 SWEEP_C4_RAYS=$n SWEEP_VARIANTS="wave=5,allclear=0:wave=5,allclear=1,wctas=3:wave=5,allclear=1:wave=5,allclear=1,wctas=3:wave=5,allclear=1" timeout 200 python tools/sweep.py c4 > gpurun_out/r2_c4_$n.log 2>&1
done
( time timeout 600 python bench.py ) > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --extras none --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
cat gpurun_out/r2_c4_*.log | cut -c1-200; tail -c 300 gpurun_out/r2_bench.err; wc -l gpurun_out/r2_launches.csv
