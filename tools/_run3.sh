mkdir -p gpurun_out
( time timeout 500 python -m pytest tests -m gpu -x -q -k "wavefront or probe or checked" ) > gpurun_out/r3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3_pytest.log
SWEEP_VARIANTS="wave=5,reuse=0:wave=5,reuse=1:wave=5,reuse=2:wave=5,reuse=1,wctas=3:wave=5,reuse=1,margin=6:wave=5,reuse=1,wrefill=4:wave=5,reuse=1,check=32:wave=4,reuse=1:wave=5,reuse=0" timeout 300 python tools/sweep.py c4 > gpurun_out/r3_c4_sweep.log 2>&1
for n in 1048576 4194304; do
VRT_WAVE_DENSE_RAYS_PER_BRICK=0 SWEEP_C4_RAYS=$n SWEEP_VARIANTS="wave=5,reuse=0,wctas=3:wave=5,reuse=0:wave=5,reuse=1,wctas=3:wave=5,reuse=1:wave=5,reuse=2" timeout 200 python tools/sweep.py c4 > gpurun_out/r3_c4_$n.log 2>&1
done
tail -4 gpurun_out/r3_pytest.log; cat gpurun_out/r3_c4_sweep.log gpurun_out/r3_c4_1048576.log gpurun_out/r3_c4_4194304.log | cut -c1-220
