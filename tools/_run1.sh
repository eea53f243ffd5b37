mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r1_gpu.txt
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
SWEEP_VARIANTS="wave=5,allclear=0:wave=5,allclear=1:wave=5,allclear=1,wctas=3:wave=5,allclear=1,margin=6:wave=5,allclear=1,wrefill=4:wave=5,allclear=1,wrefill=16:wave=5,allclear=1,check=32:wave=5,allclear=1,check=8:wave=6,allclear=1:wave=4,allclear=1:wave=5,allclear=1,tail=50:wave=-1,allclear=1" timeout 300 python tools/sweep.py c4 > gpurun_out/r1_c4_sweep.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/r1_bench.log 2> gpurun_out/r1_bench.err
METRICS=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_summary as n; print(','.join(n.KEYS))")
timeout 400 ncu --metrics $METRICS --clock-control none -k regex:wave -c 1 -o gpurun_out/r1_c4_wave_allclear python tools/profile_c4.py wave=5 > gpurun_out/r1_ncu.log 2>&1
SWEEP_C4_RAYS=1048576 SWEEP_VARIANTS="wave=5,allclear=0:wave=5,allclear=1:wave=5,allclear=1,wctas=3:wave=6,allclear=1:wave=-1" timeout 200 python tools/sweep.py c4 > gpurun_out/r1_c4_small.log 2>&1
tail -3 gpurun_out/r1_pytest.log; cat gpurun_out/r1_c4_sweep.log | cut -c1-200; tail -c 600 gpurun_out/r1_bench.err
