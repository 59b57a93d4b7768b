# rmsc03 run kernel (order tables searched 32 entries per pass): plain run first, then one ncu --set full capture of the second launch.
set -u
N=2368          # one whole wave: 148 SMs x 16 resident one-warp CTAs
CMD="python tools/quick_rmsc03_throughput.py $N"
$CMD > gpurun_out/rmsc03_plain.log 2> gpurun_out/rmsc03_plain.err || { echo "plain run failed"; tail -5 gpurun_out/rmsc03_plain.err; exit 1; }
cat gpurun_out/rmsc03_plain.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:abx_run_kernel -s 1 -c 1 -f -o gpurun_out/prof_rmsc03 $CMD > gpurun_out/ncu_rmsc03.log 2>&1
ls -la gpurun_out | tail -4
