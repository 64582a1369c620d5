set -x
P="python tools/profile_run.py"
NCU="ncu --set full --clock-control none --import-source on -f"
# car kernels in steady state (300 steps in); render launches: 1 at reset + 3 per step (main, deferred, auto-reset)
$NCU -k regex:car_render_kernel -s 901 -c 2 -o gpurun_out/r02f_car_render_stack $P --what car --double 1 --steps 302 > /dev/null 2>&1
$NCU -k regex:car_render_kernel -s 901 -c 1 -o gpurun_out/r02f_car_render_ring $P --what car --double 1 --steps 302 --stack-mode ring > /dev/null 2>&1
$NCU -k regex:car_stack_shift_kernel -s 300 -c 1 -o gpurun_out/r02f_car_shift $P --what car --double 1 --steps 302 --stack-mode stack-shift > /dev/null 2>&1
$NCU -k regex:car_frame_aux_kernel -s 901 -c 1 -o gpurun_out/r02f_car_aux $P --what car --double 1 --steps 302 > /dev/null 2>&1
$NCU -k regex:car_sensor_kernel -s 300 -c 1 -o gpurun_out/r02f_car_sensor $P --what car --double 1 --steps 302 > /dev/null 2>&1
$NCU -k regex:car_collide_kernel -s 300 -c 1 -o gpurun_out/r02f_car_collide $P --what car --double 1 --steps 302 > /dev/null 2>&1
$NCU -k regex:car_step_kernel -s 600 -c 2 -o gpurun_out/r02f_car_step $P --what car --double 1 --steps 302 > /dev/null 2>&1
$NCU -k regex:car_step_kernel -s 300 -c 1 -o gpurun_out/r02f_car_step_single $P --what car --double 0 --steps 302 > /dev/null 2>&1
$NCU -k regex:pong_raster_quad_kernel -s 20 -c 1 -o gpurun_out/r02f_quad42 $P --what pong --dim 42 --steps 25 > /dev/null 2>&1
# launch lists
ncu --metrics gpu__time_duration.sum --clock-control none -s 4800 -c 40 --csv --log-file gpurun_out/r02f_launches_car_double.csv $P --what car --double 1 --steps 305 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 40 --csv --log-file gpurun_out/r02f_launches_car_single.csv $P --what car --double 0 --steps 305 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --car-steps 3 > gpurun_out/r02f_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -15
