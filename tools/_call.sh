mkdir -p gpurun_out
# launch list of one frame (duration only, one pass per kernel)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r3e_launches.csv python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/r3e_launches.log 2>&1
tail -2 gpurun_out/r3e_launches.log | cut -c1-200
# full capture: the first bench-size sdf-only launch of the frame and one shading launch
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_mlp_tc<\(int\)0>' -s 2 -c 1 -o gpurun_out/r3e_tc0_bench python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/r3e_ncu0.log 2>&1
tail -2 gpurun_out/r3e_ncu0.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -c 2 -o gpurun_out/r3e_tc python tools/prof_kernels.py > gpurun_out/r3e_ncu1.log 2>&1
tail -2 gpurun_out/r3e_ncu1.log | cut -c1-200
ls -la gpurun_out/r3e*
