# round 2, session 2, call 6: one ncu --set full capture (with source counters) of the two attention kernels at the cfg-2 layer shape
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'qk_kernel|sv_kernel' -s 4 -c 2 -f -o gpurun_out/r02s2_attn python tools/profile_fused.py > gpurun_out/r02s2_ncu.log 2>&1; tail -2 gpurun_out/r02s2_ncu.log
ls -la gpurun_out/*.ncu-rep
