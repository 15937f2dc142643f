python -m pytest tests/test_kernels_gpu.py -x -q -k "two_pass" 2>&1 | grep -v Warning | cut -c1-400 | tail -8
python tools/profile_target.py mll 3 && ncu --set full --clock-control none --import-source on -k regex:mll_pass -c 3 -s 3 -o gpurun_out/prof_hs --force-overwrite python tools/profile_target.py mll 3 > gpurun_out/ncu_hs.log 2>&1
tail -3 gpurun_out/ncu_hs.log
