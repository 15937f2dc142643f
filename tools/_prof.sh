python tools/profile_target.py pvar 2 > gpurun_out/plain_r1m_pvar.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"cross_pair|pair_reduce|fft_pass" -c 8 -s 8 -o gpurun_out/prof_r1m_pvar --force-overwrite python tools/profile_target.py pvar 2 > gpurun_out/ncu_r1m_pvar.log 2>&1
tail -2 gpurun_out/ncu_r1m_pvar.log
python tools/profile_target.py fwht 2 > gpurun_out/plain_r1m_fwht.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"wht_pass" -c 4 -s 4 -o gpurun_out/prof_r1m_fwht --force-overwrite python tools/profile_target.py fwht 2 > gpurun_out/ncu_r1m_fwht.log 2>&1
tail -2 gpurun_out/ncu_r1m_fwht.log
