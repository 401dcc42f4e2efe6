python tools/profile_step.py --config 2 --n 1000000 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lidf -f -o gpurun_out/r02_prof_lidf2 python tools/profile_step.py --config 2 --n 1000000 > gpurun_out/r02_ncu_full_lidf2.log 2>&1
tail -2 gpurun_out/r02_ncu_full_lidf2.log
ncu -i gpurun_out/r02_prof_lidf2.ncu-rep --page raw --csv > gpurun_out/r02_raw_lidf2.csv
python tools/ncu_summary.py gpurun_out/r02_raw_lidf2.csv
