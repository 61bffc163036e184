set -x
export REPS=2
W="stiff32 stiff64"
python profiles/prof_r2.py $W > gpurun_out/r2_prof2_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'racu_jit' -o /tmp/r02_stiff python profiles/prof_r2.py $W > gpurun_out/r2_prof2_ncu.log 2>&1
tail -3 gpurun_out/r2_prof2_ncu.log
ncu -i /tmp/r02_stiff.ncu-rep --page raw --csv > gpurun_out/r02_stiff_raw.csv
ncu -i /tmp/r02_stiff.ncu-rep --page details --csv > gpurun_out/r02_stiff_details.csv
ls -la gpurun_out
