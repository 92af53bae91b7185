timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_final_a.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -s 16 -c 4 -f -o gpurun_out/r02_step_final2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_final_b.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/r02_step_final2.ncu-rep gpurun_out/r02_launches_final.csv
