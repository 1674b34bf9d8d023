# ncu --set full of the benched job's own k_step and k_net2 launches around tick 2000 (mid-job), then stop the job
O=gpurun_out
python tools/tick_sims.py 2000,2001,2002 > $O/r02_tick_sims.json 2> $O/r02_tick_sims.err
ncu --set full --clock-control none --import-source on -k regex:'k_step|k_net2' -s 4000 -c 6 --kill 1 -f -o $O/r02_final_bench \
  python bench.py --steps 1 --warmup 0 --host-loop python --no-ablation --no-cpu-baseline > $O/ncu_final.log 2>&1
ncu -i $O/r02_final_bench.ncu-rep --page raw --csv > $O/r02_final_bench_ncu_full.csv 2>> $O/ncu_final.log
tail -5 $O/ncu_final.log; cat $O/r02_tick_sims.json; wc -c $O/r02_final_bench_ncu_full.csv
