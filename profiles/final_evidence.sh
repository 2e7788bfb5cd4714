# the evidence run of the final build (one gpurun call): bench lines, ncu captures, launch list
set -x
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/g_steps20.json 2> gpurun_out/g_steps20.err
timeout 300 python bench.py > gpurun_out/g_default.json 2> gpurun_out/g_default.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g_reference.json 2> gpurun_out/g_reference.err
for c in c2 c3 c4 c5_1m; do timeout 300 python bench.py --config $c --steps 100 --warmup 5 --repeats 3 > gpurun_out/g_$c.json 2> gpurun_out/g_$c.err; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 400 -c 1 -f -o gpurun_out/step_r2_final python profiles/step_time.py 65536 > gpurun_out/g_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 1 -f -o gpurun_out/step_r2_final_1m python profiles/step_time.py 1048576 > gpurun_out/g_ncu2.log 2>&1
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 33100 --launch-count 400 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 20 --warmup 3 --graph off --repeats 10 --setup-steps 2048 > gpurun_out/g_ncu3.log 2>&1
timeout 120 python profiles/step_time.py 65536 1048576 > gpurun_out/g_step_time.txt 2>&1
ls -la gpurun_out | tail -20
