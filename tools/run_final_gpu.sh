# final evidence of the round (run under gpurun): tests, smoke, bench lines, launch list, captures of the new kernels
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r03_gpu_tests.log; cat gpurun_out/r03_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03_smoke.log 2>&1; tail -2 gpurun_out/r03_smoke.log
python bench.py > gpurun_out/r03_bench_1gpu.json 2> gpurun_out/r03_bench_1gpu.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03_bench_reference_arm.json 2>/dev/null; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_launches.csv python bench.py --steps 5 --warmup 3 --skip-cpu --skip-fits > gpurun_out/ncu_l.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:misti_jsfs_pair_kernel -s 3 -c 1 -o gpurun_out/r03_pair -f python bench.py --steps 3 --warmup 3 --skip-cpu --skip-fits > gpurun_out/ncu_pair.log 2>&1; echo ncu pair rc=$?
ncu --set full --clock-control none --import-source on -k regex:misti_post_split_quad_kernel -s 3 -c 1 -o gpurun_out/r03_post_quad -f python bench.py --steps 3 --warmup 3 --skip-cpu --skip-fits > gpurun_out/ncu_pq.log 2>&1; echo ncu quad rc=$?
