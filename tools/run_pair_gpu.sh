# pair kernel: GPU tests, parity against the 16-lane kernel, bench, launch list and ncu capture (run under gpurun)
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3_gpu_tests.log; tail -4 gpurun_out/r3_gpu_tests.log
python tools/pair_check.py > gpurun_out/r03_pair_check.json 2> gpurun_out/pair_check.err; echo pair_check rc=$?
python bench.py --steps 20 --warmup 5 --skip-fits > gpurun_out/r03_bench_quick.json 2> gpurun_out/r03_bench_quick.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_launches.csv python bench.py --steps 5 --warmup 3 --skip-cpu --skip-fits > gpurun_out/ncu_l.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:misti_jsfs_pair_kernel -s 3 -c 1 -o gpurun_out/r03_pair -f python bench.py --steps 3 --warmup 3 --skip-cpu --skip-fits > gpurun_out/ncu_pair.log 2>&1; echo ncu rc=$?
