python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/f_tests.log
python bench.py > gpurun_out/f_cfg3.json 2> gpurun_out/f_cfg3.err
python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/f_cfg5.json 2> gpurun_out/f_cfg5.err
python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/f_cfg2.json 2> gpurun_out/f_cfg2.err
python tools/refit_sweep.py > gpurun_out/f_sweep.md 2> gpurun_out/f_sweep.err
cat gpurun_out/f_tests.log; tail -c 600 gpurun_out/f_cfg3.json; tail -5 gpurun_out/f_sweep.md
