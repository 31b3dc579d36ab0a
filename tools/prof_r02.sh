#!/bin/bash
# One GPU call: parity tests ($2 = "full" for the whole GPU suite), a bench line, the ncu launch list of one bench run and a
# --set full capture of the polish kernels.
set -x
if [ "$2" = "full" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/$1_tests.log 2>&1
else python -m pytest tests/test_gpu_polish.py tests/test_gpu_fullsize.py::test_config2_full_size -x -q > gpurun_out/$1_tests.log 2>&1; fi
tail -3 gpurun_out/$1_tests.log
python bench.py --steps 10 --warmup 3 --no-t3 --no-cpu-baseline > gpurun_out/$1_bench.json 2> gpurun_out/$1_bench.err; tail -c 600 gpurun_out/$1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$1_launches.csv python bench.py --steps 1 --warmup 3 --no-t3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_tile|k_goodk|k_compact' --launch-skip 6 --launch-count 3 -o gpurun_out/$1_full python bench.py --steps 1 --warmup 3 --no-t3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/ | tail -5
