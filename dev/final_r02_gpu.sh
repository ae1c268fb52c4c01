#!/bin/bash
# Round-2 closing call on one B200: GPU tests, smoke(), and the r02-named evidence of the headline step
# (launch list of the bench command + one --set full capture of covproj_tma_kernel).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_close_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_close_gputest.log )
tail -3 gpurun_out/r02_close_gputest.log
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_close_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_close_smoke.log )
tail -2 gpurun_out/r02_close_smoke.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02_headline.csv \
    python bench.py --steps 2 --warmup 3 --no-other --no-splat > gpurun_out/r02_close_ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
PROF_ONLY=covproj timeout 150 ncu --set full --clock-control none --import-source on -k regex:covproj_tma -c 2 -f \
    -o gpurun_out/prof_r02_covproj python profiles/prof_driver.py > gpurun_out/r02_close_ncu_full.log 2>&1
echo "ncu set full rc=$?"
ls -la gpurun_out | tail -8
