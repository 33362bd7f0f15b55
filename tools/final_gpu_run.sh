#!/bin/bash
# One GPU call for the round's final records (run from the repo root under gpurun): reference-pin tests with their error
# prints, the full GPU suite, smoke, the DRAM-traffic capture stamped with the current source hash, the bench lines and
# the ncu launch list of the bench command.  Everything lands in gpurun_out/r2c_*.
T=r2c
python -m pytest tests/test_reference_pin.py -m gpu -q -s > gpurun_out/${T}_pin_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pin_tests.log
timeout 240 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_reference_pin.py > gpurun_out/${T}_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
timeout 240 ncu --profile-from-start off --set full --clock-control none -f -o /tmp/${T}_step python tools/step_traffic.py 256 > gpurun_out/${T}_step_ncu.log 2>&1 \
  && ncu -i /tmp/${T}_step.ncu-rep --page raw --csv > gpurun_out/${T}_step_raw.csv 2>/dev/null \
  && python tools/traffic_from_ncu.py gpurun_out/${T}_step_raw.csv 256 profiles/r2_step_traffic.json && cp profiles/r2_step_traffic.json gpurun_out/${T}_step_traffic.json
timeout 300 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?" >> gpurun_out/${T}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
tail -4 gpurun_out/${T}_pin_tests.log; tail -3 gpurun_out/${T}_gpu_tests.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-300 gpurun_out/${T}_bench.json
