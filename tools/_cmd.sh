timeout 500 python -m pytest tests/test_ppo_update_gpu.py tests/test_trpo_gpu.py -x -q > gpurun_out/ppo_test_s4y.log 2>&1; echo rc=$?
HIDDEN=80,80,80 ENVS=16384,4 timeout 200 python tools/ppo_update_time.py > gpurun_out/ppo_time_s4y.log 2>&1
