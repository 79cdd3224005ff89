timeout 400 python -m pytest tests/test_ppo_update_gpu.py -x -q > gpurun_out/ppo_test_s4l.log 2>&1; echo rc=$?
HIDDEN=80,80,80 ENVS=4096,4 timeout 200 python tools/ppo_update_time.py > gpurun_out/ppo_time_s4l.log 2>&1
HIDDEN=64,64,64 ENVS=4096 timeout 200 python tools/ppo_update_time.py >> gpurun_out/ppo_time_s4l.log 2>&1
HIDDEN=64,64 ENVS=4096 timeout 200 python tools/ppo_update_time.py >> gpurun_out/ppo_time_s4l.log 2>&1
