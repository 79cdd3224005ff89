for v in libml4ca_b200 libvar_sh64g4; do echo "== $v"; ML4CA_LIB=$PWD/ml4ca_b200/$v.so HIDDEN=64,64 ENVS=16384 timeout 200 python tools/ppo_update_time.py; done > gpurun_out/ppo_time_s4z.log 2>&1
ML4CA_LIB=$PWD/ml4ca_b200/libvar_sh64g4.so timeout 500 python -m pytest tests/test_ppo_update_gpu.py tests/test_trpo_gpu.py -x -q > gpurun_out/ppo_test_s4z.log 2>&1; echo rc=$?
