timeout 900 python -m pytest tests/test_gpu_libsdod.py -x -q -s -k "text or prompt" 2>&1 | tail -5
timeout 300 python tools/step_time.py 32 r2_b32_base 2>&1 | sed -n 1,14p
SDOD_GN_GROUP_MAXMB=4096 timeout 300 python tools/step_time.py 32 r2_b32_gng 2>&1 | sed -n 1,16p
SDOD_LN_FUSE_MAXCTAS=100000 timeout 300 python tools/step_time.py 32 r2_b32_ln 2>&1 | sed -n 1,16p
SDOD_LN_FUSE_MAXCTAS=100000 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "unet" 2>&1 | tail -3
