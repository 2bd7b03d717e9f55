SDOD_LN_FUSE=0 timeout 300 python tools/step_time.py 2 lnf0 2>&1 | sed -n 2,3p
timeout 300 python tools/step_time.py 2 lnf1 2>&1 | sed -n 2,3p
SDOD_LN_FUSE_MINCTAS=40 timeout 300 python tools/step_time.py 2 lnf40 2>&1 | sed -n 2,3p
SDOD_LN_FUSE_MINCTAS=100 timeout 300 python tools/step_time.py 2 lnf100 2>&1 | sed -n 2,3p
