export SDOD_STREAMK=0
for hw in 8 16 32 64; do timeout 200 python tools/step_time.py 2 hw$hw $hw 2>&1 | sed -n 1,3p; done
SDOD_PDL=0 timeout 200 python tools/step_time.py 2 nopdl 2>&1 | sed -n 1,3p
SDOD_SPLITK_CLUSTER=0 timeout 200 python tools/step_time.py 2 nocl 2>&1 | sed -n 1,3p
SDOD_SPLITK_CLUSTER=0 timeout 200 python tools/step_time.py 1 b1_nocl 2>&1 | sed -n 1,3p
SDOD_SPLITK_CLUSTER=0 timeout 200 python tools/step_time.py 4 b4_nocl 2>&1 | sed -n 1,3p
