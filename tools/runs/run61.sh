echo "default pairs + PDL on pairs"; SDOD_PAIR_PDL=1 timeout 90 python tools/step_time.py 2 pq1 2>&1 | sed -n 2,2p
echo "conv pairs at any size + PDL"; SDOD_GEMM_PAIR=3 SDOD_PAIR_PDL=1 timeout 90 python tools/step_time.py 2 pq3 2>&1 | sed -n 2,2p
echo "conv pairs at any size, no PDL"; SDOD_GEMM_PAIR=3 timeout 90 python tools/step_time.py 2 pq3n 2>&1 | sed -n 2,2p
echo "B32 default pairs + PDL"; SDOD_PAIR_PDL=1 timeout 120 python tools/step_time.py 32 pq1b32 2>&1 | sed -n 2,2p
echo "B32 default"; timeout 120 python tools/step_time.py 32 pq0b32 2>&1 | sed -n 2,2p
grep -E " conv3" gpurun_out/step_time_pq3.txt | head -12
