echo "pairs forced, no PDL on pairs"; SDOD_GEMM_PAIR=2 timeout 120 python tools/step_time.py 2 pp0 2>&1 | sed -n 2,2p
echo "pairs forced, PDL on pairs"; SDOD_GEMM_PAIR=2 SDOD_PAIR_PDL=1 timeout 90 python tools/step_time.py 2 pp1 2>&1 | sed -n 2,2p; echo rc=$?
nvidia-smi --query-gpu=name,temperature.gpu --format=csv,noheader
