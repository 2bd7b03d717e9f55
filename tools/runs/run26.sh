timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "upsample or conv" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -s 2>&1 | grep -E "passed|failed|rel|PSNR|dB" | tail -12
for v in 1 0; do echo "SDOD_CONV_UP2=$v"; SDOD_CONV_UP2=$v timeout 300 python tools/step_time.py 32 up2_$v 2>&1 | sed -n 2,2p; SDOD_CONV_UP2=$v timeout 300 python tools/step_time.py 2 up2b2_$v 2>&1 | sed -n 2,2p; done
grep -E "conv3up2|upsample|conv3 HW4096 Cin640 Cout640|conv3 HW1024 Cin1280 Cout1280|conv3 HW256 Cin1280 Cout1280 " gpurun_out/step_time_up2_1.txt gpurun_out/step_time_up2_0.txt | head -12
for v in 1 0; do echo "VAE SDOD_CONV_UP2=$v"; SDOD_CONV_UP2=$v timeout 200 python - <<'PY'
import sys,os,torch
sys.path.insert(0,'stable-diffusion-on-device_b200')
from sdod import model as M
vae=M.VaeDecoder(None,seed=1,latent_hw=64,max_batch=8); z=torch.randn(8,4,64,64,device='cuda')
for _ in range(3): vae(z)
ts=[]
for _ in range(7):
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); vae(z); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("vae decode batch 8: median %.2f ms  min %.2f"%(sorted(ts)[3],min(ts)))
PY
done
