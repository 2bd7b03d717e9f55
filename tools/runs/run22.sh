# ncu --set full captures summarised ON the box (the raw reports exceed the 64 MiB copy-back limit); only text + csv come back
mkdir -p /tmp/rep gpurun_out
cap() {  # name, kernel regex, count, skip, command...
  name=$1; rx=$2; cnt=$3; skip=$4; shift 4
  timeout 600 ncu --set full --clock-control none -k regex:"$rx" -s $skip -c $cnt -o /tmp/rep/r02_full_$name -f "$@" > /tmp/rep/ncu_$name.log 2>&1
  echo "# $name: ncu --set full --clock-control none -k regex:$rx -s $skip -c $cnt $*" >> gpurun_out/r02_kernels_ncu_full.txt
  python tools/ncu_full_summary.py /tmp/rep/r02_full_$name.ncu-rep >> gpurun_out/r02_kernels_ncu_full.txt 2>&1
}
rm -f gpurun_out/r02_kernels_ncu_full.txt
timeout 120 python tools/unet_step.py 1 2 > /dev/null 2>&1
cap gn_group 'gn_nhwc_group_kernel' 3 8 python tools/unet_step.py 1 2
cap splitk 'splitk_reduce_kernel' 2 4 python tools/unet_step.py 1 2
cap attn_b2 'attention_kernel' 8 0 python tools/unet_step.py 1 2
cap gemm_b2 'gemm_tcgen05_kernel' 10 20 python tools/unet_step.py 1 2
cap gn_c1 'gn_nchw_cluster_kernel' 1 2 python tools/gn_c1.py
cap sampler 'cfg_dpm_step_kernel|randn_kernel|timestep_sinusoid' 3 0 python -c "import __graft_entry__ as g; g.smoke()"
cap ln_b32 'layer_norm_kernel' 2 2 python tools/unet_step.py 1 32
HOT_ONCE=1 cap geglu_b8 'gemm_tcgen05_kernel' 1 0 python tools/hot_kernels.py geglu 8
HOT_ONCE=1 cap attn_b8 'attention_kernel' 1 0 python tools/hot_kernels.py attn 8
cap vae 'gn_nhwc|gemm_tcgen05_kernel|vae_post|softmax' 6 40 python tools/vae_step.py 1 8
wc -l gpurun_out/r02_kernels_ncu_full.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_tcgen05|gn_|layer_norm|attention|splitk|cast_|concat|im2col|upsample|silu|vae_post|softmax|latent" -c 1200 --csv --log-file gpurun_out/r02_vae_b8_launches.csv python tools/vae_step.py 2 8 > /tmp/rep/ncu_vae_list.log 2>&1; tail -1 /tmp/rep/ncu_vae_list.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_tcgen05|gn_|layer_norm|attention|splitk|cast_|concat|im2col|upsample|silu" -c 1100 --csv --log-file gpurun_out/r02_unet_step_b32_launches.csv python tools/unet_step.py 2 32 > /tmp/rep/ncu_b32_list.log 2>&1; tail -1 /tmp/rep/ncu_b32_list.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_tcgen05|gn_|layer_norm|attention|splitk|cast_|concat|im2col|upsample|silu" -c 700 --csv --log-file gpurun_out/r02_unet_step_b2_launches_c.csv python tools/unet_step.py 2 2 > /tmp/rep/ncu_b2_list.log 2>&1; tail -1 /tmp/rep/ncu_b2_list.log
for mb in 48 1000000; do echo "VAE C4 SDOD_GN_GROUP_MAXMB=$mb"; SDOD_GN_GROUP_MAXMB=$mb timeout 200 python - <<'PY'
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
ls -la gpurun_out | tail -8; du -sh gpurun_out
