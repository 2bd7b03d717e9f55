# ncu --set full captures of the kernels that had no committed evidence, one or two launches each
cap() {  # name, kernel regex, count, skip, command...
  name=$1; rx=$2; cnt=$3; skip=$4; shift 4
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/r02_full_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$? $(ls -la gpurun_out/r02_full_$name.ncu-rep 2>/dev/null | awk '{print $5}')"
}
timeout 120 python tools/unet_step.py 1 2 > /dev/null 2>&1
cap gn_group 'gn_nhwc_group_kernel' 3 8 python tools/unet_step.py 1 2
cap splitk 'splitk_reduce_kernel' 2 4 python tools/unet_step.py 1 2
cap attn_b2 'attention_kernel' 8 0 python tools/unet_step.py 1 2
cap gemm_b2 'gemm_tcgen05_kernel' 12 20 python tools/unet_step.py 1 2
cap gn_c1 'gn_nchw_cluster_kernel' 1 2 python tools/gn_c1.py
cap sampler 'cfg_dpm_step_kernel|randn_kernel|timestep_sinusoid' 3 0 python -c "import __graft_entry__ as g; g.smoke()"
cap ln_b32 'layer_norm_kernel' 2 2 python tools/unet_step.py 1 32
HOT_ONCE=1 cap geglu_b8 'gemm_tcgen05_kernel' 1 0 python tools/hot_kernels.py geglu 8
cap vae 'gn_nhwc|gemm_tcgen05_kernel|vae_post|softmax' 6 40 python tools/vae_step.py 1 8
# launch lists with DRAM bytes: batch-32 UNet pass and the VAE decoder (batch 8)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_tcgen05|gn_|layer_norm|attention|splitk|cast_|concat|im2col|upsample|silu|vae_post|softmax|latent" -c 1200 --csv --log-file gpurun_out/r02_vae_b8_launches.csv python tools/vae_step.py 2 8 > gpurun_out/ncu_vae_list.log 2>&1; tail -1 gpurun_out/ncu_vae_list.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_tcgen05|gn_|layer_norm|attention|splitk|cast_|concat|im2col|upsample|silu" -c 1100 --csv --log-file gpurun_out/r02_unet_step_b32_launches.csv python tools/unet_step.py 2 32 > gpurun_out/ncu_b32_list.log 2>&1; tail -1 gpurun_out/ncu_b32_list.log
timeout 300 python tools/config_sweep.py > gpurun_out/r02_config_sweep.json 2>&1; tail -c 1500 gpurun_out/r02_config_sweep.json
