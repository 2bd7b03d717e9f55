mkdir -p /tmp/rep
timeout 120 python tools/unet_step.py 1 32 > /dev/null 2>&1
cap() { name=$1; rx=$2; skip=$3; cnt=$4; shift 4; timeout 400 ncu --set full --clock-control none -k regex:"$rx" -s $skip -c $cnt -o /tmp/rep/$name -f "$@" > /tmp/rep/$name.log 2>&1; echo "$name rc=$?"; }
cap ln_tma 'layer_norm_tma' 0 1 python tools/unet_step.py 1 32
cap attn40 'attention_kernel<40' 0 2 python tools/unet_step.py 1 32
cap attn80 'attention_kernel<80' 0 2 python tools/unet_step.py 1 32
cap gemm_p160 'gemm_tcgen05_kernel<160, 1, 0, 1>' 0 4 python tools/unet_step.py 1 32
cap gn_group 'gn_nhwc_group_kernel<float' 1 1 python tools/unet_step.py 1 32
cap gn_c1 'gn_nchw_cluster' 3 1 python tools/gn_c1.py
cap gemm_ln_b2 'gemm_tcgen05_kernel<160, 1, 0, 0>' 5 3 python tools/unet_step.py 1 2
python tools/ncu_full_summary.py /tmp/rep/ln_tma.ncu-rep /tmp/rep/attn40.ncu-rep /tmp/rep/attn80.ncu-rep /tmp/rep/gemm_p160.ncu-rep /tmp/rep/gn_group.ncu-rep /tmp/rep/gn_c1.ncu-rep /tmp/rep/gemm_ln_b2.ncu-rep > gpurun_out/r02_kernels_ncu_full_final.txt 2>&1
wc -l gpurun_out/r02_kernels_ncu_full_final.txt; cut -c1-260 gpurun_out/r02_kernels_ncu_full_final.txt | head -60
