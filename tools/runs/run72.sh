mkdir -p /tmp/rep
timeout 120 python tools/unet_step.py 1 32 > /dev/null 2>&1
cap() { name=$1; rx=$2; skip=$3; cnt=$4; shift 4; timeout 400 ncu --set full --clock-control none -k regex:"$rx" -s $skip -c $cnt -o /tmp/rep/$name -f "$@" > /tmp/rep/$name.log 2>&1; echo "$name rc=$?"; }
cap attn40 'attention_kernel' 0 2 python tools/unet_step.py 1 32
cap attn80 'attention_kernel' 4 2 python tools/unet_step.py 1 32
cap gemm_b32 'gemm_tcgen05' 6 12 python tools/unet_step.py 1 32
cap gn_group 'gn_nhwc_group' 1 2 python tools/unet_step.py 1 32
cap gemm_b2 'gemm_tcgen05' 6 8 python tools/unet_step.py 1 2
python tools/ncu_full_summary.py /tmp/rep/attn40.ncu-rep /tmp/rep/attn80.ncu-rep /tmp/rep/gn_group.ncu-rep /tmp/rep/gemm_b32.ncu-rep /tmp/rep/gemm_b2.ncu-rep >> gpurun_out/r02_kernels_ncu_full_final.txt 2>&1
wc -l gpurun_out/r02_kernels_ncu_full_final.txt; cut -c1-330 gpurun_out/r02_kernels_ncu_full_final.txt | tail -80
