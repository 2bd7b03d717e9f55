# evidence pass: GEMM phase timeline, GN C1, ncu launch list of the batch-2 step (time + DRAM bytes), compute-sanitizer over op-level tests
timeout 300 python tools/gemm_timeline.py r2b 2>&1 | tail -18
timeout 200 python tools/gn_c1.py > gpurun_out/r02_gn_c1_a.json 2>&1; tail -c 900 gpurun_out/r02_gn_c1_a.json
timeout 120 python tools/unet_step.py 2 2 > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_tcgen05|gn_|layer_norm|attention|splitk|cast_|concat|im2col|upsample|silu" -c 700 --csv --log-file gpurun_out/r02_unet_step_b2_launches_b.csv python tools/unet_step.py 2 2 > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
SEL='test_gemm_every_tile_width or test_gemm_epilogues or test_gemm_persistent_scheduler_epilogues or test_conv3x3_cta_pair_epilogues or test_gemm_split_k or test_gemm_fused_layer_norm_epilogue or test_fused_attention or test_gn_config1 or test_gn_single_launch or test_cfg_dpm_sampler or test_layer_norm'
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r02_sanitizer_$tool.log python -m pytest tests/test_gpu_ops.py -x -q -k "$SEL" > gpurun_out/r02_sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r02_sanitizer_${tool}_pytest.log; grep -c "=========" gpurun_out/r02_sanitizer_$tool.log; tail -4 gpurun_out/r02_sanitizer_$tool.log
done
