# ncu --set full (with source) of the cloth kernels at fold_cloth1_para size (128 envs x 512 nodes, 50 substeps)
set -x
CMD="python profiles/other_configs.py cloth_para"
$CMD > gpurun_out/cloth_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_cloth_bwd" -s 5 -c 1 -o gpurun_out/prof_cloth_bwd $CMD > gpurun_out/ncu_cloth_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_cloth_fwd" -s 5 -c 1 -o gpurun_out/prof_cloth_fwd $CMD > gpurun_out/ncu_cloth_fwd.log 2>&1
ls -la gpurun_out/prof_cloth*
