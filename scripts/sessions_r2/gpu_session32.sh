# ncu --set full of the exchange variant of the streaming kernel and of the two compact projector kernels
mkdir -p gpurun_out
bash scripts/ncu_kernel.sh s32_dense_nlc "stream_dense_kernel" 6 CHEFSI_B200_NLC=1
bash scripts/ncu_kernel.sh s32_nloc_project_compact "nloc_kernel<20, 1, 0" 4 CHEFSI_B200_NLC=1
bash scripts/ncu_kernel.sh s32_nloc_expand_write "nloc_kernel<20, 1, 4" 4 CHEFSI_B200_NLC=1
for t in s32_dense_nlc s32_nloc_project_compact s32_nloc_expand_write; do cat gpurun_out/prof_$t.txt | head -40; done
