#!/bin/bash
# round 2, GPU call 30: DRAM bytes of one whole 1024^2 VAE decode (batch 1 and 4; norm inside the convs and as a pass),
# kernel by kernel, against the 13.46 GB / image of SURVEY.md §8(d)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for B in 1 4; do for mode in 1 0; do
  DK_CUDA_GRAPHS=0 DK_PROFILE_RANGE=1 DK_VAE_NORM_IN_CONV=$mode timeout 300 ncu --profile-from-start off \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r02_vae_dram_B${B}_norm${mode}.csv python tools/profile_vae.py $B 2 > gpurun_out/ncu_vae.log 2>&1
  echo "== B=$B norm_in_conv=$mode"
  python tools/sum_dram.py gpurun_out/r02_vae_dram_B${B}_norm${mode}.csv $(python -c "print(13.46*$B)") | tee gpurun_out/r02_vae_dram_B${B}_norm${mode}.txt | tail -14
  gzip -f gpurun_out/r02_vae_dram_B${B}_norm${mode}.csv
done; done
