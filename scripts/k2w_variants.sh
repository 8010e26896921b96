#!/bin/bash
# Builds libcmc_b200 with build-time variants of the K2w jackknife path into build_variants/ (here, no GPU needed);
# on the GPU box: for f in build_variants/*.so; do CMC_B200_LIB=$PWD/$f python scripts/time_msc_windows.py; done
set -e
cd "$(dirname "$0")/.."
python -m multimodal_biosignal_analysis_b200.build > /dev/null
mkdir -p build_variants
OBJS=$(ls multimodal_biosignal_analysis_b200/csrc/build/*.o | grep -v msc_windows.o)
for v in "1 0 2" "1 1 3" "1 1 4" "0 1 3" "0 0 3" "1 0 3" "0 0 2" "0 1 4"; do   # RCP_MUFU Y_SMEM MINB
  set -- $v
  tag="rcp$1_ysm$2_minb$3"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I include \
       -DCMC_K2W_RCP_MUFU=$1 -DCMC_K2W_Y_SMEM=$2 "-DCMC_K2W_MINB(K)=((K)<=8?$3:1)" -Xptxas -v \
       -c multimodal_biosignal_analysis_b200/csrc/msc_windows.cu -o build_variants/msc_$tag.o 2> build_variants/ptxas_$tag.log
  nvcc -shared -o build_variants/libcmc_$tag.so $OBJS build_variants/msc_$tag.o \
       -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --cudart static
  rm build_variants/msc_$tag.o
  echo "$tag: $(grep -A2 'msc_windows_kernelILi5ELb1' build_variants/ptxas_$tag.log | grep -oE '[0-9]+ bytes spill stores|Used [0-9]+ registers' | tr '\n' ' ')"
done
