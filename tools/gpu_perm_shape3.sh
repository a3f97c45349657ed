#!/bin/bash
mkdir -p gpurun_out
PERM_SIZES=28,32,36,40 python tools/gpu_perm_shape.py tools/variants/libwb_perm_base.so tools/variants/libwb_perm_nc2.so tools/variants/libwb_perm_nc8.so tools/variants/libwb_perm_g4.so tools/variants/libwb_perm_nc8g4.so tools/variants/libwb_perm_tree.so 2>&1 | tee gpurun_out/perm_shape3.txt
