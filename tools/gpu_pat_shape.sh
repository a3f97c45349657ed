#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_pat_shape.py tools/variants/libwb_pat_base.so tools/variants/libwb_pat_chunk4.so tools/variants/libwb_pat_chunk8.so tools/variants/libwb_pat_chunk32.so tools/variants/libwb_pat_chunk64.so 2>&1 | tee gpurun_out/pat_shape.txt
