#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_tor_shape.py tools/variants/libwb_tor_base.so tools/variants/libwb_tor_t256.so tools/variants/libwb_tor_t128.so 2>&1 | tee gpurun_out/tor_shape2.txt
