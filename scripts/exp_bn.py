import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.bench_kernels import conv_case, gemm_case  # noqa: E402
for bn in (0, 160, 224, 256):
    print("block_n", bn)
    conv_case("unet L1 640->640", 112, 30, 40, 640, 640, block_n=bn)
    conv_case("unet L1 1280->640", 112, 30, 40, 1280, 640, block_n=bn)
for bn in (0, 160, 192, 256):
    print("block_n", bn)
    conv_case("unet L0 320->320", 112, 60, 80, 320, 320, block_n=bn)
