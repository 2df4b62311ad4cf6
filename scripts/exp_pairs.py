import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.bench_kernels import conv_case  # noqa: E402
conv_case("unet L0 320->320", 112, 60, 80, 320, 320)
conv_case("unet L0 640->320", 112, 60, 80, 640, 320)
conv_case("unet L1 640->640", 112, 30, 40, 640, 640)
conv_case("unet L1 1280->640", 112, 30, 40, 1280, 640)
conv_case("unet L2 1280->1280", 112, 15, 20, 1280, 1280)
