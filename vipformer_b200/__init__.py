"""vipformer_b200 -- B200-native (sm_100a) implementation of ViPFormer's pre-training hot path.

Host-side mirror of the reference's `vipformer.model` / `vipformer.preproc`
call signatures over hand-written CUDA kernels reached through the C ABI in
include/vpf.h (libvpf_b200.so).  No CPU path, no Triton, no dispatch.
"""
__version__ = "0.1.0"
