// 16-bit tensor-core operand format of the GCRN / DiffUNet1 / DiffWave kernels and of the host packers that feed them.
// Default: IEEE fp16 (same storage and tensor-core rate as bf16, 10-bit instead of 7-bit mantissa; every fp32 -> operand
// conversion on the device saturates at +-65504 instead of overflowing, and the packers refuse weights beyond that).
// Build with -DPDSE_OP_BF16 (PDSE_OPERANDS=bf16 in the environment of build.py) for bf16 operands.
// csrc/dbaiat.cu always uses fp16: all of its operands are LayerNorm-bounded.
#pragma once
#if defined(PDSE_OP_BF16)
#define PDSE_OP_FP16 0
#else
#define PDSE_OP_FP16 1
#endif
