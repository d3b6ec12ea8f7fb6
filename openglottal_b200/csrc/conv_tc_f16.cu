// conv_tc.cu with f16 operands (tcgen05.mma.kind::f16 takes f16 or bf16 A/B at the same rate):
// exports launch_conv_tc_f16 / conv_tc_init_f16. See ptx.cuh, "Operand type".
#define OGL_F16 1
#include "conv_tc.cu"
