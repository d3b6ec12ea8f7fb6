// s2d_tc.cu with f16 operands: exports launch_s2d_tc_f16 / s2d_tc_init_f16 (the host-side packers
// build_s2d_host / build_stem_tc_blob exist once, in s2d_tc.cu). See ptx.cuh, "Operand type".
#define OGL_F16 1
#include "s2d_tc.cu"
