// upcat_tc.cu with f16 operands: exports launch_upcat_tc_f16 / upcat_tc_init_f16 (the host packer
// build_upcat_host exists once, in upcat_tc.cu). See ptx.cuh, "Operand type".
#define OGL_F16 1
#include "upcat_tc.cu"
